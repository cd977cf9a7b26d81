#!/usr/bin/env bash
# Usage: make_fixtures.sh <repo root of the B200 engine>   (run from the reference checkout, see README.md)
set -euo pipefail
REPO="$1"
OUT="$REPO/tests/golden/go_fixtures"
python3 "$REPO/go_ref/dump_batch.py" "$OUT"
CGO_ENABLED=1 go build -o /tmp/goref ./cmd/goref
for f in "$OUT"/*.tgb "$OUT"/*.ytb; do
  [ -e "$f" ] || continue
  /tmp/goref -in "$f" -out "${f%.*}.jsonl" -links "${f%.*}.links.txt"
done
ls -l "$OUT"
