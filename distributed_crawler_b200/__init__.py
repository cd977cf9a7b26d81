"""distributed_crawler_b200 — B200-native message-ingest engine for the distributed-crawler hot path
(parse -> link-extract -> filter/dedup -> JSONL).  See DESIGN.md.

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of include/tgingest.h);
this package is the thin host mirror of the reference interfaces used by tests and bench.py.
"""
from . import abi  # noqa: F401
