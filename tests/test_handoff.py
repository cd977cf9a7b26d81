"""SURVEY 8f rank 3 — frontier -> validator hand-off: the pending_edges rows of a tandem batch.

CPU: the oracle against a direct Python restatement of the reference loop (crawl/runner.go:1230-1306: self skip,
sm.IsInvalidChannel with its 30-day TTL, FilterUsername, seenInBatch, one row per surviving edge) and of the validator's
two cache look-ups (crawl/validator.go:205-226).  GPU: tgi_pending_edges against the oracle, row for row."""
import numpy as np
import pytest

from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import names_to_keys32
from oracle import pyoracle
from oracle.pyoracle import Oracle

TANDEM = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_FILTER | abi.RUN_SKIP_SELF | abi.RUN_SKIP_INVALID
NOW = 1_760_000_000
TTL = 30 * 24 * 3600


def _sets(batch_links):
    """pick invalid / discovered channels among the names the corpus really produces"""
    names = sorted({bytes(l["name"][: l["len"]]) for l in batch_links})
    inv = names[0::7]                      # every 7th name is a known-invalid channel ...
    stamps = [NOW - (i % 3) * 20 * 24 * 3600 for i in range(len(inv))]  # ... marked 0, 20 or 40 days ago (40 > TTL: expired)
    disc = names[3::5]
    return inv, stamps, disc


def _reference_rows(o_plain_links, link_off, recs, chan_names, inv, stamps, disc):
    """crawl/runner.go:1230-1306 + crawl/validator.go:205-226 over the per-record outlinks (first-insertion order)"""
    inv_t = dict(zip(inv, stamps))
    is_invalid = lambda nm: nm in inv_t and NOW - inv_t[nm] < TTL
    disc_s, seen, rows = set(disc), set(), []
    for r in range(len(recs)):
        owner = chan_names[int(recs["chan_idx"][r])]
        for k in range(int(link_off[r]), int(link_off[r + 1])):
            l = o_plain_links[k]
            nm = bytes(l["name"][: l["len"]])
            if nm == owner:                      # :1231 self reference
                continue
            if is_invalid(nm):                   # :1247
                continue
            if l["filter_reason"] != 0:          # :1261 FilterUsername
                continue
            if nm in seen:                       # :1267
                continue
            seen.add(nm)
            status = abi.EDGE_INVALID_CACHED if is_invalid(nm) else abi.EDGE_DUPLICATE if nm in disc_s else abi.EDGE_PENDING
            rows.append((nm, r, int(recs["chan_idx"][r]), int(l["src"]), status))
    return rows


def _chan_names(b):
    out = []
    for ch in b.chans:
        o = int(ch["str_off"]) + int(ch["title_len"])
        out.append(b.chan_strs[o:o + int(ch["name_len"])].tobytes())
    return out


def _as_tuples(rows):
    return [(bytes(e["destination"][: e["dest_len"]]), int(e["record"]), int(e["chan_idx"]), int(e["source_type"]), int(e["status"])) for e in rows]


def test_oracle_rows_follow_the_reference_loop():
    c = Corpus(30_000, profile=3, nthreads=2)
    plain = Oracle().telegram(c.batch, abi.RUN_LINKS)
    inv, stamps, disc = _sets(plain.links)
    o = Oracle()
    o.set_add(abi.SET_INVALID, names_to_keys32(inv), np.array(stamps, np.int64))
    o.set_add(abi.SET_DISCOVERED, names_to_keys32(disc))
    o.set_now(NOW)
    r = o.telegram(c.batch, TANDEM, nthreads=3)
    got = _as_tuples(o.pending_edges(NOW))
    want = _reference_rows(plain.links, plain.link_off, c.batch.recs, _chan_names(c.batch), inv, stamps, disc)
    assert got == want and len(got) == r.n_new > 100
    st = [g[4] for g in got]
    assert st.count(abi.EDGE_DUPLICATE) > 0 and st.count(abi.EDGE_PENDING) > 0 and st.count(abi.EDGE_INVALID_CACHED) == 0
    assert (r.links["flags"] & abi.LF_INVALID).sum() > 0
    # expired invalid marks do not filter: those names are edges again
    expired = {nm for nm, t in zip(inv, stamps) if NOW - t >= TTL}
    assert expired & {g[0] for g in got}


@pytest.mark.gpu
def test_gpu_pending_edges_match_oracle():
    from distributed_crawler_b200.engine import Engine
    c = Corpus(200_000, profile=3)
    plain = Oracle().telegram(c.batch, abi.RUN_LINKS, nthreads=8)
    inv, stamps, disc = _sets(plain.links)
    o, e = Oracle(), Engine()
    for x in (o, e):
        x.set_add(abi.SET_INVALID, names_to_keys32(inv), np.array(stamps, np.int64))
        x.set_add(abi.SET_DISCOVERED, names_to_keys32(disc))
        x.set_now(NOW)
    assert e.set_size(abi.SET_INVALID) == len(inv) and e.set_size(abi.SET_DISCOVERED) == len(disc)
    for part in (c.batch.slice(0, 120_000), c.batch.slice(120_000, 200_000)):  # the dedup set carries over
        ro = o.telegram(part, TANDEM, nthreads=8)
        e.telegram_submit(1, part, TANDEM)
        rg = e.telegram_wait(1, copy=True)
        assert np.array_equal(ro.links, rg.links) and ro.n_new == rg.n_new
        # a channel marked invalid after the batch ran shows up as cached-invalid in the rows
        late = bytes(ro.links[ro.links["flags"] & abi.LF_NEW != 0][0]["name"][:32]).rstrip(b"\\0")
        rows_o, rows_g = o.pending_edges(NOW), e.pending_edges(1, NOW)
        assert np.array_equal(rows_o, rows_g) and len(rows_g) == rg.n_new
        e.release(1)
    e.close()


STATUS = {"pending": abi.EDGE_PENDING, "duplicate": abi.EDGE_DUPLICATE, "invalid": abi.EDGE_INVALID_CACHED}


def _tandem_case(x, t):
    """one reference tandem vector through x (Oracle or Engine): names of the edges that would be inserted"""
    from distributed_crawler_b200.pack import Channel, pack_telegram
    from helpers import msg
    if t.get("invalid"):
        x.set_add(abi.SET_INVALID, names_to_keys32([n.encode() for n in t["invalid"]]), np.zeros(len(t["invalid"]), np.int64))
    x.set_now(NOW)
    b = pack_telegram([msg("messageText", t["text"], [tuple(e) for e in t["entities"]])],
                      [Channel(name=t["owner_url"], username=t["owner_url"])])
    return b


def test_reference_tandem_vectors_oracle(vectors):
    """crawl/runner_tandem_test.go: WithEdges (two rows), NoEdges (none), InvalidChannelSkipped (IsInvalidChannel -> no
    InsertPendingEdge) — the rows tgi_pending_edges would hand to the one INSERT."""
    for name, t in vectors["tandem"].items():
        o = Oracle()
        b = _tandem_case(o, t)
        r = o.telegram(b, TANDEM)
        rows = o.pending_edges(NOW)
        got = [bytes(e["destination"][: e["dest_len"]]).decode() for e in rows]
        assert got == t["expected_edges"], f'{name} (crawl/runner_tandem_test.go:{t["go_line"]})'
        assert r.n_new == len(t["expected_edges"])
        for inv in t.get("invalid", []):
            hit = [l for l in r.links if bytes(l["name"][: l["len"]]).decode() == inv]
            assert hit and all(l["flags"] & abi.LF_INVALID for l in hit)


def _validator_case(x, v):
    from distributed_crawler_b200.pack import Channel, pack_telegram
    from helpers import msg
    d = v["destination"]
    b = pack_telegram([msg("messageText", "see t.me/" + d)], [Channel(name="source_channel", username="source_channel")])
    plain = abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_FILTER | abi.RUN_SKIP_SELF  # the edge was inserted BEFORE the caches knew
    return b, plain, d


def test_reference_validator_cache_vectors_oracle(vectors):
    """crawl/validator_test.go TestValidateSingleEdge_{Valid,AlreadyInvalid,AlreadyDiscovered}: what the two cache
    look-ups of validateSingleEdge (validator.go:205-226) answer, as the status column of the packed rows."""
    for v in vectors["validator_cache"]:
        o = Oracle()
        b, flags, d = _validator_case(o, v)
        o.telegram(b, flags)
        if v["invalid"]:
            o.set_add(abi.SET_INVALID, names_to_keys32([d.encode()]), np.zeros(1, np.int64))
        if v["discovered"]:
            o.set_add(abi.SET_DISCOVERED, names_to_keys32([d.encode()]))
        rows = o.pending_edges(NOW)
        assert len(rows) == 1 and bytes(rows[0]["destination"][: rows[0]["dest_len"]]).decode() == d
        assert rows[0]["status"] == STATUS[v["status"]], f'{v["name"]} ({v["go_file"]}:{v["go_line"]})'


@pytest.mark.gpu
def test_reference_tandem_and_validator_vectors_gpu(vectors):
    from distributed_crawler_b200.engine import Engine
    for name, t in vectors["tandem"].items():
        e = Engine()
        b = _tandem_case(e, t)
        e.telegram_submit(0, b, TANDEM)
        r = e.telegram_wait(0, copy=True)
        rows = e.pending_edges(0, NOW)
        assert [bytes(x["destination"][: x["dest_len"]]).decode() for x in rows] == t["expected_edges"], name
        assert r.n_new == len(t["expected_edges"])
        e.release(0)
        e.close()
    for v in vectors["validator_cache"]:
        e = Engine()
        b, flags, d = _validator_case(e, v)
        e.telegram_submit(0, b, flags)
        e.telegram_wait(0)
        if v["invalid"]:
            e.set_add(abi.SET_INVALID, names_to_keys32([d.encode()]), np.zeros(1, np.int64))
        if v["discovered"]:
            e.set_add(abi.SET_DISCOVERED, names_to_keys32([d.encode()]))
        rows = e.pending_edges(0, NOW)
        assert len(rows) == 1 and rows[0]["status"] == STATUS[v["status"]], v["name"]
        e.release(0)
        e.close()
