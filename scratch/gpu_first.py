import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
from distributed_crawler_b200.pack import *
from oracle.pyoracle import Oracle

def compare(b, fl, label, **cfgkw):
    o = Oracle(**cfgkw); e = Engine(**cfgkw)
    ro = o.telegram(b, fl); rg = e.telegram(b, fl)
    ok = True
    if not np.array_equal(ro.status, rg.status):
        bad = np.nonzero(ro.status != rg.status)[0]; print(label, "STATUS MISMATCH", len(bad), bad[:10], ro.status[bad[:10]], rg.status[bad[:10]]); ok=False
    if fl & abi.RUN_JSONL:
        if not np.array_equal(ro.line_off, rg.line_off):
            bad = np.nonzero(np.diff(ro.line_off.astype(np.int64)) != np.diff(rg.line_off.astype(np.int64)))[0]
            print(label, "LINELEN MISMATCH", len(bad), bad[:10]); ok=False
            for i in bad[:6]:
                a, c = ro.line(i), rg.line(i)
                j = next((k for k in range(min(len(a),len(c))) if a[k] != c[k]), min(len(a),len(c)))
                print("  rec", i, "len", len(a), len(c), "first diff at", j, a[max(0,j-80):j+40], "|||", c[max(0,j-80):j+40])
        elif not np.array_equal(ro.jsonl, rg.jsonl):
            nb = 0
            for i in range(b.n):
                if ro.line(i) != rg.line(i):
                    nb += 1
                    if nb <= 3:
                        a, c = ro.line(i), rg.line(i)
                        j = next(k for k in range(len(a)) if a[k] != c[k])
                        print(label, "LINE MISMATCH rec", i, "at", j, a[max(0,j-60):j+60], "|||", c[max(0,j-60):j+60])
            print(label, "lines differing:", nb); ok=False
    if fl & abi.RUN_LINKS:
        if not (np.array_equal(ro.link_off, rg.link_off) and np.array_equal(ro.links, rg.links)):
            print(label, "LINKS MISMATCH", ro.n_links if hasattr(ro,'n_links') else len(ro.links), len(rg.links)); ok=False
            for i in range(b.n):
                if ro.record_links(i) != rg.record_links(i):
                    print("  rec", i, ro.record_links(i), rg.record_links(i)); break
            if len(ro.links)==len(rg.links):
                bad=np.nonzero(ro.links != rg.links)[0]; print("  differing link idx", bad[:5], ro.links[bad[:3]], rg.links[bad[:3]])
    if fl & abi.RUN_FRONTIER:
        fo, fg = o.frontier_export(), e.frontier_export()
        if not np.array_equal(fo, fg): print(label, "FRONTIER MISMATCH", fo.shape, fg.shape); ok=False
        if ro.n_new != rg.n_new: print(label, "n_new", ro.n_new, rg.n_new); ok=False
    print(label, "OK" if ok else "FAIL", "n=", b.n, "kernel_ms", rg.kernel_ms, "parse", rg.parse_ms, "emit", rg.emit_ms, "launches", rg.gpu_launches, "jsonl", rg.jsonl_len)
    return ok

m=[Message(text=FormattedText("Check out https://t.me/channelname for news & <b>"), reactions=[("👍",3),("❤",2)]),
   Message(text=FormattedText("Привет @testchan", [TextEntity(7,9,"mention")])),
   Message(content_type="messageSticker")]
ALL = abi.RUN_JSONL|abi.RUN_LINKS|abi.RUN_FRONTIER|abi.RUN_SKIP_SELF
compare(pack_telegram(m), ALL, "tiny", crawl_label=b"l<b>l")
for n, prof in ((2000, 2), (100000, 2), (100000, 3), (20000, 1)):
    c = Corpus(n, profile=prof)
    compare(c.batch, ALL, f"corpus{prof}-{n}")
    compare(c.batch, abi.RUN_LINKS|abi.RUN_FRONTIER|abi.RUN_FILTER|abi.RUN_SKIP_SELF, f"corpus{prof}-{n}-tandem", tz_offset_sec=3600, min_post_date=1700000000)
