import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from yt_corpus import make_youtube
from distributed_crawler_b200 import abi
from distributed_crawler_b200.engine import Engine
from oracle.pyoracle import Oracle
b, vids, chans = make_youtube(5, seed=16)
fl = abi.RUN_JSONL | abi.RUN_LINKS
ro = Oracle().youtube(b, fl); rg = Engine().youtube(b, fl)
print(ro.link_off, rg.link_off)
for nm, r in (("o", ro), ("g", rg)):
    for k in range(len(r.links)):
        L = r.links[k]
        print(nm, k, bytes(L['name'][:int(L['len'])]), int(L['len']), int(L['src']), int(L['flags']), int(L['filter_reason']), bytes(L['name']))
