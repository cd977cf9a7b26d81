import os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from distributed_crawler_b200 import abi
from distributed_crawler_b200.corpus import Corpus
from distributed_crawler_b200.engine import Engine
from distributed_crawler_b200.frontier_merge import EngineFrontier, merge_frontier
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 200000
c = Corpus(n, first=rank * n, profile=3)
e = Engine(device=local)
fs = EngineFrontier(e, dev)
r = e.telegram(c.batch, abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF)
print(rank, "local", e.frontier_size(), r.n_new, flush=True)
g, upto = merge_frontier(fs, 0)
print(rank, "global", g, flush=True)
if rank == 0:
    e1 = Engine(device=local)
    e1.telegram(Corpus(2 * n, profile=3).batch, abi.RUN_LINKS | abi.RUN_FRONTIER | abi.RUN_SKIP_SELF)
    print("single-process union", e1.frontier_size(), "sets equal:", {bytes(x) for x in e1.frontier_export()} == {bytes(x) for x in e.frontier_export()})
dist.barrier(); dist.destroy_process_group()
