"""Hand-off latency in isolation (run with torch.distributed.run, N ranks): bare ring hops of 1 / 8 / 64 rows of n_embd floats,
peer mailboxes vs NCCL, no layers in between.  Prints microseconds per hop."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from distributedllm_b200 import capi, ggjt  # noqa: E402
from distributedllm_b200.pipeline import join_pipeline, torch_collectives  # noqa: E402
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sh = ggjt.SHAPES["7b"]
sl = capi.Slice(bench.slice_file("7b", rank, rank), local, 512)          # one layer per rank: only the mailbox matters here
bcast, gather = torch_collectives(dist, torch.device("cuda", local))
transport = join_pipeline(sl, rank, world, bcast, gather)
lib = capi.lib()
out = {"world": world, "n_embd": sh.n_embd, "us_per_hop": {}}
for name, peer in (("peer_mailbox", 1), ("nccl", 0)):
    if peer and transport != "peer":
        continue
    capi.check(lib.b200_pipeline_set_transport(sl.handle, peer))
    dist.barrier()
    for rows in (1, 8, 64):
        us = C.c_float()
        sl.sync(); dist.barrier()
        capi.check(lib.b200_pipeline_pingpong(sl.handle, rows, 200, C.byref(us)))
        t = torch.tensor([us.value], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["us_per_hop"]["%s_%drows" % (name, rows)] = round(float(t[0]) / world, 2)
capi.check(lib.b200_pipeline_set_transport(sl.handle, 1 if transport == "peer" else 0))
if rank == 0:
    print(json.dumps(out), flush=True)
capi.check(lib.b200_pipeline_destroy(sl.handle))
dist.barrier()
dist.destroy_process_group()
