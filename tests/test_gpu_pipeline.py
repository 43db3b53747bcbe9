"""GPU, 2 ranks: the layer-slice pipeline (b200_pipeline_*) against the un-sliced model on one GPU, once per hand-off
transport: the peer-memory mailboxes (store + flag over NVLink inside the step graph) and the single ncclSend/ncclRecv."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, ctypes as C
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from distributedllm_b200 import capi, ggjt
from distributedllm_b200.pipeline import layer_ranges
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sh = ggjt.SHAPES["tiny128"]
d = %(tmp)r
WT = int(os.environ.get("B200_TEST_WTYPE", str(ggjt.T_Q4_0)))
a, b = layer_ranges(sh.n_layer, world)[rank]
p = os.path.join(d, "s_%%d_%%d.bin" %% (a, b))
if not os.path.exists(p):
    ggjt.write_synth_slice(p, sh, a, b, WT, seed=0)
sl = capi.Slice(p, local, 64, n_sessions=4)
lib = capi.lib()
from distributedllm_b200.pipeline import join_pipeline, torch_collectives
bcast, gather = torch_collectives(dist, torch.device("cuda", local))
want_peer = os.environ.get("B200_PP_PEER", "1") != "0"
transport = join_pipeline(sl, rank, world, bcast, gather, peer=want_peer)
lib.b200_pipeline_result.restype = C.c_void_p; lib.b200_pipeline_result.argtypes = [C.c_void_p]
_rt = C.CDLL("libcudart.so.12")
class _Rt:                                   # cudaMemcpy from pageable memory may return before the DMA tail lands, and the
    def cudaMemcpy(self, *a):                # slice stream is non-blocking (not ordered behind the legacy stream): sync
        rc = _rt.cudaMemcpy(*a)
        _rt.cudaDeviceSynchronize()
        return rc
cudart = _Rt()
rng = np.random.default_rng(21)
ok = True
if rank == 0:
    whole = os.path.join(d, "whole.bin"); ggjt.write_synth_slice(whole, sh, 0, sh.n_layer - 1, WT, seed=0)
    ref = capi.Slice(whole, local, 64, n_sessions=4)
for n in (7, 1, 1, 5, 1):
    x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
    if rank == 0:
        assert cudart.cudaMemcpy(C.c_void_p(sl.dev_in), C.c_void_p(x.ctypes.data), C.c_size_t(x.nbytes), 1) == 0
    capi.check(lib.b200_pipeline_step(sl.handle, C.c_void_p(sl.dev_in), n, 1))
    sl.sync()
    if rank == 0:
        out = np.empty_like(x)
        assert cudart.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(lib.b200_pipeline_result(sl.handle)), C.c_size_t(out.nbytes), 2) == 0
        want = ref.forward(x)
        ok = ok and bool((out.view(np.uint32) == want.view(np.uint32)).all())
# sessions through the pipeline: ragged prompts into sessions 1..3, then batched steps (one token per session)
def fetch(n):
    out = np.empty((n, sh.n_embd), np.float32)
    assert cudart.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(lib.b200_pipeline_result(sl.handle)), C.c_size_t(out.nbytes), 2) == 0
    return out
for k, n in ((1, 3), (2, 9), (3, 1)):
    x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
    if rank == 0:
        assert cudart.cudaMemcpy(C.c_void_p(sl.dev_in), C.c_void_p(x.ctypes.data), C.c_size_t(x.nbytes), 1) == 0
    capi.check(lib.b200_pipeline_step_session(sl.handle, k, C.c_void_p(sl.dev_in), n, 1))
    sl.sync()
    if rank == 0:
        ok = ok and bool((fetch(n).view(np.uint32) == ref.session_forward(k, x).view(np.uint32)).all())
ids = np.array([3, 1, 2], np.int32)
for step in range(4):
    x = rng.standard_normal((3, sh.n_embd), dtype=np.float32)
    if rank == 0:
        assert cudart.cudaMemcpy(C.c_void_p(sl.dev_in), C.c_void_p(x.ctypes.data), C.c_size_t(x.nbytes), 1) == 0
    capi.check(lib.b200_pipeline_step_batch(sl.handle, C.c_void_p(ids.ctypes.data), 3, C.c_void_p(sl.dev_in), 1))
    sl.sync()
    if rank == 0:
        got = fetch(3)
        for j, k in enumerate(ids):
            ok = ok and bool((got[j].view(np.uint32) == ref.session_forward(int(k), x[j:j + 1])[0].view(np.uint32)).all())
# decode steps back to back with no host synchronisation in between (graph replays; the mailbox slots must not be overrun)
if rank == 0:
    ref.clear_context()
sl.clear_context()
dist.barrier()
xs = rng.standard_normal((24, sh.n_embd), dtype=np.float32)
outs = []
for i in range(24):
    if rank == 0:
        assert cudart.cudaMemcpy(C.c_void_p(sl.dev_in), C.c_void_p(xs[i:i + 1].ctypes.data), C.c_size_t(xs[i:i + 1].nbytes), 1) == 0
    capi.check(lib.b200_pipeline_step(sl.handle, C.c_void_p(sl.dev_in), 1, 1))
    if rank == 0:
        sl.sync()
        outs.append(fetch(1))
sl.sync()
if rank == 0:
    for i in range(24):
        ok = ok and bool((outs[i].view(np.uint32) == ref.forward(xs[i:i + 1]).view(np.uint32)).all())
# throughput mode: ring = 2 steps for three sessions back to back, results collected later in issue order (b200_pipeline_collect);
# a session's next token is issued only after its previous result is back -- same tokens as stepping the sessions one by one
dist.barrier()
lib.b200_pipeline_collect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
res = torch.zeros((3, sh.n_embd), dtype=torch.float32, device="cuda")
sess = [1, 2, 3]
xs3 = rng.standard_normal((4, 3, sh.n_embd), dtype=np.float32)            # 4 rounds x 3 sessions
pipelined = np.zeros_like(xs3)
window = world if transport == "peer" else 1                               # over NCCL a step is collected before the next is issued
buf = torch.zeros((3, sh.n_embd), dtype=torch.float32, device="cuda")
for rnd in range(4):
    if rank == 0:
        buf.copy_(torch.from_numpy(xs3[rnd]))
        torch.cuda.synchronize()
    issued = 0
    for j in range(3 + window):
        if j >= window and rank == 0:
            capi.check(lib.b200_pipeline_collect(sl.handle, 1, C.c_void_p(res.data_ptr() + 4 * sh.n_embd * (j - window))))
        if j < 3:
            capi.check(lib.b200_pipeline_step_session(sl.handle, sess[j], C.c_void_p(buf.data_ptr() + 4 * sh.n_embd * j), 1, 2))
    sl.sync()
    if rank == 0:
        torch.cuda.synchronize()
        pipelined[rnd] = res.cpu().numpy()
    dist.barrier()
if rank == 0:
    for rnd in range(4):
        for j, k in enumerate(sess):
            want = ref.session_forward(k, xs3[rnd, j:j + 1])[0]
            ok = ok and bool((pipelined[rnd, j].view(np.uint32) == want.view(np.uint32)).all())
dist.barrier()
capi.check(lib.b200_pipeline_destroy(sl.handle))
err = lib.b200_pipeline_error(sl.handle)
if rank == 0:
    print(("PIPELINE_OK" if ok and not err else "PIPELINE_MISMATCH") + " transport=" + transport)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("peer,fold,wtype", [(1, 1, 2), (1, 0, 2), (0, 0, 2), (1, 1, 3)],
                         ids=["peer-folded-into-matmuls", "peer-send-recv-kernels", "nccl", "peer-folded-q4_1"])
def test_two_gpu_pipeline_bit_exact(tmp_path, peer, fold, wtype):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "tmp": str(tmp_path)})
    env = dict(os.environ, B200_PP_PEER=str(peer), B200_PP_FOLD=str(fold), B200_TEST_WTYPE=str(wtype))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(29533 + 2 * peer + fold + 4 * (wtype == 3)), str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert "PIPELINE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
    assert ("transport=peer" if peer else "transport=nccl") in out.stdout, out.stdout[-500:]
