"""A few decode steps of an 8-layer 7B Q4_0 slice at p ~ 260 for `ncu -k regex:k_decode_persistent` (B200_PERSIST=1) or
the multi-kernel step (B200_PERSIST=0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200_PERSIST", "1")
import bench  # noqa: E402
from distributedllm_b200 import capi, ggjt  # noqa: E402

sh = ggjt.SHAPES["7b"]
nl = int(os.environ.get("PROF_LAYERS", "8"))
sl = capi.Slice(bench.slice_file("7b", 0, nl - 1), 0, 512)
xp = bench.synth_inputs(256, sh.n_embd, 1)
for i in range(0, 256, 64):
    sl.forward(xp[i:i + 64])
bench._h2d(sl, xp[0:1])
for i in range(int(os.environ.get("PROF_STEPS", "4"))):
    sl.forward_device(sl.dev_in, 1, sl.dev_out)
sl.sync()
print("done", sl.n_past)
