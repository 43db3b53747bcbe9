"""Profiling driver: a 7B-shaped slice of a few layers, 256-token prefill, then un-graphed decode steps.
Launch order: prefill = ceil(256/64) calls x (6 kernels x L + 1); then each decode step = 6 x L + 1 launches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributedllm_b200 import capi, ggjt
import bench

L = int(os.environ.get("PROF_LAYERS", "4"))
steps = int(os.environ.get("PROF_STEPS", "8"))
path = bench.slice_file("7b", 0, L - 1)
sl = capi.Slice(path, 0, 512)
E = sl.n_embd
xp = bench.synth_inputs(256, E, 1)
for i in range(0, 256, 64):
    sl.forward(xp[i:i + 64])
xs = bench.synth_inputs(steps, E, 2)
bench._h2d(sl, xs[0:1])
for i in range(3):
    sl.forward_device(sl.dev_in, 1, sl.dev_out)
sl.sync()
t0 = time.perf_counter()
sl.mark(0)
for i in range(steps):
    sl.forward_device(sl.dev_in, 1, sl.dev_out)
sl.mark(1)
sl.sync()
print("decode: %.1f us/step device, %.1f us/step wall, %d layers" % (1e3 * sl.mark_elapsed_ms() / steps, 1e6 * (time.perf_counter() - t0) / steps, L))
if os.environ.get("PROF_CLASSES", "1") == "1":
    sl.profile(True)
    for i in range(steps):
        sl.forward_device(sl.dev_in, 1, sl.dev_out)
    ms, cnt = sl.profile_read()
    sl.profile(False)
    names = ("qkv", "rope", "attn", "wo", "w13", "w2", "advance")
    print("per-launch us (event-bracketed, un-graphed): " + "  ".join("%s %.1f" % (n, 1e3 * m / c) for n, m, c in zip(names, ms, cnt) if c))
sl.close()
