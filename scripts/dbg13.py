import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from distributedllm_b200 import capi, ggjt
from oracle import oracle
sh = ggjt.SHAPES["13b"]; E = sh.n_embd
p = bench.slice_file("13b", 0, 1)
gpu = capi.Slice(p, 0, 512, n_sessions=2)
ref = oracle.RefSlice(p, 16, 512)
xp = bench.synth_inputs(256, E, 6); xs = bench.synth_inputs(2, E, 7)
def bits(a): return np.ascontiguousarray(a, np.float32).view(np.uint32)
bad = 0
outs_g = [gpu.session_forward(0, xp[i:i+64]) for i in range(0, 256, 64)]
outs_r = [ref.forward(xp[i:i+32]) for i in range(0, 256, 32)]
g = np.concatenate(outs_g); r = np.concatenate(outs_r)
print("prefill mismatches per 32-token chunk:", [(int((bits(g[i:i+32]) != bits(r[i:i+32])).sum())) for i in range(0, 256, 32)])
a, b = gpu.session_forward(0, xs[0:1]), ref.forward(xs[0:1])
print("decode mismatches", int((bits(a) != bits(b)).sum()), float(np.abs(a-b).max()), float(np.abs(b).max()))
