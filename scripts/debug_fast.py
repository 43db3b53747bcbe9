import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributedllm_b200 import capi, ggjt
sh = ggjt.SHAPES["tiny128b"]; d = tempfile.mkdtemp(); p = os.path.join(d, "m.bin")
ggjt.write_synth_slice(p, sh, 0, 0, ggjt.T_Q4_0, 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = np.random.default_rng(4).standard_normal((N, sh.n_embd), dtype=np.float32)
a, b = capi.Slice(p, 0, 512), capi.Slice(p, 0, 512)
b.set_fast_prefill(True, 32)
ya, yb = a.forward(x), b.forward(x)
E, FF = sh.n_embd, sh.n_ff
def rel(u, v): return float(np.sqrt(np.mean((u - v) ** 2)) / (np.sqrt(np.mean(u ** 2)) + 1e-30))
for name, which, cnt in (("qkv", 0, N * 3 * E), ("att", 1, N * E), ("ffin", 2, N * E), ("gate", 3, N * FF)):
    u, v = a.debug_read(which, cnt), b.debug_read(which, cnt)
    print("%-5s rel rms %.3e  max|d| %.3e  (max|ref| %.3e)" % (name, rel(u, v), np.abs(u - v).max(), np.abs(u).max()))
print("out   rel rms %.3e" % rel(ya, yb))
