"""Profiling driver for BASELINE config 4: a 1-layer LLaMA-7B F16 slice at n_ctx 2048, 1024-token prefill, un-graphed decode
steps at p ~ 1024 (for `ncu -k regex:k_gemv_f16|k_attn128`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200_GRAPH", "0")
from distributedllm_b200 import capi, ggjt
import bench
sh = ggjt.SHAPES["7b"]
p = os.path.join(bench.model_dir(), "7b_f16_layers_0_0.bin")
if not os.path.isfile(p):
    ggjt.write_fast_f16_slice(p + ".tmp", sh, 0, 0, 0)
    os.replace(p + ".tmp", p)
sl = capi.Slice(p, 0, 2048)
x = bench.synth_inputs(1024, sh.n_embd, 5)
for i in range(0, 1024, 128):
    sl.forward(x[i:i + 128])
bench._h2d(sl, x[0:1])
for i in range(int(os.environ.get("PROF_STEPS", "4"))):
    sl.forward_device(sl.dev_in, 1, sl.dev_out)
sl.sync()
print("done", sl.n_past)
