"""F16 slice, exact prefill: device-timed tokens/s of 128-token calls (BASELINE config 4 shapes, 4-layer slice), with the
multi-column kernel (default) and with one column per CTA (B200_F16_MC=0).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributedllm_b200 import capi, ggjt
import bench

L = 4
sh = ggjt.SHAPES["7b"]
p = os.path.join(bench.model_dir(), "7b_f16_layers_0_%d.bin" % (L - 1))
if not os.path.isfile(p):
    ggjt.write_fast_f16_slice(p + ".tmp", sh, 0, L - 1, 0)
    os.replace(p + ".tmp", p)
out = {"config": "LLaMA-7B F16, %d-layer slice, n_ctx 2048, exact prefill in 128-token calls" % L}
for mc in ("1", "8", "0"):
    os.environ["B200_F16_MC"] = mc
    sl = capi.Slice(p, 0, 2048)
    x = bench.synth_inputs(128, sl.n_embd, 3)
    bench._h2d(sl, x)
    sl.forward_device(sl.dev_in, 128, sl.dev_out)            # warm-up: module load, attributes
    sl.clear_context()
    res = {}
    for label, calls in (("positions_0_511", 4), ("positions_512_1023", 4)):
        sl.sync(); sl.mark(0)
        for _ in range(calls):
            sl.forward_device(sl.dev_in, 128, sl.dev_out)
        sl.mark(1); sl.sync()
        ms = sl.mark_elapsed_ms()
        res[label] = {"tok_per_s_slice": 128 * calls / (ms * 1e-3), "us_per_token_layer": 1e3 * ms / (128 * calls * L),
                      "tok_per_s_32_layers_equiv": 128 * calls / (ms * 1e-3) * L / 32}
    out[{"1": "multi_column_4", "8": "multi_column_8", "0": "one_column_per_cta"}[mc]] = res
    sl.close()
print(json.dumps(out))
