"""One-GPU measurements for the BASELINE.json configurations bench.py does not print (bench.py's line is config 2):
  config 4  LLaMA-7B F16, one of 8 slices (4 layers), n_ctx 2048: decode at positions 1024..; GB/s of the F16 weight stream
  config 5  LLaMA-13B Q4_0, one of 8 slices (5 layers), n_ctx 512, 8 sessions: batched step (throughput mode) vs bs 1
Output: one JSON object per config on stdout.  Device-resident buffers, CUDA events on the slice's stream."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributedllm_b200 import capi, ggjt
import bench

PEAK, PEAK_SRC = bench.measured_peak()


def timed(sl, fn, steps, warm=3):
    for _ in range(warm):
        fn()
    sl.sync()
    sl.mark(0)
    for _ in range(steps):
        fn()
    sl.mark(1)
    sl.sync()
    return sl.mark_elapsed_ms() / steps


def config4():
    sh = ggjt.SHAPES["7b"]
    L = int(os.environ.get("CFG4_LAYERS", "4"))
    p = os.path.join(bench.model_dir(), "7b_f16_layers_0_%d.bin" % (L - 1))
    if not os.path.isfile(p):
        ggjt.write_fast_f16_slice(p + ".tmp", sh, 0, L - 1, 0)
        os.replace(p + ".tmp", p)
    sl = capi.Slice(p, 0, 2048)
    E = sl.n_embd
    x = bench.synth_inputs(1024, E, 5)
    t0 = time.perf_counter()
    for i in range(0, 1024, 128):
        sl.forward(x[i:i + 128])
    prefill_s = time.perf_counter() - t0
    bench._h2d(sl, x[0:1])
    steps = 64
    ms = timed(sl, lambda: sl.forward_device(sl.dev_in, 1, sl.dev_out), steps)
    sl.profile(True)
    for _ in range(4):
        sl.forward_device(sl.dev_in, 1, sl.dev_out)
    pm, pc = sl.profile_read()
    sl.profile(False)
    classes = {n: round(1e3 * m / c, 1) for n, m, c in zip(("qkv", "rope", "attn", "wo", "w13", "w2", "advance"), pm, pc) if c}
    pos = 1024 + 3 + steps / 2
    wbytes = sl.info.weight_bytes
    kv = sl.info.kv_bytes_per_pos * pos
    out = {"config": "LLaMA-7B F16, %d-layer slice (1 of 8), n_ctx 2048, decode at p~%d" % (L, pos), "ms_per_step": ms,
           "us_per_layer": 1e3 * ms / L, "achieved_gbs": (wbytes + kv) / (ms * 1e-3) / 1e9, "peak_gbs": PEAK,
           "frac": (wbytes + kv) / (ms * 1e-3) / 1e9 / PEAK, "weight_bytes": wbytes, "kv_bytes_read": kv, "us_per_launch_by_class": classes,
           "tokens_per_s_32_layers_equiv": 1e3 / (ms * 32 / L), "prefill_1024_tok_per_s_slice": 1024 / prefill_s}
    sl.close()
    return out


def config5():
    sh = ggjt.SHAPES["13b"]
    L = 5
    p = bench.slice_file("13b", 0, L - 1)
    B = 8
    sl = capi.Slice(p, 0, 512, n_sessions=B)
    E = sl.n_embd
    x = bench.synth_inputs(256, E, 6)
    for b in range(B):
        for i in range(0, 256, 64):
            sl.session_forward(b, x[i:i + 64])
    bench._h2d(sl, x[0:B])
    ids = list(range(B))
    steps = 64
    ms8 = timed(sl, lambda: sl.batch_forward_device(ids, sl.dev_in, sl.dev_out), steps)
    sl.profile(True)
    for _ in range(4):
        sl.batch_forward_device(ids, sl.dev_in, sl.dev_out)
    pm, pc = sl.profile_read()
    sl.profile(False)
    classes = {n: round(1e3 * m / c, 1) for n, m, c in zip(("qkv", "rope", "attn", "wo", "w13", "w2", "advance"), pm, pc) if c}
    pos = 256 + 3 + steps / 2
    sl.session_clear()
    for i in range(0, 256, 64):
        sl.forward(x[i:i + 64])
    ms1 = timed(sl, lambda: sl.forward_device(sl.dev_in, 1, sl.dev_out), steps)
    wbytes = sl.info.weight_bytes
    kv = sl.info.kv_bytes_per_pos * pos
    bytes8 = wbytes + B * kv
    out = {"config": "LLaMA-13B Q4_0, 5-layer slice (1 of 8), n_ctx 512, batch 8 sessions at p~%d" % pos,
           "ms_per_batched_step": ms8, "ms_per_bs1_step": ms1, "us_per_layer_batched": 1e3 * ms8 / L, "us_per_layer_bs1": 1e3 * ms1 / L,
           "slice_tokens_per_s_batched": B * 1e3 / ms8, "slice_tokens_per_s_bs1": 1e3 / ms1,
           "achieved_gbs_batched": bytes8 / (ms8 * 1e-3) / 1e9, "frac_batched": bytes8 / (ms8 * 1e-3) / 1e9 / PEAK,
           "us_per_launch_batched_by_class": classes, "frac_bs1": (wbytes + kv) / (ms1 * 1e-3) / 1e9 / PEAK, "peak_gbs": PEAK,
           "model_tokens_per_s_8_slices_batched_serial": B * 1e3 / (ms8 * 8), "model_tokens_per_s_8_slices_pipelined_bs1": 1e3 / ms1}
    sl.close()
    return out


def config_q4_1():
    """Not a BASELINE config: the reference's other provisioning choice (`quantize q4_1`, provision.py:179), 7B, 8-layer slice,
    decode at p~270 against the same slice in Q4_0, and a parity spot check of the big shapes against the CPU oracle."""
    from oracle import oracle
    sh = ggjt.SHAPES["7b"]
    L = 8
    out = {"config": "LLaMA-7B Q4_1 vs Q4_0, %d-layer slice, n_ctx 512, decode at p~270" % L, "peak_gbs": PEAK}
    for tag, wt in (("q4_1", ggjt.T_Q4_1), ("q4_0", ggjt.T_Q4_0)):
        p = os.path.join(bench.model_dir(), "7b_%s_layers_0_%d.bin" % (tag, L - 1))
        if not os.path.isfile(p):
            ggjt.write_fast_q4_slice(p + ".tmp", sh, 0, L - 1, 0, wtype=wt)
            os.replace(p + ".tmp", p)
        sl = capi.Slice(p, 0, 512)
        x = bench.synth_inputs(256, sl.n_embd, 5)
        for i in range(0, 256, 64):
            sl.forward(x[i:i + 64])
        bench._h2d(sl, x[0:1])
        steps = 32
        ms = timed(sl, lambda: sl.forward_device(sl.dev_in, 1, sl.dev_out), steps)
        pos = 256 + 3 + steps / 2
        byts = sl.info.weight_bytes + sl.info.kv_bytes_per_pos * pos
        out[tag] = {"us_per_layer": 1e3 * ms / L, "tokens_per_s_32_layers_equiv": 1e3 / (ms * 32 / L),
                    "achieved_gbs": byts / (ms * 1e-3) / 1e9, "frac": byts / (ms * 1e-3) / 1e9 / PEAK, "weight_bytes": sl.info.weight_bytes}
        sl.close()
        if wt == ggjt.T_Q4_1:
            # parity at the full 7B widths: a 2-layer slice, prompt of 9 + 3 single-token steps, against the C restatement
            p2 = os.path.join(bench.model_dir(), "7b_q4_1_layers_0_1.bin")
            if not os.path.isfile(p2):
                ggjt.write_fast_q4_slice(p2 + ".tmp", sh, 0, 1, 0, wtype=wt)
                os.replace(p2 + ".tmp", p2)
            g, c = capi.Slice(p2, 0, 64), oracle.PortSlice(p2, 64)
            bad = 0
            for n in (9, 1, 1, 1):
                xx = bench.synth_inputs(n, g.n_embd, 40 + n)
                bad += int((np.ascontiguousarray(g.forward(xx)).view(np.uint32) != np.ascontiguousarray(c.forward(xx)).view(np.uint32)).sum())
            out["parity_mismatches_7b_2_layers"] = bad
            g.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["4", "5"]
    if "q4_1" in which:
        print(json.dumps(config_q4_1()))
    if "5" in which:
        print(json.dumps(config5()))
    if "4" in which:
        print(json.dumps(config4()))
