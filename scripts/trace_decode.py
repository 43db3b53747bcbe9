"""Timeline of one graphed decode step from in-kernel %globaltimer stamps (B200_TRACE=1)."""
import os, sys, ctypes as C
os.environ["B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributedllm_b200 import capi
import bench
L = int(os.environ.get("PROF_LAYERS", "4"))
path = bench.slice_file("7b", 0, L - 1)
sl = capi.Slice(path, 0, 512)
xp = bench.synth_inputs(256, sl.n_embd, 1)
for i in range(0, 256, 64):
    sl.forward(xp[i:i + 64])
bench._h2d(sl, xp[0:1])
for i in range(6):
    sl.forward_device(sl.dev_in, 1, sl.dev_out)
sl.sync()
lib = capi.lib()
buf = np.zeros((512, 1024, 8), np.uint64); cls = np.zeros(512, np.int32); ctas = np.zeros(512, np.int32)
lib.b200_debug_trace_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
n = lib.b200_debug_trace_read(sl.handle, buf.ctypes.data, cls.ctypes.data, ctas.ctypes.data, 512)
names = ["qkv", "rope", "attn", "wo", "w13", "w2", "adv"]
# the decode graph's launches are the LAST 5*L entries
first = n - 5 * L
t_base = None
print("launch  kernel  ctas | first_start last_start | first_ready(dep) | prologue_done(max) | last_copy_issued(max) | end(max)   [us from step start]")
for i in range(first, n):
    k = ctas[i]; d = buf[i, :k].astype(np.int64)
    if t_base is None: t_base = d[:, 0].min()
    f = lambda col, fn: (fn(d[:, col][d[:, col] > 0]) - t_base) / 1e3 if (d[:, col] > 0).any() else float("nan")
    print("%3d %6s %5d | %8.2f %8.2f | %8.2f | %8.2f | %8.2f | %8.2f" % (i - first, names[cls[i]], k, f(0, np.min), f(0, np.max), f(1, np.min), f(2, np.max), f(4, np.max), f(3, np.max)))
print("\nper-CTA durations [us]: prologue = t2-t1, main = t3-t2 (min / median / p90 / max)")
for i in range(first, min(n, first + 5)):
    k = ctas[i]; d = buf[i, :k].astype(np.int64)
    if (d[:, 2] > 0).any():
        pro = (d[:, 2] - d[:, 1]) / 1e3; main = (d[:, 3] - d[:, 2]) / 1e3
        q = lambda a: "%.2f / %.2f / %.2f / %.2f" % (a.min(), np.median(a), np.percentile(a, 90), a.max())
        print("%6s prologue %s | main %s | t1 spread %.2f" % (names[cls[i]], q(pro), q(main), (d[:, 1].max() - d[:, 1].min()) / 1e3))
        if names[cls[i]] == "attn":
            m = lambda a_, b_: np.median((d[:, a_] - d[:, b_]) / 1e3)
            print("        median phases: wait->rope/q %.2f | scores %.2f | cluster sync %.2f | softmax %.2f | V.p %.2f (incl. sync) | finish %.2f"
                  % (m(2, 1), m(4, 2), m(5, 4), m(6, 5), m(7, 6), m(3, 7)))
        if names[cls[i]] != "attn" and (d[:, 7] > 0).any():
            u = buf[i, :k, 7]
            wait, tot = (u >> np.uint64(32)).astype(np.float64), (u & np.uint64(0xFFFFFFFF)).astype(np.float64)
            fr = wait / np.maximum(tot, 1.0)
            print("        main loop of warp 0: %.0f cycles median, of which waiting for weight stages %.0f%% / %.0f%% / %.0f%% (min / median / max)"
                  % (np.median(tot), 100 * fr.min(), 100 * np.median(fr), 100 * fr.max()))
        if names[cls[i]] == "attn":
            pass
        elif (d[:, 5] > 0).any():
            print("        prologue split: load+sum %s | reduce+scale %s | quant+bar %s" % (q((d[:, 5] - d[:, 1]) / 1e3), q((d[:, 6] - d[:, 5]) / 1e3), q((d[:, 2] - d[:, 6]) / 1e3)))
    else:
        tot = (d[:, 3] - d[:, 1]) / 1e3
        print("%6s total %.2f / %.2f / %.2f" % (names[cls[i]], tot.min(), np.median(tot), tot.max()))
        m = lambda a_, b_: np.median((d[:, a_] - d[:, b_]) / 1e3)
        print("        median phases: rope/q %.2f | scores %.2f | cluster sync %.2f | softmax %.2f | V.p+sync %.2f | finish %.2f" % (m(2, 1), m(4, 2), m(5, 4), m(6, 5), m(7, 6), m(3, 7)))
sl.close()
