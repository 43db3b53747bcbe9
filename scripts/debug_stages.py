"""Debug aid: compare the GPU's intermediate buffers of ONE layer, N tokens, with the oracle's op-level functions."""
import ctypes as C, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributedllm_b200 import ggjt, capi
from oracle import oracle

shape = sys.argv[1] if len(sys.argv) > 1 else "tiny"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sh = ggjt.SHAPES[shape]
d = tempfile.mkdtemp()
p = os.path.join(d, "m.bin")
ggjt.write_synth_slice(p, sh, 0, 0, ggjt.T_Q4_0, 0)
L = oracle.port_lib()
f = ggjt.read_file(p, True)
E, FF, H = sh.n_embd, sh.n_ff, sh.n_head
def raw(n): return np.frombuffer(f.read_raw("layers.0." + n), dtype=np.uint8)
def mm(wname, x, rows, K):
    w = raw(wname); nb = K // 32
    out = np.zeros((x.shape[0], rows), np.float32)
    for n in range(x.shape[0]):
        aq = np.zeros(K, np.int8); ad = np.zeros(nb, np.uint16)
        xr = np.ascontiguousarray(x[n])
        L.orc_quant_q8_0(xr.ctypes.data, K, aq.ctypes.data, ad.ctypes.data)
        for r in range(rows):
            out[n, r] = L.orc_dot_q4_0_q8_0(w[r*nb*18:].ctypes.data, aq.ctypes.data, ad.ctypes.data, K)
    return out
rng = np.random.default_rng(1)
x = rng.standard_normal((N, E), dtype=np.float32)
an = np.frombuffer(f.read_raw("layers.0.attention_norm.weight"), np.float32)
cur = np.zeros_like(x)
for n in range(N):
    L.orc_rmsnorm(np.ascontiguousarray(x[n]).ctypes.data, an.ctypes.data, E, cur[n].ctypes.data)
q = mm("attention.wq.weight", cur, E, E); k = mm("attention.wk.weight", cur, E, E); v = mm("attention.wv.weight", cur, E, E)
g = capi.Slice(p, 0, 64)
y = g.forward(x)
def rd(which, count, dt=np.float32):
    out = np.zeros(count, np.uint32)
    capi.lib().b200_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]
    capi.check(capi.lib().b200_debug_read(g.handle, which, 0, count, out.ctypes.data))
    return out.view(dt)
qkv = rd(0, N*3*E).reshape(N, 3*E)
def cmp(name, a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    bad = int((a.view(np.uint32) != b.view(np.uint32)).sum())
    print("%-8s mismatches %d / %d   max|diff| %.3e  ref[:4] %s gpu[:4] %s" % (name, bad, a.size, float(np.abs(a-b).max()), a.ravel()[:4], b.ravel()[:4]))
cmp("q(pre)", q, qkv[:, :E]); cmp("k(pre)", k, qkv[:, E:2*E]); cmp("v", v, qkv[:, 2*E:])
port = oracle.PortSlice(p, 64)
yo = port.forward(x)
cmp("layer", yo, y)

# ---- rope / cache / attention / ffn stages (n_past = 0)
D = E // H
qr = q.copy(); kr = k.copy()
for n in range(N):
    L.orc_rope(qr[n].ctypes.data, H, D, n); L.orc_rope(kr[n].ctypes.data, H, D, n)
q16 = rd(6, N*E//2, np.uint16)[:N*E].reshape(N, E)
kc = rd(7, N*E//2, np.uint16)[:N*E].reshape(N, E)
vc = rd(8, N*E//2, np.uint16)[:N*E].reshape(N, E)
def cmp16(name, a, b):
    bad = int((a != b).sum()); print("%-8s mismatches %d / %d" % (name, bad, a.size), a.ravel()[:4], b.ravel()[:4])
cmp16("q16", qr.astype(np.float16).view(np.uint16), q16)
cmp16("kcache", kr.astype(np.float16).view(np.uint16), kc)
cmp16("vcache", v.astype(np.float16).view(np.uint16), vc)
# attention via oracle primitives
T = N
kq_scale = np.float32(1.0) / np.sqrt(np.float32(E) / np.float32(H)).astype(np.float32)
att = np.zeros((N, E), np.float32)
kch = kr.astype(np.float16).view(np.uint16); vch = v.astype(np.float16).view(np.uint16); qh = qr.astype(np.float16).view(np.uint16)
for n in range(N):
    for h in range(H):
        sc = np.full(T, -np.inf, np.float32)
        for t in range(n + 1):
            sc[t] = np.float32(L.orc_dot_f16(kch[t, h*D:].ctypes.data, 1, qh[n, h*D:].ctypes.data, 1, D)) * kq_scale
        L.orc_softmax_row(sc.ctypes.data, T)
        ph = sc.astype(np.float16).view(np.uint16)
        for c in range(D):
            att[n, h*D + c] = L.orc_dot_f16(vch[:, h*D + c:].ctypes.data, E, ph.ctypes.data, 1, T)
cmp("att", att, rd(1, N*E).reshape(N, E))
wo = mm("attention.wo.weight", att, E, E)
ffin = (wo + x).astype(np.float32)
cmp("ffin", ffin, rd(2, N*E).reshape(N, E))
fn = np.frombuffer(f.read_raw("layers.0.ffn_norm.weight"), np.float32)
cur2 = np.zeros_like(x)
for n in range(N):
    L.orc_rmsnorm(np.ascontiguousarray(ffin[n]).ctypes.data, fn.ctypes.data, E, cur2[n].ctypes.data)
g1 = mm("feed_forward.w1.weight", cur2, FF, E); g3 = mm("feed_forward.w3.weight", cur2, FF, E)
gate = np.array([[np.float32(L.orc_silu(float(a))) for a in row] for row in g1], np.float32) * g3
cmp("gate", gate.astype(np.float32), rd(3, N*FF).reshape(N, FF))
gg = rd(3, N*FF).reshape(N, FF)
badrows = np.nonzero((gate.astype(np.float32).view(np.uint32) != gg.view(np.uint32)).any(axis=0))[0]
print("bad rows:", badrows[:40], "... count", len(badrows))
r0 = int(badrows[0])
print("row", r0, "ref g1", g1[:, r0], "g3", g3[:, r0], "ref gate", gate[:, r0], "gpu", gg[:, r0])
# is gpu == silu(g3)*g1 or g1*g3 or something?
sil = lambda a: np.float32(L.orc_silu(float(a)))
print("alt silu(g3)*g1", [sil(g3[n, r0]) * g1[n, r0] for n in range(N)])
ts = np.zeros(65536, np.uint16); te = np.zeros(65536, np.uint16); L.orc_tables(te.ctypes.data, ts.ctypes.data)
