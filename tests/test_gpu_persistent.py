"""GPU: the persistent single-token step (csrc/persist.cuh, B200_PERSIST=1) -- one kernel for all layers of the slice,
weights streamed by TMA through every dependency of the layer, phases ordered by grid-wide counters -- must be
bit-identical to the multi-kernel step, i.e. to the oracle / the compiled reference."""
import os

import numpy as np
import pytest

from distributedllm_b200 import ggjt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "libllmref.so"))
THREADS = min(16, os.cpu_count() or 4)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("wtype", [ggjt.T_Q4_0, ggjt.T_Q8_0], ids=["q4_0", "q8_0"])
@pytest.mark.parametrize("ctas", [0, 8])
@pytest.mark.parametrize("graph", [1, 0])
def test_persistent_step_is_bit_identical(tmp_models, monkeypatch, wtype, ctas, graph):
    from distributedllm_b200 import capi
    from oracle import oracle
    monkeypatch.setenv("B200_PERSIST", "1")
    monkeypatch.setenv("B200_PERSIST_CTAS", str(ctas))     # 8 CTAs = 32 groups for 48 qkv tiles: several rounds of tiles per group (the wrap-around path)
    monkeypatch.setenv("B200_GRAPH", str(graph))
    sh = ggjt.SHAPES["tiny128"]
    path = tmp_models("tiny128", wtype, 0, 2, seed=7)
    gpu, cpu = capi.Slice(path, 0, 96), oracle.PortSlice(path, 96)
    rng = np.random.default_rng(13)
    x = rng.standard_normal((37, sh.n_embd), dtype=np.float32)
    assert (_bits(gpu.forward(x)) == _bits(cpu.forward(x))).all()          # prompt: the multi-token path
    for step in range(40):                                                   # T crosses 64 (two 32-slot bodies + tails)
        x = rng.standard_normal((1, sh.n_embd), dtype=np.float32)
        g, c = gpu.forward(x), cpu.forward(x)
        assert (_bits(g) == _bits(c)).all(), "step %d: %d floats differ" % (step, int((_bits(g) != _bits(c)).sum()))
    gpu.clear_context()
    cpu.clear_context()
    for step in range(3):                                                    # from an empty context: T = 1, 2, 3
        x = rng.standard_normal((1, sh.n_embd), dtype=np.float32)
        assert (_bits(gpu.forward(x)) == _bits(cpu.forward(x))).all(), step
    gpu.close()
    cpu.close()


def test_persistent_step_full_size_layers(tmp_path, monkeypatch):
    """7B (E 4096, 32 heads) and 13B (E 5120, 40 heads: five warps of RMSNorm partials) layer shapes, device-resident
    decode steps deep in the context."""
    from distributedllm_b200 import capi
    from oracle import oracle
    monkeypatch.setenv("B200_PERSIST", "1")
    for name, layers, seed in (("7b", 2, 21), ("13b", 1, 22)):
        sh = ggjt.SHAPES[name]
        p = str(tmp_path / ("%s.bin" % name))
        ggjt.write_fast_q4_slice(p, sh, 0, layers - 1, seed=seed)
        gpu = capi.Slice(p, 0, 512)
        ref = oracle.RefSlice(p, THREADS, 512) if HAVE_REF else oracle.PortSlice(p, 512)
        rng = np.random.default_rng(seed)
        pos = 0
        while pos < 290:
            n = min(32, 290 - pos)
            x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
            assert (_bits(gpu.forward(x)) == _bits(ref.forward(x))).all()
            pos += n
        for step in range(6):
            x = rng.standard_normal((1, sh.n_embd), dtype=np.float32)
            g, r = gpu.forward(x), ref.forward(x)
            assert (_bits(g) == _bits(r)).all(), "%s step %d: %d floats differ" % (name, step, int((_bits(g) != _bits(r)).sum()))
        gpu.close()
        ref.close()


def test_persistent_step_sessions(tmp_models, monkeypatch):
    from distributedllm_b200 import capi
    monkeypatch.setenv("B200_PERSIST", "1")
    sh = ggjt.SHAPES["tiny128"]
    path = tmp_models("tiny128", ggjt.T_Q4_0, 0, 1, seed=9)
    gpu = capi.Slice(path, 0, 64, n_sessions=3)
    monkeypatch.setenv("B200_PERSIST", "0")
    priv = [capi.Slice(path, 0, 64) for _ in range(3)]
    rng = np.random.default_rng(2)
    for k, n in enumerate((3, 1, 6)):
        x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
        assert (_bits(gpu.session_forward(k, x)) == _bits(priv[k].forward(x))).all()
    for step in range(5):
        for k in (2, 0, 1):
            x = rng.standard_normal((1, sh.n_embd), dtype=np.float32)
            assert (_bits(gpu.session_forward(k, x)) == _bits(priv[k].forward(x))).all(), (step, k)
    for s in priv + [gpu]:
        s.close()
