"""ctypes front-ends for the two CPU checkers (TEST INFRASTRUCTURE, see oracle/__init__.py).

* `PortSlice`  -- oracle/liboracle.so, the plain-C restatement (slice_oracle.c).
* `RefSlice`   -- oracle/_ref/libllmref.so, the UNMODIFIED reference translation unit
                  (tensor_processor.cpp) behind ref_shim.cpp's C door.  Only present when
                  oracle/_ref was built (in the build container; the binaries travel).
Both take the reference's own slice files and mirror `llm.load_slice / propagate_forward /
clear_context` (tensor_processor.cpp:1995-2030, 2127-2163) on numpy buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from distributedllm_b200 import ggjt

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libllmref.so")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "port"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


def have_port() -> bool:
    return os.path.isfile(PORT_SO)


def have_ref() -> bool:
    return os.path.isfile(REF_SO)


_port = None


def port_lib() -> C.CDLL:
    global _port
    if _port is None:
        if not have_port():
            build(ref=False)
        L = C.CDLL(PORT_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int] * 6
        L.orc_set_layer.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.orc_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_forward.restype = C.c_int
        for fn in ("orc_clear", "orc_free"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.orc_n_past.argtypes = [C.c_void_p]
        L.orc_set_n_past.argtypes = [C.c_void_p, C.c_int]
        L.orc_dot_q4_0_q8_0.restype = C.c_float
        L.orc_dot_q4_0_q8_0.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dot_q8_0_q8_0.restype = C.c_float
        L.orc_dot_q8_0_q8_0.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dot_q4_1_q8_1.restype = C.c_float
        L.orc_dot_q4_1_q8_1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_quant_q8_1.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_dot_f16.restype = C.c_float
        L.orc_dot_f16.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int]
        L.orc_quant_q8_0.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_rope.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_softmax_row.argtypes = [C.c_void_p, C.c_int]
        L.orc_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_fp32_to_fp16.restype = C.c_uint16
        L.orc_fp32_to_fp16.argtypes = [C.c_float]
        L.orc_silu.restype = C.c_float
        L.orc_silu.argtypes = [C.c_float]
        _port = L
    return _port


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class PortSlice:
    """The C restatement on one slice file."""

    def __init__(self, path: str, n_ctx: int = 512):
        self.lib = port_lib()
        self.file = ggjt.read_file(path, sliced=True)
        hp = self.file.hparams
        self.n_embd, self.n_layer, self.first_layer = hp.n_embd, hp.n_layer, hp.first_layer
        self.mm = np.memmap(path, dtype=np.uint8, mode="r")
        wt = self.file.tensors["layers.%d.attention.wq.weight" % hp.first_layer].ttype
        self.h = self.lib.orc_create(hp.n_embd, hp.n_head, hp.n_ff, hp.n_layer, n_ctx, wt)
        self._keep = []
        for i in range(hp.n_layer):
            pre = "layers.%d." % (i + hp.first_layer)
            ptrs = []
            for nm in ("attention_norm.weight", "attention.wq.weight", "attention.wk.weight", "attention.wv.weight",
                       "attention.wo.weight", "ffn_norm.weight", "feed_forward.w1.weight", "feed_forward.w2.weight",
                       "feed_forward.w3.weight"):
                t = self.file.tensors[pre + nm]
                a = np.array(self.mm[t.offset:t.offset + t.nbytes])          # private, aligned copy
                self._keep.append(a)
                ptrs.append(_ptr(a))
            self.lib.orc_set_layer(self.h, i, *ptrs)

    def forward(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.n_embd)
        out = np.empty_like(x)
        rc = self.lib.orc_forward(self.h, _ptr(x), x.shape[0], _ptr(out))
        if rc != 0:
            raise RuntimeError("oracle forward failed: %d" % rc)
        return out

    @property
    def n_past(self) -> int:
        return self.lib.orc_n_past(self.h)

    def set_n_past(self, p: int) -> None:
        self.lib.orc_set_n_past(self.h, p)

    def clear_context(self) -> None:
        self.lib.orc_clear(self.h)

    def close(self) -> None:
        if self.h:
            self.lib.orc_free(self.h)
            self.h = None


_ref = None


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        L.ref_slice_load.restype = C.c_void_p
        L.ref_slice_load.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.ref_slice_n_embd.argtypes = [C.c_void_p]
        L.ref_slice_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_slice_clear.argtypes = [C.c_void_p]
        L.ref_slice_free.argtypes = [C.c_void_p]
        L.ref_embed.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_logits.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_next_token.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.ref_tokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
        _ref = L
    return _ref


class RefSlice:
    """The compiled reference (TransformerSlice, tensor_processor.cpp:1488-1562)."""

    MAX_CHUNK = 32     # the reference's eval arena overflows for long calls

    def __init__(self, path: str, n_threads: int = 3, n_ctx: int = 512):
        self.lib = ref_lib()
        self.h = self.lib.ref_slice_load(path.encode(), n_threads, n_ctx)
        self.n_embd = self.lib.ref_slice_n_embd(self.h)

    def forward(self, x: np.ndarray) -> np.ndarray:
        """One reference call (N = rows of x), as `llm.propagate_forward` would make it."""
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.n_embd)
        out = np.empty_like(x)
        rc = self.lib.ref_slice_forward(self.h, _ptr(x), x.shape[0], _ptr(out))
        if rc != 0:
            raise RuntimeError("reference forward failed: %d" % rc)
        return out

    def clear_context(self) -> None:
        self.lib.ref_slice_clear(self.h)

    def close(self) -> None:
        if self.h:
            self.lib.ref_slice_free(self.h)
            self.h = None


def ref_embed(extra_path: str, tokens, n_embd: int) -> np.ndarray:
    t = np.ascontiguousarray(tokens, dtype=np.int32)
    out = np.empty((len(t), n_embd), dtype=np.float32)
    ref_lib().ref_embed(extra_path.encode(), _ptr(t), len(t), _ptr(out), 3)
    return out


def ref_logits(extra_path: str, emb: np.ndarray, n_vocab: int, all_logits: bool) -> np.ndarray:
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    n = emb.shape[0] if all_logits else 1
    out = np.empty((n, n_vocab), dtype=np.float32)
    ref_lib().ref_logits(extra_path.encode(), _ptr(emb), emb.size, int(all_logits), _ptr(out))
    return out


def ref_tokenize(extra_path: str, prompt: str) -> list:
    buf = np.empty(4096, dtype=np.int32)
    n = ref_lib().ref_tokenize(extra_path.encode(), prompt.encode(), _ptr(buf), 4096)
    return buf[:n].tolist()
