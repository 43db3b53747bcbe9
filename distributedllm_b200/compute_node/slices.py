"""Slice container of a compute node (reference: distllm/compute_node/slices.py:5-95).

`SliceContainer.load / forward / clear_context` keep the reference's semantics:
  * metadata {'format': 'test'} selects the two-byte `DummySlice` (k, b) -> k*x + b used by handler tests
    (slices.py:19-26, 64-71);
  * anything else is a GGJT slice file run on the GPU by `GGMLSlice` through the `llm` module
    (slices.py:74-91): llm.load_slice(path); llm.propagate_forward(values); llm.clear_context() != 0 raises.
A non-list result of llm.propagate_forward (the module returns an int status when the eval fails,
tensor_processor.cpp:2148-2151) raises NeuralComputationError, which the route maps to
`neural_computation_error`.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import Any, Optional

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def import_llm():
    """The `llm` extension is built in-tree next to libb200slice.so; make it importable by its bare name,
    the way the reference's Dockerfile does with PYTHONPATH=/libs."""
    if _PKG_DIR not in sys.path:
        sys.path.insert(0, _PKG_DIR)
    import llm
    if not hasattr(llm, "propagate_forward"):
        raise ImportError("the imported `llm` module is not the slice runtime")
    return llm


@dataclass
class Tensor:
    shape: tuple
    values: Any            # list[float] on the reference wire; numpy float32 also accepted


class SliceNotLoadedError(Exception):
    pass


class NeuralComputationError(Exception):
    pass


class ModelSlice:
    def __call__(self, tensor: Tensor) -> Tensor:
        raise NotImplementedError

    def clear_context(self) -> None:
        pass


class DummySlice(ModelSlice):
    def __init__(self, k, b):
        self.k, self.b = k, b

    def __call__(self, tensor: Tensor) -> Tensor:
        if hasattr(tensor.values, "dtype"):
            import numpy as np
            return Tensor(tensor.shape, (np.float64(self.k) * tensor.values.astype(np.float64) + self.b).astype(np.float32))
        return Tensor(tensor.shape, [self.k * v + self.b for v in tensor.values])


LOAD_OPTION_KEYS = ("n_ctx", "device", "n_sessions")


def load_options(metadata) -> dict:
    """Per-slice load options carried by the slice's upload metadata: context length, GPU ordinal and
    session count -- the knobs the reference hard-codes at tensor_processor.cpp:1997-2006.  Accepted at the top level
    of the metadata JSON or under a "b200" sub-object; anything else in the metadata is ignored here."""
    opts = {}
    for src in (metadata or {}, (metadata or {}).get("b200") or {}):
        for k in LOAD_OPTION_KEYS:
            v = src.get(k) if isinstance(src, dict) else None
            if v is None:
                continue
            v = int(v)
            if v < 0 or (k != "device" and v == 0):
                raise ValueError("slice metadata: %s must be positive (got %r)" % (k, v))
            opts[k] = v
    return opts


class GGMLSlice(ModelSlice):
    """A reference-format slice file resident on the GPU."""

    def __init__(self, file_path: str, **options):
        self.path = file_path
        self.options = {k: int(v) for k, v in options.items() if k in LOAD_OPTION_KEYS and v is not None}
        self.llm = import_llm()
        self.llm.load_slice(self.path, **self.options)     # keyword extras are additive; llm.load_slice(path) still works

    def __call__(self, tensor: Tensor) -> Tensor:
        if hasattr(tensor.values, "dtype"):
            # binary wire format: float32 buffer in, bytes out, no per-float Python objects (llm.propagate_forward_buffer)
            import numpy as np
            try:
                out = self.llm.propagate_forward_buffer(np.ascontiguousarray(tensor.values, dtype=np.float32))
            except RuntimeError as e:            # the buffer entry point raises where the list one returns a status
                raise NeuralComputationError(str(e))
            return Tensor(tensor.shape, np.frombuffer(out, dtype=np.float32))
        out = self.llm.propagate_forward(tensor.values)
        if not isinstance(out, list):
            raise NeuralComputationError("slice forward failed with status %r" % (out,))
        return Tensor(tensor.shape, out)

    def clear_context(self) -> None:
        if self.llm.clear_context() != 0:
            raise Exception("Error occurred when clearing context")


class SliceContainer:
    def __init__(self, fs_backend):
        self.fs_backend = fs_backend
        self.slice: Optional[ModelSlice] = None
        self.metadata = None

    def load(self, slice_path, metadata) -> None:
        self.metadata = metadata
        if metadata.get("format") == "test":
            with self.fs_backend.open_file(slice_path, mode="rb") as f:
                data = f.read()
            self.slice = DummySlice(data[0], data[1])
        else:
            self.slice = GGMLSlice(slice_path, **load_options(metadata))

    def forward(self, tensor: Tensor) -> Tensor:
        if self.slice is None:
            raise SliceNotLoadedError()
        return self.slice(tensor)

    def clear_context(self) -> None:
        if self.slice is not None:
            self.slice.clear_context()

    @property
    def info(self):
        return self.metadata

    @property
    def is_loaded(self) -> bool:
        return self.slice is not None
