"""Client-side orchestration of a layer-sliced model (reference: distllm/cli_api/common.py:9-154).

`DistributedLLM.generate / perplexity / propagate_tensor / clear_context` and `Sampler` keep the reference's
behaviour (including the shape (1, len) it sends and T=0 -> logits / 1e-5 "greedy" with repeat penalty).  The
extra layers (tokenizer, embeddings, lm_head) run through the `llm` module, which here keeps the extra-layers
file resident on the GPU instead of re-reading it on every call.

`LocalPipeline` is the on-box fast path: all slices of the nodes_map live on GPUs of this machine and the
activation never leaves HBM between them (one process, one slice handle per GPU)."""
from __future__ import annotations

import json
from typing import List, Sequence, Tuple

import numpy as np

from .compute_node.slices import import_llm
from .control_center import Connection


def parse_address(address: str) -> Tuple[str, int]:
    host, port = address.split(":")
    return host, int(port)


def load_one_slice(model_id, address_str, a, b) -> bool:
    conn = Connection(address=parse_address(address_str))
    status = conn.get_status()
    if status["status"] == "up":
        meta = status["metadata"]
        if (meta["model"], meta["layer_from"], meta["layer_to"]) == (model_id, a, b):
            return True
    for s in conn.list_all_slices():
        if (s["model"], s["layer_from"], s["layer_to"]) == (model_id, a, b):
            conn.load_slice(s["name"])
            return True
    return False


def get_llm(config_path, registry_path="models_registry/registry.json"):
    with open(config_path) as f:
        config = json.load(f)
    items = list(config["nodes_map"].items())
    for address_str, (a, b) in items:
        load_one_slice(config["model_id"], address_str, a, b)
    nodes = [parse_address(addr) for addr, _ in sorted(items, key=lambda t: t[1])]
    with open(registry_path) as f:
        registry = json.load(f)
    return DistributedLLM(nodes, registry[config["model_id"]]["extra_layers_file"])


def _softmax(x: np.ndarray, axis=-1) -> np.ndarray:
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


class Sampler:
    def __init__(self, temperature=0.7, repeat_penalty=1.1, rng=None):
        self.T = temperature
        self.penalty = repeat_penalty
        self.previous_ids: List[int] = []
        self.eps = 10 ** (-5)
        self.rng = rng or np.random

    def __call__(self, logits) -> int:
        logits = np.array(logits)
        ids = np.arange(len(logits))
        seen = np.isin(ids, self.previous_ids)
        logits = logits / ((seen * self.penalty + ~seen) * (self.T + self.eps))
        token_id = int(self.rng.choice(ids, p=_softmax(logits)))
        self.previous_ids.append(token_id)
        return token_id


class DistributedLLM:
    def __init__(self, addresses: Sequence[Tuple[str, int]], extra_layers_path: str, wire: str = "list"):
        """wire: "list"  -- the reference's per-node star of float lists (common.py:148-154);
                 "bytes" -- same star, tensors as one binary field;
                 "chain" -- binary, and the nodes pass the activation along themselves: one client round trip per step."""
        if wire not in ("list", "bytes", "chain"):
            raise ValueError("wire must be list, bytes or chain")
        self.addresses = list(addresses)
        self.extra_layers_path = extra_layers_path
        self.wire = wire
        self.llm = import_llm()

    def generate(self, prompt, max_steps=200, temperature=0.0, repeat_penalty=1.1):
        self.clear_context()
        extra = self.extra_layers_path
        tokens = self.llm.tokenize_prompt(extra, prompt)
        sampler = Sampler(temperature, repeat_penalty)
        for _ in range(max_steps):
            emb = self.propagate_tensor(self.llm.prepare_embeddings(extra, tokens))
            token_id = sampler(self.llm.get_logits(extra, emb, False))
            tokens = [token_id]
            yield self.llm.decode_token(extra, token_id)

    def generate_greedy(self, prompt, max_steps=200) -> List[int]:
        """Pure argmax decoding through llm.get_next_token (tensor_processor.cpp:1894-1908): the parity path."""
        self.clear_context()
        extra = self.extra_layers_path
        tokens = self.llm.tokenize_prompt(extra, prompt)
        out = []
        for _ in range(max_steps):
            emb = self.propagate_tensor(self.llm.prepare_embeddings(extra, tokens))
            token_id = self.llm.get_next_token(extra, emb)
            out.append(token_id)
            tokens = [token_id]
        return out

    def perplexity(self, text) -> float:
        self.clear_context()
        extra = self.extra_layers_path
        tokens = self.llm.tokenize_prompt(extra, text)
        emb = self.propagate_tensor(self.llm.prepare_embeddings(extra, tokens[:-1]))
        n = len(tokens) - 1
        logits = np.array(self.llm.get_logits(extra, emb, True)).reshape(n, -1)
        p = _softmax(logits, axis=1)[np.arange(n), tokens[1:]]
        return float(np.exp(-np.log(p).sum() / n))

    def clear_context(self):
        for address in self.addresses:
            Connection(address).clear_context()

    def propagate_tensor(self, embeddings):
        shape = (1, len(embeddings))
        if self.wire == "list":
            for address in self.addresses:
                embeddings = Connection(address).propagate_forward(embeddings, shape)["values"]
            return embeddings
        x = np.asarray(embeddings, dtype=np.float32)
        if self.wire == "bytes":
            for address in self.addresses:
                x = Connection(address).propagate_forward_bytes(x, shape)
        else:
            route = ["%s:%d" % (h, p) for h, p in self.addresses[1:]]
            x = Connection(self.addresses[0]).propagate_forward_bytes(x, shape, route)
        return x.tolist()


class LocalPipeline:
    """All slices of a nodes_map on the GPUs of THIS box: slice i on device i, activations chained device to
    device (peer copy) -- the single-process equivalent of the NCCL pipeline bench.py runs with one rank per GPU."""

    def __init__(self, slice_paths: Sequence[str], devices: Sequence[int] = None, n_ctx: int = 0):
        from . import capi
        self.capi = capi
        devices = list(devices) if devices is not None else list(range(len(slice_paths)))
        self.slices = [capi.Slice(p, d, n_ctx) for p, d in zip(slice_paths, devices)]
        self.slices.sort(key=lambda s: s.info.first_layer)

    def propagate_tensor(self, embeddings) -> np.ndarray:
        x = np.ascontiguousarray(embeddings, dtype=np.float32)
        for s in self.slices:
            x = s.forward(x)
        return x

    def clear_context(self):
        for s in self.slices:
            s.clear_context()

    def close(self):
        for s in self.slices:
            s.close()
