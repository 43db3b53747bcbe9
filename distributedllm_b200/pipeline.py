"""Layer-slice pipeline across ranks (one process per GPU): who holds which layers, and the hand-off schedule.

The reference's `nodes_map` gives every node a contiguous layer range and the CLIENT relays the activation from
node to node (README.md:116-131, cli_api/common.py:148-154).  On one NVSwitch box the relay is replaced by a direct
hop rank r -> r+1; this module holds the rank-side logic shared by the two transports:

  * `b200_pipeline_step` (C ABI, csrc/runtime.cu): ncclSend / ncclRecv on the slice's CUDA stream -- production;
  * `TorchDistTransport`: torch.distributed send/recv of host tensors -- any backend; the world_size-2 `gloo`
    tests drive the same schedule on CPU with it.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def layer_ranges(n_layer: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [first, last] layer ranges, sizes differing by at most one (earlier ranks take the extra)."""
    if world < 1 or world > n_layer:
        raise ValueError("cannot cut %d layers into %d slices" % (n_layer, world))
    base, extra = divmod(n_layer, world)
    out, a = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append((a, a + k - 1))
        a += k
    return out


def ranges_from_nodes_map(nodes_map: dict) -> List[Tuple[str, Tuple[int, int]]]:
    """nodes_map {"ip:port": [a, b]} -> [(address, (a, b))] in pipeline order; validates contiguity."""
    items = sorted(((addr, (int(ab[0]), int(ab[1]))) for addr, ab in nodes_map.items()), key=lambda t: t[1])
    nxt = items[0][1][0] if items else 0
    for addr, (a, b) in items:
        if a != nxt or b < a:
            raise ValueError("nodes_map layer ranges are not contiguous at %s: [%d, %d]" % (addr, a, b))
        nxt = b + 1
    return items


class TorchDistTransport:
    def __init__(self, dist):
        self.dist = dist
        import torch
        self.torch = torch

    def send(self, x: np.ndarray, dst: int) -> None:
        self.dist.send(self.torch.from_numpy(np.ascontiguousarray(x, np.float32)), dst)

    def recv(self, shape, src: int) -> np.ndarray:
        t = self.torch.empty(tuple(shape), dtype=self.torch.float32)
        self.dist.recv(t, src)
        return t.numpy()


class PipelineStage:
    """One rank of the pipeline.  `forward` is the rank's slice ([n, n_embd] f32 -> same shape)."""

    def __init__(self, forward: Callable[[np.ndarray], np.ndarray], rank: int, world: int, n_embd: int, transport):
        self.forward, self.rank, self.world, self.n_embd, self.tr = forward, rank, world, n_embd, transport

    def step(self, x: Optional[np.ndarray], n_tokens: int, ring: bool = True) -> Optional[np.ndarray]:
        """Rank 0 feeds `x`; every other rank receives from its predecessor.  With `ring`, the last rank's output
        returns to rank 0 (where the client-side lm_head lives) and is this call's result there."""
        r, w = self.rank, self.world
        if r > 0:
            x = self.tr.recv((n_tokens, self.n_embd), r - 1)
        elif x is None:
            raise ValueError("rank 0 needs the input activation")
        y = self.forward(np.ascontiguousarray(x, np.float32).reshape(n_tokens, self.n_embd))
        if r < w - 1:
            self.tr.send(y, r + 1)
        elif ring and w > 1:
            self.tr.send(y, 0)
        if r == 0 and w > 1:
            return self.tr.recv((n_tokens, self.n_embd), w - 1) if ring else None
        return y if r == w - 1 else None


def join_pipeline(sl, rank: int, world: int, broadcast_bytes: Callable[[Optional[bytes], int], bytes],
                  all_gather_bytes: Callable[[bytes], Sequence[bytes]], peer: bool = True) -> str:
    """Bring one rank's slice (`capi.Slice`) into the layer-slice pipeline of `world` ranks and pick the hand-off
    transport.  The host only moves a few bytes of set-up data through the two collectives it is given
    (`torch.distributed` in bench.py / the tests; anything else works):

      1. rank 0 draws the ncclUniqueId, everyone joins the communicator (`b200_pipeline_init`) -- the single
         ncclSend/ncclRecv hop stays available as the tested fallback;
      2. every rank exports the cudaIpc handle of its mailbox, the handles are all-gathered, every rank maps its ring
         neighbours (`b200_pipeline_mailbox_connect`): from then on the hop is a peer-memory store + flag inside the
         step's CUDA graph;
      3. the ranks agree: if ANY rank could not map a neighbour (no P2P / IPC in this container), all stay on NCCL.
    Returns "peer" or "nccl"."""
    from . import capi
    lib = capi.lib()
    raw = np.zeros(128, np.uint8)
    if rank == 0:
        capi.check(lib.b200_pipeline_unique_id(raw.ctypes.data))
    raw = np.frombuffer(broadcast_bytes(raw.tobytes() if rank == 0 else None, 128), np.uint8).copy()
    capi.check(lib.b200_pipeline_init(sl.handle, rank, world, raw.ctypes.data))
    if not peer or world < 2:
        return "nccl"
    h = np.zeros(64, np.uint8)
    ok = lib.b200_pipeline_mailbox_export(sl.handle, h.ctypes.data) == 0
    handles = list(all_gather_bytes(h.tobytes() + bytes([1 if ok else 0])))
    ok = all(x[64] == 1 for x in handles)
    if ok:
        buf = np.frombuffer(b"".join(x[:64] for x in handles), np.uint8).copy()
        ok = lib.b200_pipeline_mailbox_connect(sl.handle, buf.ctypes.data, world) == 0
        ok = ok and lib.b200_pipeline_transport(sl.handle) == 1
    agreed = all(x == b"\x01" for x in all_gather_bytes(b"\x01" if ok else b"\x00"))
    capi.check(lib.b200_pipeline_set_transport(sl.handle, 1 if agreed else 0))
    return "peer" if agreed else "nccl"


def torch_collectives(dist, device):
    """(broadcast_bytes, all_gather_bytes) over an initialised torch.distributed group (any backend)."""
    import torch

    def broadcast_bytes(data: Optional[bytes], n: int) -> bytes:
        t = torch.zeros(n, dtype=torch.uint8, device=device)
        if data is not None:
            t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def all_gather_bytes(data: bytes):
        mine = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
        outs = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, mine)
        return [bytes(o.cpu().numpy().tobytes()) for o in outs]

    return broadcast_bytes, all_gather_bytes
