"""The N>1 path on CPU: two ranks over `gloo`, each holding one layer slice (the C oracle stands in for the GPU
slice), hand the activation rank 0 -> rank 1 -> back to rank 0.  The result must be bit-identical to the
un-sliced model, for a prefill and for decode steps."""
import os
import socket
import sys

import numpy as np
import pytest

from distributedllm_b200 import ggjt
from distributedllm_b200.pipeline import layer_ranges, ranges_from_nodes_map


def test_layer_ranges():
    assert layer_ranges(32, 1) == [(0, 31)]
    assert layer_ranges(32, 4) == [(0, 7), (8, 15), (16, 23), (24, 31)]
    assert layer_ranges(26, 4) == [(0, 6), (7, 13), (14, 19), (20, 25)]
    assert layer_ranges(40, 8) == [(5 * i, 5 * i + 4) for i in range(8)]
    with pytest.raises(ValueError):
        layer_ranges(2, 3)


def test_nodes_map_order_and_contiguity():
    nm = {"10.0.0.2:9090": [17, 25], "10.0.0.1:9090": [0, 16]}
    assert ranges_from_nodes_map(nm) == [("10.0.0.1:9090", (0, 16)), ("10.0.0.2:9090", (17, 25))]
    with pytest.raises(ValueError):
        ranges_from_nodes_map({"a:1": [0, 3], "b:1": [5, 7]})


def _worker(rank, world, port, tmpdir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from distributedllm_b200.pipeline import PipelineStage, TorchDistTransport, layer_ranges
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = ggjt.SHAPES["tiny"]
    a, b = layer_ranges(sh.n_layer, world)[rank]
    sl = oracle.PortSlice(os.path.join(tmpdir, "slice_%d_%d.bin" % (a, b)), 64)
    stage = PipelineStage(sl.forward, rank, world, sh.n_embd, TorchDistTransport(dist))
    rng = np.random.default_rng(11)
    outs = []
    for n in (9, 1, 1, 3):
        x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
        y = stage.step(x if rank == 0 else None, n, ring=True)
        if rank == 0:
            outs.append(y)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put([o.tobytes() for o in outs])


def test_two_rank_pipeline_matches_single_slice(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle
    sh = ggjt.SHAPES["tiny"]
    for a, b in layer_ranges(sh.n_layer, 2) + [(0, sh.n_layer - 1)]:
        ggjt.write_synth_slice(str(tmp_path / ("slice_%d_%d.bin" % (a, b))), sh, a, b, ggjt.T_Q4_0, seed=0)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = oracle.PortSlice(str(tmp_path / ("slice_0_%d.bin" % (sh.n_layer - 1))), 64)
    rng = np.random.default_rng(11)
    for raw, n in zip(got, (9, 1, 1, 3)):
        x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
        assert whole.forward(x).tobytes() == raw


def test_join_pipeline_falls_back_to_nccl_unless_every_rank_maps_its_neighbours(monkeypatch):
    """Host logic of the transport choice (pipeline.join_pipeline) with a fake C ABI, one thread per rank over in-memory
    collectives: the peer-memory mailboxes are used only if EVERY rank exported its handle, connected, and reports the peer
    transport; one failing rank sends all of them to NCCL."""
    import threading
    import types

    from distributedllm_b200 import capi, pipeline

    def run(world, failing_rank, stage):
        tls = threading.local()
        calls = {r: [] for r in range(world)}

        def fake_lib():
            rank = tls.rank
            L = types.SimpleNamespace()
            L.b200_pipeline_unique_id = lambda p: 0
            L.b200_pipeline_init = lambda h, r, w, p: calls[rank].append("init") or 0
            L.b200_pipeline_mailbox_export = lambda h, p: (1 if (rank == failing_rank and stage == "export") else 0)
            L.b200_pipeline_mailbox_connect = lambda h, p, w: (4 if (rank == failing_rank and stage == "connect") else 0)
            L.b200_pipeline_transport = lambda h: 1
            L.b200_pipeline_set_transport = lambda h, peer: calls[rank].append(("set", peer)) or 0
            L.b200_last_error = lambda: b""
            return L
        monkeypatch.setattr(capi, "lib", fake_lib)
        barrier = threading.Barrier(world)
        slots = [None] * world
        root = {}
        results = {}

        def worker(rank):
            tls.rank = rank

            def bcast(data, n):
                if rank == 0:
                    root["v"] = data
                barrier.wait()
                v = root["v"]
                barrier.wait()
                return v

            def gather(data):
                slots[rank] = data
                barrier.wait()
                out = list(slots)
                barrier.wait()
                return out
            results[rank] = pipeline.join_pipeline(types.SimpleNamespace(handle=rank), rank, world, bcast, gather)
        ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(30)
        return results, calls

    res, calls = run(4, failing_rank=-1, stage="")
    assert set(res.values()) == {"peer"} and all(c[-1] == ("set", 1) for c in calls.values())
    for stage in ("export", "connect"):
        res, calls = run(4, failing_rank=2, stage=stage)
        assert set(res.values()) == {"nccl"}, (stage, res)
        assert all(c[-1] == ("set", 0) for c in calls.values())
