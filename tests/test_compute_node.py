"""Compute-node handlers, uploads and the RPC client against an in-process TCP server -- the reference's
tests/unit/test_compute_node.py / test_control_center.py scenarios, with the DummySlice (k*x + b) standing in
for the model (slices.py:19-26, 64-71 of the reference)."""
import hashlib
import io
import json
import threading

import pytest

from distributedllm_b200 import protocol
from distributedllm_b200.compute_node import serve
from distributedllm_b200.compute_node.routes import routes
from distributedllm_b200.compute_node.tcp_handler import RequestContext
from distributedllm_b200.control_center import Connection, OperationFailedError


def _upload(ctx, data: bytes, meta: dict, checksum=None) -> int:
    sid = routes["request_file_submission_begin"](ctx, protocol.RequestFileSubmissionBegin(json.dumps(meta))).submission_id
    for i in range(0, len(data), 3):
        r = routes["request_submit_part"](ctx, protocol.RequestSubmitPart(sid, i // 3, data[i:i + 3]))
        assert r.part_size == len(data[i:i + 3])
    return sid, routes["request_file_submission_end"](ctx, protocol.RequestFileSubmissionEnd(
        sid, checksum or hashlib.sha256(data).hexdigest()))


def test_status_brand_new_then_up():
    ctx = RequestContext.default(names=["a", "b"])
    assert json.loads(routes["status_request"](ctx, protocol.RequestStatus()).status_json) == {"status": "brand_new"}
    meta = {"type": "slice", "model": "m", "layer_from": 0, "layer_to": 3, "format": "test"}
    _, end = _upload(ctx, bytes([2, 5]), meta)
    assert (end.file_name, end.total_size) == ("a", 2)
    loaded = routes["load_slice_request"](ctx, protocol.RequestLoadSlice(name="a"))
    assert (loaded.name, loaded.model) == ("a", "m")
    st = json.loads(routes["status_request"](ctx, protocol.RequestStatus()).status_json)
    assert st["status"] == "up" and st["metadata"] == meta


def test_propagate_forward_dummy_slice_and_errors():
    ctx = RequestContext.default(names=["a"])
    req = protocol.RequestPropagateForward(2, 2, [1.0, 2.0, 3.0, 4.0])
    r = routes["propagate_forward_request"](ctx, req)
    assert (r.msg, r.error, r.operation) == ("operation_failure", "slice_not_loaded", "propagate_forward_request")
    _upload(ctx, bytes([3, 1]), {"type": "slice", "model": "m", "layer_from": 0, "layer_to": 0, "format": "test"})
    routes["load_slice_request"](ctx, protocol.RequestLoadSlice(name="a"))
    r = routes["propagate_forward_request"](ctx, req)
    assert (r.msg, r.axis0, r.axis1, r.values) == ("tensor_response", 2, 2, [4.0, 7.0, 10.0, 13.0])
    assert routes["clear_context_request"](ctx, protocol.RequestClearContext()).msg == "clear_context_response"
    bad = RequestContext.with_failing_loader(names=["a"])
    assert routes["propagate_forward_request"](bad, req).error == "neural_computation_error"


def test_load_slice_failures():
    ctx = RequestContext.with_failing_loader(names=["a"])
    assert routes["load_slice_request"](ctx, protocol.RequestLoadSlice(name="nope")).error == "slice_not_found"
    _upload(ctx, b"xy", {"type": "slice", "model": "m", "layer_from": 0, "layer_to": 0, "format": "test"})
    assert routes["load_slice_request"](ctx, protocol.RequestLoadSlice(name="a")).error == "slice_load_error"


def test_upload_state_machine():
    ctx = RequestContext.default(names=["a", "b"])
    begin = protocol.RequestFileSubmissionBegin(json.dumps({"type": "any"}))
    sid = routes["request_file_submission_begin"](ctx, begin).submission_id
    assert routes["request_file_submission_begin"](ctx, begin).error == "parallel_upload_forbidden"
    assert routes["request_submit_part"](ctx, protocol.RequestSubmitPart(sid + 5, 0, b"z")).error == "upload_not_found"
    assert routes["request_file_submission_end"](ctx, protocol.RequestFileSubmissionEnd(sid + 5, "0")).error == "upload_not_found"
    routes["request_submit_part"](ctx, protocol.RequestSubmitPart(sid, 0, b"hello"))
    assert routes["request_file_submission_end"](ctx, protocol.RequestFileSubmissionEnd(sid, "deadbeef")).error == "file_upload_failed"
    assert ctx.registry.failed == [sid] and ctx.registry.in_progress == []
    # non-slice uploads are not listed as slices
    _upload(ctx, b"abc", {"type": "any"})
    assert json.loads(routes["slices_request"](ctx, protocol.RequestAllSlices()).slices_json) == []


@pytest.fixture()
def node(tmp_path):
    import distributedllm_b200.compute_node.tcp_handler as th
    th._PROD = None
    srv = serve.make_server("127.0.0.1", 0, str(tmp_path / "uploads"))
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    yield Connection(("127.0.0.1", srv.server_address[1])), srv
    srv.shutdown()
    srv.server_close()
    th._PROD = None


def test_client_against_live_node(node, tmp_path):
    conn, _ = node
    assert conn.get_status() == {"status": "brand_new"}
    with pytest.raises(OperationFailedError):
        conn.propagate_forward([1.0], (1, 1))
    res = conn.push_slice(io.BytesIO(bytes([2, 1])), "toy", {"layer_from": 4, "layer_to": 7, "format": "test"}, chunk_size=1)
    assert res == {"file_name": "orb", "total_size": 2}
    assert conn.list_all_slices() == [{"name": "orb", "model": "toy", "layer_from": 4, "layer_to": 7}]
    with pytest.raises(OperationFailedError):
        conn.load_slice("missing")
    assert conn.load_slice("orb") == {"name": "orb", "model": "toy"}
    out = conn.propagate_forward([0.5, -1.0, 2.0], (1, 3))
    assert out == {"shape": [1, 3], "values": [2.0, -1.0, 5.0]}
    assert conn.clear_context() == {}
    # registry persisted for the next start
    state = json.load(open(tmp_path / "uploads" / "registry_data.json"))
    assert state["finished"] == [0]


def test_shape_mismatch_is_an_error(node, monkeypatch):
    conn, _ = node
    conn.push_slice(io.BytesIO(bytes([1, 0])), "toy", {"layer_from": 0, "layer_to": 0, "format": "test"})
    conn.load_slice("orb")
    real = conn._get_response

    def lie(request, sock=None):
        r = real(request, sock)
        if r.msg == "tensor_response":
            r.axis1 += 1
        return r
    monkeypatch.setattr(conn, "_get_response", lie)
    with pytest.raises(OperationFailedError):
        conn.propagate_forward([1.0, 2.0], (1, 2))


# ---------------------------------------------------------------- additive wire format
def _toy_node(tmp_path, name, k, b, loaded=True):
    """A node with its own context (several per process) holding DummySlice(k, b)."""
    ctx = RequestContext.default(str(tmp_path / name), names=[name])
    srv = serve.make_server("127.0.0.1", 0, str(tmp_path / name), context=ctx)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    conn = Connection(("127.0.0.1", srv.server_address[1]))
    if loaded:
        conn.push_slice(io.BytesIO(bytes([k, b])), "toy", {"layer_from": 0, "layer_to": 0, "format": "test"})
        conn.load_slice(name)
    return srv, conn


def test_binary_tensor_messages_round_trip():
    import numpy as np
    x = np.arange(6, dtype=np.float32) / 4
    m = protocol.RequestPropagateBytes(2, 3, x.tobytes(), json.dumps(["10.0.0.2:9090"]))
    name, body = protocol.decode_frame(m.encode()[4:])
    assert name == "propagate_bytes_request" and protocol.restore_message(name, body) == m
    # the field codecs are the reference's: a list-typed tensor of the same floats carries the same payload bytes
    as_list = protocol.RequestPropagateForward(2, 3, x.tolist()).encode()
    assert x.tobytes() in as_list and x.tobytes() in m.encode()
    r = protocol.ResponsePropagateBytes(2, 3, x.tobytes())
    name, body = protocol.decode_frame(r.encode()[4:])
    assert protocol.restore_message(name, body) == r


def test_binary_propagate_and_node_to_node_chain(tmp_path):
    import numpy as np
    nodes = [_toy_node(tmp_path, "n%d" % i, k, b) for i, (k, b) in enumerate([(2, 1), (3, 0), (1, 5)])]
    try:
        conns = [c for _, c in nodes]
        x = np.array([0.5, -1.0, 2.0, 8.0], dtype=np.float32)
        # star topology, list wire (the reference's) vs bytes wire: same numbers
        want = x.tolist()
        for c in conns:
            want = c.propagate_forward(want, (1, 4))["values"]
        got = x
        for c in conns:
            got = c.propagate_forward_bytes(got, (1, 4))
        assert got.dtype == np.float32 and got.tolist() == want == [((2 * v + 1) * 3) + 5 for v in x.tolist()]
        # chained: ONE request to the first node, which forwards along the route
        route = ["127.0.0.1:%d" % s.server_address[1] for s, _ in nodes[1:]]
        chained = conns[0].propagate_forward_bytes(x, (1, 4), route)
        assert chained.tolist() == want
        # client wrapper
        from distributedllm_b200.client import DistributedLLM
        llm = DistributedLLM.__new__(DistributedLLM)
        llm.addresses = [("127.0.0.1", s.server_address[1]) for s, _ in nodes]
        for wire in ("list", "bytes", "chain"):
            llm.wire = wire
            assert llm.propagate_tensor(x.tolist()) == want, wire
    finally:
        for s, _ in nodes:
            s.shutdown()
            s.server_close()


def test_chain_failures_come_back_to_the_client(tmp_path):
    import numpy as np
    a = _toy_node(tmp_path, "a", 1, 1)
    b = _toy_node(tmp_path, "b", 1, 1, loaded=False)        # no slice loaded on the second hop
    try:
        x = np.ones(3, dtype=np.float32)
        with pytest.raises(OperationFailedError, match="slice_not_loaded"):
            a[1].propagate_forward_bytes(x, (1, 3), ["127.0.0.1:%d" % b[0].server_address[1]])
        with pytest.raises(OperationFailedError, match="chain_hop_failed"):
            a[1].propagate_forward_bytes(x, (1, 3), ["127.0.0.1:1"])          # nothing listens there
        # a tensor that is not a whole number of float32, and an unparsable hop
        r = a[1]._get_response(protocol.RequestPropagateBytes(1, 3, b"\x00" * 5, "[]"))
        assert r.get_message() == "operation_failure" and r.error == "neural_computation_error"
        with pytest.raises(OperationFailedError, match="chain_hop_failed"):
            a[1].propagate_forward_bytes(x, (1, 3), ["not-a-hop"])
    finally:
        for s, _ in (a, b):
            s.shutdown()
            s.server_close()


# ---- round 2: load metadata, restart of the production singleton, route validation -------------------
def test_load_options_come_from_the_slice_metadata():
    from distributedllm_b200.compute_node.slices import load_options
    assert load_options({"type": "slice", "model": "m", "layer_from": 0, "layer_to": 3}) == {}
    assert load_options({"n_ctx": 2048, "device": 3, "n_sessions": 8, "model": "m"}) == {"n_ctx": 2048, "device": 3, "n_sessions": 8}
    assert load_options({"b200": {"n_ctx": "1024"}, "n_ctx": 512}) == {"n_ctx": 1024}          # the sub-object wins
    assert load_options({"device": 0}) == {"device": 0}
    for bad in ({"n_ctx": 0}, {"n_sessions": -1}, {"device": -2}):
        with pytest.raises(ValueError):
            load_options(bad)


def test_slice_container_passes_load_options_to_the_llm_module(monkeypatch):
    """routes.py hands the upload's metadata to SliceContainer.load; n_ctx / device / n_sessions reach llm.load_slice as
    keyword extras (no process environment involved), and two loads may differ."""
    import types

    from distributedllm_b200.compute_node import slices
    calls = []
    fake = types.SimpleNamespace(load_slice=lambda path, **kw: calls.append((path, kw)) or 0, propagate_forward=None)
    monkeypatch.setattr(slices, "import_llm", lambda: fake)
    c = slices.SliceContainer(None)
    c.load("/x/a.bin", {"type": "slice", "model": "m", "n_ctx": 2048, "n_sessions": 8})
    c.load("/x/b.bin", {"type": "slice", "model": "m", "b200": {"n_ctx": 512, "device": 1}})
    c.load("/x/c.bin", {"type": "slice", "model": "m"})
    assert calls == [("/x/a.bin", {"n_ctx": 2048, "n_sessions": 8}), ("/x/b.bin", {"n_ctx": 512, "device": 1}), ("/x/c.bin", {})]


def test_production_context_survives_a_restart_with_a_respelled_uploads_dir(tmp_path, monkeypatch):
    """the singleton was keyed on registry.root, which restore_registry overwrites with the root stored in
    registry_data.json -- a node restarted with './uploads' instead of 'uploads' lost every uploaded slice."""
    import distributedllm_b200.compute_node.tcp_handler as th
    monkeypatch.chdir(tmp_path)
    th._PROD = None
    ctx = RequestContext.production("uploads", ["a", "b"])
    serve.restore_registry(ctx, "uploads")
    _upload(ctx, bytes([2, 5]), {"type": "slice", "model": "m", "layer_from": 0, "layer_to": 3, "format": "test"})
    assert ctx.registry.finished == [0]
    # restart: fresh process state, the directory spelled differently
    th._PROD = None
    spelled = "./uploads"
    ctx2 = RequestContext.production(spelled, ["a", "b"])
    serve.restore_registry(ctx2, spelled)
    assert RequestContext.production(spelled, ["a", "b"]) is ctx2            # the first request must NOT rebuild it
    assert ctx2.registry.finished == [0] and ctx2.registry.next_id == 1
    assert ctx2.registry.root == spelled
    assert [s["name"] for s in json.loads(routes["slices_request"](ctx2, protocol.RequestAllSlices()).slices_json)] == ["a"]
    sid, _ = _upload(ctx2, bytes([1, 1]), {"type": "slice", "model": "m", "layer_from": 4, "layer_to": 7, "format": "test"})
    assert sid == 1                                                           # does not overwrite upload_0
    th._PROD = None


def test_route_is_validated_before_the_forward_runs():
    from distributedllm_b200.compute_node import routes as R
    ctx = RequestContext.default(names=["a"])
    ctx.self_addresses = ("127.0.0.1:9000",)
    ran = []
    ctx.slice_container.forward = lambda t: ran.append(1) or t
    data = b"\0" * 8

    def ask(route):
        return routes["propagate_bytes_request"](ctx, protocol.RequestPropagateBytes(1, 2, data, json.dumps(route)))
    assert ask(["nonsense"]).error == "chain_hop_failed"
    assert ask(["host:99999"]).error == "chain_hop_failed"
    assert ask(["127.0.0.1:9000"]).error == "chain_hop_failed"                # loops back to this node
    assert ask(["10.0.0.1:1"] * (R.MAX_ROUTE_HOPS + 1)).error == "chain_hop_failed"
    assert ask({"not": "a list"}).error == "chain_hop_failed"
    assert ran == []                                                          # nothing was computed for a bad route
    ctx.peers = ("10.0.0.2:7000",)
    assert ask(["10.0.0.3:7000"]).error == "chain_hop_failed"                 # not in the peer list
    assert ran == []
    assert ask([]).msg == "tensor_bytes_response" and ran == [1]
