"""Request handlers of the compute node, one per wire message (reference: distllm/compute_node/routes.py).

Same request names, same replies, same error codes:
  status_request                -> status_response {'status': 'up'|'brand_new', 'metadata'?}
  slices_request                -> slices_list_response [{name, model, layer_from, layer_to}]
  load_slice_request            -> loaded_slice_response | operation_failure(slice_load_error | slice_not_found)
  request_file_submission_begin -> file_submission_begin_response | operation_failure(parallel_upload_forbidden)
  request_submit_part           -> submit_part_response | operation_failure(upload_not_found)
  request_file_submission_end   -> file_submission_end_response | operation_failure(upload_not_found | file_upload_failed)
  propagate_forward_request     -> tensor_response | operation_failure(neural_computation_error | slice_not_loaded)
  clear_context_request         -> clear_context_response | operation_failure(clear_context_failure)
Additive:
  propagate_bytes_request       -> tensor_bytes_response | operation_failure(... | chain_hop_failed): binary tensor, and
                                   node-to-node chaining along `route` so the client makes ONE round trip per step
"""
from __future__ import annotations

import json
from typing import Callable, Dict

from .. import protocol
from .slices import NeuralComputationError, SliceNotLoadedError, Tensor
from .uploads import FailedUploadError, ParallelUploadError, UploadNotFoundError

routes: Dict[str, Callable] = {}


def route(name: str):
    def deco(fn):
        routes[name] = fn
        return fn
    return deco


def _failure(message, error: str, description: str = ""):
    return protocol.ResponseWithError(operation=message.get_message(), error=error, description=description)


def _read_json(ctx, path):
    with ctx.manager.fs_backend.open_file(path, "r") as f:
        return json.loads(f.read())


def list_slices(ctx):
    out = []
    for sid in ctx.registry.finished:
        meta = _read_json(ctx, ctx.registry.get_location(sid).metadata_path)
        if meta.get("type") != "slice":
            continue
        out.append(dict(name=ctx.name_gen.id_to_name(sid), model=meta["model"],
                        layer_from=meta["layer_from"], layer_to=meta["layer_to"]))
    return out


@route("status_request")
def handle_status(ctx, message):
    c = ctx.slice_container
    status = {"status": "up" if c.is_loaded else "brand_new"}
    if c.is_loaded:
        status["metadata"] = c.metadata
    return protocol.JsonResponseWithStatus(json.dumps(status))


@route("slices_request")
def handle_slices(ctx, message):
    return protocol.JsonResponseWithSlices(json.dumps(list_slices(ctx)))


@route("load_slice_request")
def handle_load_slice(ctx, message):
    for entry in list_slices(ctx):
        if entry["name"] != message.name:
            continue
        loc = ctx.registry.get_location(ctx.name_gen.name_to_id(message.name))
        try:
            ctx.slice_container.load(loc.upload_path, _read_json(ctx, loc.metadata_path))
        except Exception:
            return _failure(message, "slice_load_error")
        return protocol.JsonResponseWithLoadedSlice(name=message.name, model=entry["model"])
    return _failure(message, "slice_not_found")


@route("request_file_submission_begin")
def handle_submission_begin(ctx, message):
    try:
        return protocol.ResponseFileSubmissionBegin(ctx.manager.prepare_upload(json.loads(message.metadata_json)))
    except ParallelUploadError:
        return _failure(message, "parallel_upload_forbidden")


@route("request_submit_part")
def handle_submit_part(ctx, message):
    try:
        return protocol.ResponseSubmitPart(ctx.manager.upload_part(message.submission_id, message.data))
    except UploadNotFoundError:
        return _failure(message, "upload_not_found")


@route("request_file_submission_end")
def handle_submission_end(ctx, message):
    try:
        total = ctx.manager.finilize_upload(message.submission_id, message.checksum)
    except FileNotFoundError:
        return _failure(message, "upload_not_found")
    except FailedUploadError:
        return _failure(message, "file_upload_failed")
    name = ctx.name_gen.id_to_name(message.submission_id)
    if not name:
        return _failure(message, "file_upload_failed")
    return protocol.ResponseFileSubmissionEnd(name, total)


@route("propagate_forward_request")
def handle_propagate_forward(ctx, message):
    tensor = Tensor((message.axis0, message.axis1), message.values)
    try:
        out = ctx.slice_container.forward(tensor)
    except NeuralComputationError:
        return _failure(message, "neural_computation_error")
    except SliceNotLoadedError:
        return _failure(message, "slice_not_loaded")
    axis0, axis1 = out.shape
    return protocol.ResponsePropagateForward(axis0, axis1, out.values)


MAX_ROUTE_HOPS = 64


def parse_hop(hop: str):
    if not isinstance(hop, str):
        raise ValueError("hop must be 'host:port'")
    host, _, port = hop.rpartition(":")
    if not host or not port.isdigit() or not 0 < int(port) < 65536:
        raise ValueError("hop %r is not 'host:port'" % (hop,))
    return host, int(port)


def check_route(ctx, hops):
    """A route is client-supplied: bound its length, validate every hop's format BEFORE running the forward, refuse
    loops back to this node, and -- when the node was started with a peer list -- hops outside it."""
    if not isinstance(hops, list) or len(hops) > MAX_ROUTE_HOPS:
        raise ValueError("route must be a list of at most %d hops" % MAX_ROUTE_HOPS)
    peers, me = set(getattr(ctx, "peers", ()) or ()), set(getattr(ctx, "self_addresses", ()) or ())
    for hop in hops:
        parse_hop(hop)
        if hop in me:
            raise ValueError("route loops back to this node (%s)" % hop)
        if peers and hop not in peers:
            raise ValueError("hop %s is not in this node's peer list" % hop)


@route("propagate_bytes_request")
def handle_propagate_bytes(ctx, message):
    import numpy as np
    if len(message.data) % 4:
        return _failure(message, "neural_computation_error", "tensor is not a whole number of float32")
    try:
        hops = json.loads(message.route) if message.route else []
        check_route(ctx, hops)
    except ValueError as e:
        return _failure(message, "chain_hop_failed", "bad route: %s" % e)
    tensor = Tensor((message.axis0, message.axis1), np.frombuffer(message.data, dtype=np.float32))
    try:
        out = ctx.slice_container.forward(tensor)
    except NeuralComputationError:
        return _failure(message, "neural_computation_error")
    except SliceNotLoadedError:
        return _failure(message, "slice_not_loaded")
    data = np.ascontiguousarray(out.values, dtype=np.float32).tobytes()
    if not hops:
        return protocol.ResponsePropagateBytes(out.shape[0], out.shape[1], data)
    # hand the activation to the next node ourselves and relay whatever comes back (a tensor or a failure)
    try:
        import socket
        with socket.create_connection(parse_hop(hops[0]), timeout=ctx_timeout(ctx)) as sock:
            protocol.RequestPropagateBytes(out.shape[0], out.shape[1], data, json.dumps(hops[1:])).send(sock)
            name, body = protocol.receive_message(sock)
        return protocol.restore_message(name, body)
    except Exception as e:      # unreachable / misbehaving next hop
        return _failure(message, "chain_hop_failed", "%s: %r" % (hops[0], e))


def ctx_timeout(ctx):
    return getattr(ctx, "hop_timeout", 120.0)


@route("clear_context_request")
def handle_clear_context(ctx, message):
    try:
        ctx.slice_container.clear_context()
        return protocol.ResponseClearContext()
    except Exception as e:
        return _failure(message, "clear_context_failure", repr(e))
