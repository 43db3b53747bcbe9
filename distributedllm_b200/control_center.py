"""RPC client of a compute node (reference: distllm/control_center.py:88-261).

`Connection` keeps the reference's methods and failure behaviour: a fresh TCP connection per request
(control_center.py:119-226), `OperationFailedError` on `operation_failure` replies, on a size mismatch after an
upload and on an echoed-shape mismatch after propagate_forward (232-244); chunk retries x3 (167-188)."""
from __future__ import annotations

import hashlib
import json
import socket as _socket
from typing import Optional

from . import protocol


class OperationFailedError(Exception):
    pass


class NodeProvisioningError(Exception):
    pass


def connect(address):
    sock = _socket.socket(_socket.AF_INET, _socket.SOCK_STREAM)
    sock.setsockopt(_socket.IPPROTO_TCP, _socket.TCP_NODELAY, 1)
    sock.connect(address)
    return sock


def disconnect(sock):
    sock.close()


class Connection:
    def __init__(self, address):
        self.address = address
        self.connect = connect
        self.disconnect = disconnect

    # ---- one request / one reply
    def _get_response(self, request, sock=None):
        sock = sock or self.connect(self.address)
        try:
            request.send(sock)
            name, body = protocol.receive_message(sock)
        finally:
            try:
                self.disconnect(sock)
            except Exception:
                pass
        return protocol.restore_message(name, body)

    # ---- uploads
    def push_slice(self, f, model, metadata=None, chunk_size=1024 * 1024, file_size=None, progress_bar=False):
        meta = {"type": "slice", "model": model}
        meta.update(metadata or {})
        return self.push_file(f, meta, chunk_size, file_size, progress_bar)

    def push_file(self, f, metadata=None, chunk_size=1024 * 1024, file_size=None, progress_bar=False):
        reply = self._get_response(protocol.RequestFileSubmissionBegin(json.dumps(metadata)))
        if reply.msg == "operation_failure":
            raise OperationFailedError
        sid = reply.submission_id
        bar = None
        if progress_bar:
            try:
                from tqdm import tqdm
                bar = tqdm(total=file_size, desc="Uploading slice", unit="bytes", unit_scale=True)
            except ImportError:
                bar = None
        hasher, total, part = hashlib.sha256(), 0, 0
        while True:
            chunk = f.read(chunk_size)
            if not chunk:
                break
            hasher.update(chunk)
            total += len(chunk)
            self._send_chunk(chunk, sid, part)
            part += 1
            if bar:
                bar.update(len(chunk))
        if bar:
            bar.close()
        reply = self._get_response(protocol.RequestFileSubmissionEnd(sid, hasher.hexdigest()))
        if reply.msg == "operation_failure":
            raise OperationFailedError
        if reply.msg != "file_submission_end_response":
            raise OperationFailedError("Unexpected message code in response: %s" % reply.msg)
        if reply.total_size != total:
            raise OperationFailedError
        return reply.get_body()

    def _send_chunk(self, data, submission_id, part, max_retries=3):
        error = ""
        for _ in range(max_retries):
            reply = self._get_response(protocol.RequestSubmitPart(submission_id, part, data))
            if reply.msg == "submit_part_response":
                if reply.part_size == len(data):
                    return
            elif reply.msg == "operation_failure":
                error = {"integrity_error": "Part of file got corrupted during transfer",
                         "upload_not_found": "Upload not found on the side of the server"}.get(reply.error, error)
            else:
                raise OperationFailedError("Unexpected message code in response: %s" % reply.msg)
        raise OperationFailedError(error)

    # ---- control
    def list_all_slices(self):
        return json.loads(self._get_response(protocol.RequestAllSlices()).slices_json)

    def load_slice(self, name):
        reply = self._get_response(protocol.RequestLoadSlice(name=name))
        if reply.get_message() == "operation_failure":
            raise OperationFailedError("")
        return reply.get_body()

    def clear_context(self):
        reply = self._get_response(protocol.RequestClearContext())
        if reply.get_message() == "operation_failure":
            raise OperationFailedError("")
        return reply.get_body()

    def get_status(self):
        return json.loads(self._get_response(protocol.RequestStatus()).status_json)

    # ---- the hot path
    def propagate_forward(self, tensor, shape):
        axis0, axis1 = shape
        reply = self._get_response(protocol.RequestPropagateForward(axis0, axis1, tensor))
        kind = reply.get_message()
        if kind == "operation_failure":
            raise OperationFailedError
        if kind != "tensor_response":
            raise Exception("Cannot handle unrecognized message")
        if (reply.axis0, reply.axis1) != (shape[0], shape[1]):
            raise OperationFailedError
        return {"shape": [reply.axis0, reply.axis1], "values": reply.values}


    def propagate_forward_bytes(self, tensor, shape, route=()):
        """Binary wire format (additive): `tensor` is a float32 numpy array / buffer; `route` lists the "host:port" hops
        the node should forward to on its own.  Returns a float32 numpy array (the LAST hop's output)."""
        import json
        import numpy as np
        axis0, axis1 = shape
        data = np.ascontiguousarray(tensor, dtype=np.float32).tobytes()
        reply = self._get_response(protocol.RequestPropagateBytes(axis0, axis1, data, json.dumps(list(route))))
        kind = reply.get_message()
        if kind == "operation_failure":
            raise OperationFailedError("%s: %s" % (reply.error, reply.description))
        if kind != "tensor_bytes_response":
            raise Exception("Cannot handle unrecognized message")
        if (reply.axis0, reply.axis1) != (shape[0], shape[1]):
            raise OperationFailedError
        return np.frombuffer(reply.data, dtype=np.float32)


class ControlCenter:
    """Status book-keeping over a nodes map {name: (ip, port)} (reference: control_center.py:8-71)."""

    def __init__(self, nodes_map):
        self.nodes_map = nodes_map
        self.status = {"ready": False, "model": None,
                       "nodes": {name: {"connectivity": True, "ip": ip, "port": port, "status": "brand_new",
                                        "errors": [], "slice": None} for name, (ip, port) in nodes_map.items()}}

    def push_model(self, model_name, slices, meta_data=None):
        if set(slices) != set(self.nodes_map):
            raise NodeProvisioningError
        nodes = {}
        for name, node in self.status["nodes"].items():
            nodes[name] = dict(node, status="up", slice=[slices[name].layer_from, slices[name].layer_to])
        self.status = {"ready": True, "model": {"pseudoname": model_name, "family": None, "class": None, "size": None},
                       "nodes": nodes}

    def get_status(self):
        return self.status
