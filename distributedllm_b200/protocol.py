"""Wire protocol of the compute-node RPC -- byte-compatible with the reference's `distllm/protocol.py`.

Frame (distllm/protocol.py:185-228, utils.py:26-140, all integers native-endian 32-bit):

    i32 total_len | 64 B hex sha256(payload) | payload
    payload = 30 B message name (space padded) | i32 n_fields | n_fields x (str name, str type, value)
    str = i32 len + utf-8 bytes; bytes = i32 len + raw; int = i32; float = f32; list = i32 count + count x f32;
    None = the string "None"

The message set (names, field order, field types) is the reference's (protocol.py:46-166); it is declared in
ONE table below and the classes are generated from it.  Float lists are packed with `array('f')`, which emits
exactly the bytes of the reference's per-element `struct.pack('f')` at memcpy speed, and numpy float32 arrays
are accepted wherever a list is.
"""
from __future__ import annotations

import hashlib
import struct
from array import array
from dataclasses import make_dataclass
from typing import Any, Dict, Tuple

MAX_MESSAGE_SIZE = 30            # bytes reserved for the message name
_DIGEST_LEN = 64
_I32 = struct.Struct("i")
_F32 = struct.Struct("f")


class ByteCodingError(Exception):
    pass


class TooLongMessageStringError(Exception):
    pass


# ------------------------------------------------------------------------------------------ field codecs
def _enc_int(v: int) -> bytes:
    try:
        return _I32.pack(v)
    except struct.error as e:
        raise ByteCodingError(str(e))


def _enc_str(s: str) -> bytes:
    raw = s.encode("utf-8")
    return _enc_int(len(raw)) + raw


def _enc_floats(values) -> bytes:
    if hasattr(values, "dtype"):                      # numpy array
        import numpy as np
        raw = np.ascontiguousarray(values, dtype=np.float32).tobytes()
        return _enc_int(len(raw) // 4) + raw
    try:
        raw = array("f", values).tobytes()
    except (TypeError, OverflowError) as e:
        raise ByteCodingError(str(e))
    return _enc_int(len(raw) // 4) + raw


def _encode_field(value: Any) -> Tuple[str, bytes]:
    if isinstance(value, (bytes, bytearray, memoryview)):
        raw = bytes(value)
        return "bytes", _enc_int(len(raw)) + raw
    if isinstance(value, str):
        return "str", _enc_str(value)
    if isinstance(value, bool) or isinstance(value, int):
        return "int", _enc_int(int(value))
    if isinstance(value, float):
        try:
            return "float", _F32.pack(value)
        except (struct.error, OverflowError) as e:
            raise ByteCodingError(str(e))
    if isinstance(value, (list, tuple)) or hasattr(value, "dtype"):
        return "list", _enc_floats(value)
    if value is None:
        return "NoneType", _enc_str("None")
    raise Exception("Unsupport data type: %r" % (value,))


class _Cursor:
    def __init__(self, data: bytes):
        self.data = memoryview(data)
        self.pos = 0

    def take(self, n: int) -> memoryview:
        if n < 0 or self.pos + n > len(self.data):
            raise ByteCodingError("truncated message")
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def int(self) -> int:
        return _I32.unpack(self.take(4))[0]

    def str(self) -> str:
        try:
            return bytes(self.take(self.int())).decode("utf-8")
        except UnicodeDecodeError as e:
            raise ByteCodingError(str(e))

    def floats(self) -> list:
        a = array("f")
        a.frombytes(bytes(self.take(4 * self.int())))
        return a.tolist()

    def field(self, ftype: str) -> Any:
        if ftype == "bytes":
            return bytes(self.take(self.int()))
        if ftype == "str":
            return self.str()
        if ftype == "int":
            return self.int()
        if ftype == "float":
            return _F32.unpack(self.take(4))[0]
        if ftype == "list":
            return self.floats()
        if ftype == "NoneType":
            self.str()
            return None
        raise ByteCodingError("Unknown parameter type: %s" % ftype)


# ------------------------------------------------------------------------------------------ framing
def encode_message(message: str, body: Dict[str, Any]) -> bytes:
    if len(message) > MAX_MESSAGE_SIZE:
        raise TooLongMessageStringError("")
    parts = [message.ljust(MAX_MESSAGE_SIZE).encode("ascii"), _enc_int(len(body))]
    for name, value in body.items():
        ftype, raw = _encode_field(value)
        parts += [_enc_str(name), _enc_str(ftype), raw]
    payload = b"".join(parts)
    digest = hashlib.sha256(payload).hexdigest().encode("ascii")
    return _enc_int(len(digest) + len(payload)) + digest + payload


def decode_frame(data: bytes) -> Tuple[str, Dict[str, Any]]:
    """`data` = everything after the 4-byte length: digest + payload."""
    digest, payload = bytes(data[:_DIGEST_LEN]), data[_DIGEST_LEN:]
    if hashlib.sha256(payload).hexdigest().encode("ascii") != digest:
        raise Exception("Data integrity error. Hashes do not match")
    name = bytes(payload[:MAX_MESSAGE_SIZE]).decode("ascii").strip()
    cur = _Cursor(payload[MAX_MESSAGE_SIZE:])
    body = {}
    for _ in range(cur.int()):
        fname = cur.str()
        body[fname] = cur.field(cur.str())
    return name, body


def recv_exact(sock, n: int) -> bytes:
    chunks, got = [], 0
    while got < n:
        chunk = sock.recv(min(1 << 20, n - got))
        if not chunk:
            raise ConnectionError("socket closed after %d of %d bytes" % (got, n))
        chunks.append(chunk)
        got += len(chunk)
    return b"".join(chunks)


def receive_message(sock) -> Tuple[str, Dict[str, Any]]:
    (size,) = _I32.unpack(recv_exact(sock, 4))
    return decode_frame(recv_exact(sock, size))


def send_message(sock, message: str, body: Dict[str, Any]) -> None:
    sock.sendall(encode_message(message, body))


# ------------------------------------------------------------------------------------------ messages
message_registry: Dict[str, type] = {}


class Message:
    msg = ""

    def get_message(self) -> str:
        return self.msg

    def get_body(self) -> Dict[str, Any]:
        return {k: v for k, v in self.__dict__.items() if k != "msg"}

    def encode(self) -> bytes:
        return encode_message(self.msg, self.get_body())

    def send(self, sock) -> None:
        sock.sendall(self.encode())

    @classmethod
    def from_body(cls, body: Dict[str, Any]):
        return cls(**body)

    def __eq__(self, other: object) -> bool:
        return type(self) is type(other) and self.get_body() == other.get_body()


# class name, wire name, fields in wire order  (reference: protocol.py:46-166)
_SPEC = [
    ("RequestAllSlices", "slices_request", []),
    ("RequestStatus", "status_request", []),
    ("RequestLoadSlice", "load_slice_request", [("name", str)]),
    ("RequestPropagateForward", "propagate_forward_request", [("axis0", int), ("axis1", int), ("values", list)]),
    ("ResponsePropagateForward", "tensor_response", [("axis0", int), ("axis1", int), ("values", list)]),
    ("RequestClearContext", "clear_context_request", []),
    ("ResponseClearContext", "clear_context_response", []),
    ("RequestFileSubmissionBegin", "request_file_submission_begin", [("metadata_json", str)]),
    ("ResponseFileSubmissionBegin", "file_submission_begin_response", [("submission_id", int)]),
    ("RequestSubmitPart", "request_submit_part", [("submission_id", int), ("part_number", int), ("data", bytes)]),
    ("ResponseSubmitPart", "submit_part_response", [("part_size", int)]),
    ("RequestFileSubmissionEnd", "request_file_submission_end", [("submission_id", int), ("checksum", str)]),
    ("ResponseFileSubmissionEnd", "file_submission_end_response", [("file_name", str), ("total_size", int)]),
    ("JsonResponseWithStatus", "status_response", [("status_json", str)]),
    ("JsonResponseWithSlices", "slices_list_response", [("slices_json", str)]),
    ("JsonResponseWithLoadedSlice", "loaded_slice_response", [("name", str), ("model", str)]),
    ("ResponseWithError", "operation_failure", [("operation", str), ("error", str), ("description", str)]),
    ("RequestGreeting", "greeting_request", []),
    ("ResponseGreeting", "greeting_response", []),
    # ---- additive (not in the reference's message set, same framing and field codecs) ----
    # the tensor as ONE `bytes` field (raw native-endian f32, what `list` carries minus the per-float Python objects),
    # plus `route`: a JSON list of "host:port" hops still to visit -- the node forwards its output to route[0] itself
    # and relays the final reply, so the client sends once and receives once instead of once per node
    ("RequestPropagateBytes", "propagate_bytes_request", [("axis0", int), ("axis1", int), ("data", bytes), ("route", str)]),
    ("ResponsePropagateBytes", "tensor_bytes_response", [("axis0", int), ("axis1", int), ("data", bytes)]),
]

for _cls_name, _wire, _fields in _SPEC:
    _cls = make_dataclass(_cls_name, _fields, bases=(Message,), eq=False)
    _cls.msg = _wire
    _cls.__module__ = __name__
    message_registry[_wire] = _cls
    globals()[_cls_name] = _cls


def restore_message(message: str, body: Dict[str, Any]) -> Message:
    cls = message_registry.get(message)
    if cls is None:
        raise Exception("Unrecognized message %s" % message)
    return cls.from_body(body)
