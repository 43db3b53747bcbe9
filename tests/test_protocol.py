"""Wire compatibility of distributedllm_b200.protocol with the reference's frames (tests/golden/protocol.json,
produced by distllm/protocol.py itself) -- the contract pinned by the reference's tests/unit/test_protocol.py:90-132."""
import base64
import json
import os

import numpy as np
import pytest

from distributedllm_b200 import protocol

FRAMES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "protocol.json")))


class ChunkySocket:
    """recv() hands out at most `chunk` bytes, like the reference's VaryingChunkSocketMock."""

    def __init__(self, data, chunk):
        self.data, self.pos, self.chunk = data, 0, chunk

    def recv(self, n):
        out = self.data[self.pos:self.pos + min(n, self.chunk)]
        self.pos += len(out)
        return out


def _body(f):
    return {k: (base64.b64decode(v) if k in f["bytes_fields"] else v) for k, v in f["body"].items()}


@pytest.mark.parametrize("f", FRAMES, ids=[f["cls"] for f in FRAMES])
def test_encode_is_byte_identical_to_the_reference(f):
    assert getattr(protocol, f["cls"])(**_body(f)).encode() == base64.b64decode(f["frame_b64"])


@pytest.mark.parametrize("chunk", [1, 7, 1 << 20])
@pytest.mark.parametrize("f", FRAMES, ids=[f["cls"] for f in FRAMES])
def test_decode_reference_frames(f, chunk):
    name, body = protocol.receive_message(ChunkySocket(base64.b64decode(f["frame_b64"]), chunk))
    msg = protocol.restore_message(name, body)
    assert type(msg).__name__ == f["cls"]
    want = _body(f)
    for k, v in want.items():
        got = getattr(msg, k)
        if isinstance(v, list):
            assert got == [float(np.float32(x)) for x in v]
        else:
            assert got == v


def test_numpy_values_encode_like_lists():
    a = protocol.RequestPropagateForward(1, 3, np.array([1.5, -2.0, 3.25], np.float32)).encode()
    b = protocol.RequestPropagateForward(1, 3, [1.5, -2.0, 3.25]).encode()
    assert a == b


def test_corrupted_frame_is_rejected():
    frame = bytearray(protocol.RequestLoadSlice(name="orb").encode())
    frame[-1] ^= 1
    with pytest.raises(Exception, match="integrity"):
        protocol.receive_message(ChunkySocket(bytes(frame), 64))


def test_unknown_message_and_long_name():
    with pytest.raises(Exception, match="Unrecognized"):
        protocol.restore_message("no_such_message", {})
    with pytest.raises(protocol.TooLongMessageStringError):
        protocol.encode_message("x" * 31, {})


def test_message_equality_and_body():
    m = protocol.ResponseWithError(operation="o", error="e", description="")
    assert m.get_message() == "operation_failure" and m.get_body() == {"operation": "o", "error": "e", "description": ""}
    assert m == protocol.ResponseWithError("o", "e", "") and m != protocol.ResponseWithError("o", "x", "")
