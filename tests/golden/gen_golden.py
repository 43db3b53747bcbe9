"""Generate the committed golden fixtures from the REFERENCE ITSELF (run in the build container only).

  slices.npz     hidden states of oracle/_ref (= distllm/tensor_processor.cpp compiled unmodified, see
                 oracle/Makefile) on seeded synthetic slice files, for a fixed schedule of propagate_forward calls
  slices_q4_1.*  the same for Q4_1 slices (added later; `gen_golden.py q4_1` writes only these) + extra_q4_1.npz
  extra.npz      reference get_inputs / get_llm_output / llama_tokenize on a synthetic extra-layers file
  tokenizer.json reference tokenisation of fixed strings with vendor/llama.cpp/models/ggml-vocab.bin's vocabulary
  protocol.json  frames produced by the reference's distllm/protocol.py for one instance of every message

    python tests/golden/gen_golden.py          # needs /root/reference and a built oracle/_ref
"""
import base64
import hashlib
import json
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from distributedllm_b200 import ggjt  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = [  # name, shape, weight type, layer range, schedule of calls
    ("tiny_q4_0", "tiny", ggjt.T_Q4_0, (1, 2), [40, 1, 1, 7, 1, 20, 3, 1]),
    ("tiny128_q4_0", "tiny128", ggjt.T_Q4_0, (0, 1), [33, 1, 1, 1, 30, 1]),
    ("tiny3b_q4_0", "tiny3b", ggjt.T_Q4_0, (0, 1), [37, 1, 1, 5]),
    ("tiny_q8_0", "tiny", ggjt.T_Q8_0, (0, 1), [18, 1, 1, 16]),
    ("tiny_f16", "tiny", ggjt.T_F16, (2, 3), [35, 1, 1]),
]
CASES_Q4_1 = [
    ("tiny_q4_1", "tiny", ggjt.T_Q4_1, (1, 2), [40, 1, 1, 7, 1, 20, 3, 1]),
    ("tiny128_q4_1", "tiny128", ggjt.T_Q4_1, (0, 1), [33, 1, 1, 1, 30, 1]),
    ("tiny3b_q4_1", "tiny3b", ggjt.T_Q4_1, (0, 1), [37, 1, 1, 5]),
]


def gen_slices(tmp, cases, stem):
    out = {}
    meta = {}
    for name, shape, wt, (a, b), sched in cases:
        sh = ggjt.SHAPES[shape]
        path = os.path.join(tmp, name + ".bin")
        ggjt.write_synth_slice(path, sh, a, b, wt, seed=0)
        meta[name] = {"shape": shape, "wtype": wt, "layers": [a, b], "schedule": sched,
                      "file_sha256": hashlib.sha256(open(path, "rb").read()).hexdigest()}
        rng = np.random.default_rng(1234)
        ref = oracle.RefSlice(path, n_threads=3, n_ctx=512)
        for i, n in enumerate(sched):
            x = rng.standard_normal((n, sh.n_embd), dtype=np.float32)
            out["%s/x%d" % (name, i)] = x
            out["%s/y%d" % (name, i)] = ref.forward(x)
        ref.close()
    np.savez_compressed(os.path.join(HERE, stem + ".npz"), **out)
    json.dump(meta, open(os.path.join(HERE, stem + ".json"), "w"), indent=1)


def gen_q4_1(tmp):
    gen_slices(tmp, CASES_Q4_1, "slices_q4_1")
    # client side of a Q4_1 model: tok_embeddings rows dequantised as nibble * d + m, output.weight through the Q4_1 dot
    sh = ggjt.SHAPES["tiny"]
    extra = os.path.join(tmp, "extra_q4_1.bin")
    ggjt.write_synth_extra(extra, sh, ggjt.T_Q4_1, seed=0)
    toks = np.array([1, 5, 300, 44, 511, 0, 77], np.int32)
    h = np.random.default_rng(7).standard_normal((5, sh.n_embd), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "extra_q4_1.npz"), tokens=toks, emb=oracle.ref_embed(extra, toks, sh.n_embd), hidden=h,
                        logits_all=oracle.ref_logits(extra, h, sh.n_vocab, True),
                        file_sha256=np.frombuffer(hashlib.sha256(open(extra, "rb").read()).digest(), np.uint8))


def main():
    tmp = tempfile.mkdtemp()
    if sys.argv[1:] == ["q4_1"]:
        gen_q4_1(tmp)
        return
    gen_slices(tmp, CASES, "slices")
    gen_q4_1(tmp)

    # extra layers + tokenizer on a synthetic all-Q4_0 extra file
    sh = ggjt.SHAPES["tiny"]
    extra = os.path.join(tmp, "extra.bin")
    ggjt.write_synth_extra(extra, sh, ggjt.T_Q4_0, seed=0)
    toks = np.array([1, 5, 300, 44, 511, 0, 77], np.int32)
    emb = oracle.ref_embed(extra, toks, sh.n_embd)
    h = np.random.default_rng(7).standard_normal((5, sh.n_embd), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "extra.npz"), tokens=toks, emb=emb, hidden=h,
                        logits_all=oracle.ref_logits(extra, h, sh.n_vocab, True),
                        logits_last=oracle.ref_logits(extra, h, sh.n_vocab, False),
                        file_sha256=np.frombuffer(hashlib.sha256(open(extra, "rb").read()).digest(), np.uint8))

    # extra layers as the reference's own `quantize q4_0` writes them for n_embd % 256 == 0: output.weight is Q6_K
    # (llama.cpp:2523-2528).  The quantised file itself is committed (190 KB): k-quant quantisation is not restated here.
    import subprocess
    full = os.path.join(tmp, "full_f32.bin")
    ggjt.write_synth_full(full, sh, ggjt.T_F32, seed=0)
    fq = os.path.join(tmp, "full_q4.bin")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "quantize"), full, fq, "q4_0"], check=True, capture_output=True)
    eq = os.path.join(HERE, "extra_q6k.bin")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "slice_model"), "extra_layers", fq, eq], check=True, capture_output=True)
    assert ggjt.read_file(eq).tensors["output.weight"].ttype == ggjt.T_Q6_K
    h6 = np.random.default_rng(8).standard_normal((6, sh.n_embd), dtype=np.float32)
    h6[3] *= 30.0
    h6[4, :] = 0.0
    np.savez_compressed(os.path.join(HERE, "extra_q6k.npz"), hidden=h6, logits_all=oracle.ref_logits(eq, h6, sh.n_vocab, True),
                        emb=oracle.ref_embed(eq, toks, sh.n_embd), tokens=toks)

    # tokenizer with the real 32000-entry llama vocabulary
    vocab_bin = "/root/reference/vendor/llama.cpp/models/ggml-vocab.bin"
    f = ggjt.read_file(vocab_bin, sliced=False)
    vfile = os.path.join(tmp, "vocab_extra.bin")
    hp = ggjt.HParams(f.hparams.n_vocab, 32, 32, 1, 0, 32, ggjt.FTYPE_F32, ggjt.NO_FIRST_LAYER)
    ggjt.write_file(vfile, hp, f.vocab, [])
    texts = ["Hello World", " Hello World", "Hello, World!", " this is 🦙.cpp", "w048 7tuijk dsdfhu",
             "Alan Turing is", "", "  double  spaces\nnewline\ttab", "ünïcödé ✓ 日本語", "a" * 50]
    json.dump({"vocab_sha256": hashlib.sha256(open(vocab_bin, "rb").read()).hexdigest(),
               "cases": [{"text": t, "ids": oracle.ref_tokenize(vfile, t)} for t in texts]},
              open(os.path.join(HERE, "tokenizer.json"), "w"), indent=1, ensure_ascii=False)
    # the vocabulary itself (432 KB) is needed to replay the cases: keep only (len, bytes, score) records, gzip
    import gzip
    with gzip.open(os.path.join(HERE, "llama_vocab.bin.gz"), "wb") as g:
        for text, score in f.vocab:
            g.write(struct.pack("<I", len(text)) + text + struct.pack("<f", score))

    # protocol frames from the reference implementation
    sys.path.insert(0, "/root/reference")
    from distllm import protocol as ref_protocol
    msgs = [("RequestAllSlices", {}), ("RequestStatus", {}), ("RequestLoadSlice", {"name": "orb"}),
            ("RequestPropagateForward", {"axis0": 1, "axis1": 4, "values": [0.5, -1.25, 3.0, 0.1]}),
            ("ResponsePropagateForward", {"axis0": 1, "axis1": 2, "values": [1e-8, 65504.0]}),
            ("RequestClearContext", {}), ("ResponseClearContext", {}),
            ("RequestFileSubmissionBegin", {"metadata_json": '{"type": "slice", "model": "m"}'}),
            ("ResponseFileSubmissionBegin", {"submission_id": 7}),
            ("RequestSubmitPart", {"submission_id": 7, "part_number": 2, "data": bytes(range(40))}),
            ("ResponseSubmitPart", {"part_size": 40}),
            ("RequestFileSubmissionEnd", {"submission_id": 7, "checksum": "ab" * 32}),
            ("ResponseFileSubmissionEnd", {"file_name": "orb", "total_size": 1 << 20}),
            ("JsonResponseWithStatus", {"status_json": '{"status": "up"}'}),
            ("JsonResponseWithSlices", {"slices_json": "[]"}),
            ("JsonResponseWithLoadedSlice", {"name": "orb", "model": "llama"}),
            ("ResponseWithError", {"operation": "load_slice_request", "error": "slice_not_found", "description": "ü"}),
            ("RequestGreeting", {}), ("ResponseGreeting", {})]
    frames = []
    for cls, body in msgs:
        frame = getattr(ref_protocol, cls)(**body).encode()
        jb = {k: (base64.b64encode(v).decode() if isinstance(v, bytes) else v) for k, v in body.items()}
        frames.append({"cls": cls, "body": jb, "bytes_fields": [k for k, v in body.items() if isinstance(v, bytes)],
                       "frame_b64": base64.b64encode(frame).decode()})
    json.dump(frames, open(os.path.join(HERE, "protocol.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
