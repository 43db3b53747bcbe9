"""Per-hop cost of the tensor RPC: the reference's float-list wire vs the additive binary wire, against a
live node serving a 7B Q4_0 slice on this GPU.  One process: node server thread + client.  Prints one JSON object."""
import json, os, sys, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from distributedllm_b200.compute_node import serve
from distributedllm_b200.compute_node.tcp_handler import RequestContext
from distributedllm_b200.control_center import Connection

L = int(os.environ.get("WIRE_LAYERS", "32"))
path = bench.slice_file("7b", 0, L - 1)
ctx = RequestContext.default("wire_uploads")
ctx.slice_container.load(path, {"model": "llama-7b", "layer_from": 0, "layer_to": L - 1})
srv = serve.make_server("127.0.0.1", 0, "wire_uploads", context=ctx)
threading.Thread(target=srv.serve_forever, daemon=True).start()
conn = Connection(("127.0.0.1", srv.server_address[1]))
E = 4096
out = {"slice": "LLaMA-7B Q4_0, %d layers" % L}
for n_tok, reps in ((1, 60), (64, 12)):
    x = bench.synth_inputs(n_tok, E, 9)
    res = {}
    for wire in ("list", "bytes"):
        ts = []
        for r in range(reps):
            conn.clear_context()
            t0 = time.perf_counter()
            if wire == "list":
                y = conn.propagate_forward(x.ravel().tolist(), (1, n_tok * E))["values"]
            else:
                y = conn.propagate_forward_bytes(x, (1, n_tok * E))
            ts.append(time.perf_counter() - t0)
        res[wire + "_ms"] = round(1e3 * statistics.median(ts[2:]), 3)
    # the slice forward alone (same call the node makes), for scale
    sl = ctx.slice_container.slice.llm
    ts = []
    for r in range(reps):
        sl.clear_context()
        t0 = time.perf_counter()
        sl.propagate_forward_buffer(x)
        ts.append(time.perf_counter() - t0)
    res["forward_only_ms"] = round(1e3 * statistics.median(ts[2:]), 3)
    out["%d_token%s" % (n_tok, "s" if n_tok > 1 else "")] = res
print(json.dumps(out))
srv.shutdown()
