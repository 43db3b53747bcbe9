"""`run_node`: the compute-node server (reference: distllm/compute_node/serve.py:11-82).

One message per TCP connection, threads per connection; the slice forward itself is serialised by the slice
handle's mutex and releases the GIL while the GPU works."""
from __future__ import annotations

import json
import os
import socket
import socketserver

from .. import protocol
from .tcp_handler import RequestContext, TCPHandler

FUNKY_NAMES = ["orb", "pranker", "human", "alien", "sorcerer"]


def restore_registry(ctx: RequestContext, uploads_dir: str) -> None:
    path = ctx.registry.registry_data_path(uploads_dir)
    if os.path.isfile(path):
        with open(path) as f:
            ctx.registry.load_state_dict(json.loads(f.read()))
    # the directory the node was STARTED with wins over the root spelled in a restored registry_data.json ("uploads" vs
    # "./uploads" vs an absolute path, or a registry written elsewhere): locations are derived from the root
    ctx.registry.root = uploads_dir


class _Handler(socketserver.BaseRequestHandler):
    uploads_dir = "uploads"
    context = None          # an explicit RequestContext (several nodes in one process); None = process-wide singleton

    def handle(self):
        TCPHandler(self.request, self.context or RequestContext.production(self.uploads_dir, FUNKY_NAMES)).handle()


class ThreadingTCPServer(socketserver.ThreadingMixIn, socketserver.TCPServer):
    allow_reuse_address = True
    daemon_threads = True


def make_server(host: str, port: int, uploads_dir: str, context: RequestContext = None) -> ThreadingTCPServer:
    ctx = context or RequestContext.production(uploads_dir, FUNKY_NAMES)
    restore_registry(ctx, uploads_dir)
    handler = type("NodeHandler", (_Handler,), {"uploads_dir": uploads_dir, "context": context})
    return ThreadingTCPServer((host, port), handler)


def run_server(host, port, uploads_dir, reverse_connect=False):
    if reverse_connect:
        return connect_then_serve(host, port, uploads_dir)
    with make_server(host, port, uploads_dir) as server:
        server.serve_forever()


def connect_then_serve(host, port, uploads_dir="uploads"):
    """Dial out to a proxy and serve over that one connection (serve.py:35-64 of the reference)."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.connect((host, port))
        protocol.RequestGreeting().send(sock)
        name, body = protocol.receive_message(sock)
        if protocol.restore_message(name, body) != protocol.ResponseGreeting():
            raise Exception("Failed to reverse connect: handshake failed")
        handler = TCPHandler(sock, RequestContext.production(uploads_dir, FUNKY_NAMES))
        while True:
            try:
                handler.handle()
            except KeyboardInterrupt:
                break
