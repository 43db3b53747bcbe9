"""One request -> one reply over a connected socket (reference: distllm/compute_node/tcp_handler.py)."""
from __future__ import annotations

from dataclasses import dataclass

from ..protocol import receive_message, restore_message
from .routes import routes
from .slices import NeuralComputationError, SliceContainer
from .uploads import DiskFS, FunkyNameGenerator, MemoryFS, UploadManager, UploadRegistry


class FailingSliceContainer(SliceContainer):
    """Fault injection for handler tests (tcp_handler.py:39-44 of the reference)."""

    def load(self, f, metadata):
        raise Exception("Something went wrong")

    def forward(self, tensor):
        raise NeuralComputationError("Something went wrong")


@dataclass
class RequestContext:
    registry: UploadRegistry
    manager: UploadManager
    name_gen: FunkyNameGenerator
    slice_container: SliceContainer
    uploads_dir: str = "uploads"          # what the singleton is keyed on (NOT registry.root, which a restore rewrites)
    peers: tuple = ()                     # "host:port" of nodes this one may chain to (empty: any, see routes.py)
    self_addresses: tuple = ()            # this node's own "host:port" spellings: a route may not loop back

    @classmethod
    def default(cls, uploads_dir="uploads", names=None):
        fs = MemoryFS()
        registry = UploadRegistry(uploads_dir)
        return cls(registry, UploadManager(registry, fs), FunkyNameGenerator(names), SliceContainer(fs), uploads_dir)

    @classmethod
    def with_failing_loader(cls, uploads_dir="uploads", names=None):
        ctx = cls.default(uploads_dir, names)
        ctx.slice_container = FailingSliceContainer(MemoryFS())
        return ctx

    @classmethod
    def production(cls, uploads_dir="uploads", names=None):
        """Process-wide singletons, like the reference (slices.py:94-95, uploads.py:217-218)."""
        global _PROD
        if _PROD is None or _PROD.uploads_dir != uploads_dir:
            fs = DiskFS()
            registry = UploadRegistry(uploads_dir)
            _PROD = cls(registry, UploadManager(registry, fs), FunkyNameGenerator(names), SliceContainer(fs), uploads_dir)
        return _PROD


_PROD = None


class TCPHandler:
    def __init__(self, socket, context: RequestContext):
        self.socket = socket
        self.context = context

    def handle(self) -> None:
        name, body = receive_message(self.socket)
        message = restore_message(name, body)
        response = routes[message.get_message()](self.context, message)
        response.send(self.socket)
