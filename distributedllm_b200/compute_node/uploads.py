"""Chunked, checksummed file uploads of a compute node (reference: distllm/compute_node/uploads.py:9-218).

State machine per submission id: prepare_upload(metadata) -> upload_part(id, bytes)* -> finilize_upload(id, sha256)
(the reference's spelling is kept because it is part of the handler surface).  Only one upload may be in
flight (`ParallelUploadError`); a checksum mismatch moves the submission to `failed`.  Finished submissions are
persisted to <root>/registry_data.json and restored at start (serve.py:12-22 of the reference)."""
from __future__ import annotations

import hashlib
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional


class FailedUploadError(Exception):
    pass


class ParallelUploadError(Exception):
    pass


class UploadNotFoundError(Exception):
    pass


class DiskFS:
    """Real filesystem backend (the reference's DefaultFileSystemBackend)."""

    def open_file(self, path, mode="r"):
        return open(path, mode)

    def make_dirs(self, path):
        os.makedirs(path, exist_ok=True)

    def exists(self, path):
        return os.path.exists(path)


class MemoryFS:
    """In-memory backend for handler tests (role of the reference's FakeFileSystemBackend)."""

    class _File:
        def __init__(self, store, path, mode):
            self.store, self.path, self.mode = store, path, mode
            self.binary = "b" in mode
            if "w" in mode:
                store[path] = b""
            elif "a" not in mode and path not in store:
                raise FileNotFoundError(path)
            store.setdefault(path, b"")
            self.pos = 0

        def write(self, data):
            raw = data if self.binary else data.encode("utf-8")
            self.store[self.path] += raw
            return len(raw)

        def read(self, n=-1):
            raw = self.store[self.path][self.pos:] if n < 0 else self.store[self.path][self.pos:self.pos + n]
            self.pos += len(raw)
            return raw if self.binary else raw.decode("utf-8")

        def close(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def __init__(self):
        self.files: Dict[str, bytes] = {}

    def open_file(self, path, mode="r"):
        return MemoryFS._File(self.files, path, mode)

    def make_dirs(self, path):
        pass

    def exists(self, path):
        return path in self.files


@dataclass
class UploadLocation:
    upload_path: str
    metadata_path: str


class FunkyNameGenerator:
    """submission id <-> human name (uploads.py:199-213 of the reference)."""

    def __init__(self, names: Optional[List[str]] = None):
        self.names = list(names or [])

    def id_to_name(self, submission_id: int) -> Optional[str]:
        if 0 <= submission_id < len(self.names):
            return self.names[submission_id]
        return "slice_%d" % submission_id if not self.names else None

    def name_to_id(self, name: str) -> int:
        if name in self.names:
            return self.names.index(name)
        if name.startswith("slice_") and name[6:].isdigit():
            return int(name[6:])
        raise UploadNotFoundError(name)


class UploadRegistry:
    def __init__(self, root: str = "uploads"):
        self.root = root
        self.in_progress: List[int] = []
        self.finished: List[int] = []
        self.failed: List[int] = []
        self.next_id = 0

    def registry_data_path(self, root: Optional[str] = None) -> str:
        return os.path.join(root or self.root, "registry_data.json")

    def get_location(self, submission_id: int) -> UploadLocation:
        d = os.path.join(self.root, "slices", "upload_%d" % submission_id)
        return UploadLocation(os.path.join(d, "uploaded_file"), os.path.join(d, "metadata.json"))

    def allocate(self) -> int:
        if self.in_progress:
            raise ParallelUploadError()
        sid = self.next_id
        self.next_id += 1
        self.in_progress.append(sid)
        return sid

    def mark(self, sid: int, ok: bool) -> None:
        if sid in self.in_progress:
            self.in_progress.remove(sid)
        (self.finished if ok else self.failed).append(sid)

    def state_dict(self) -> dict:
        return {"root": self.root, "finished": self.finished, "failed": self.failed, "next_id": self.next_id}

    def load_state_dict(self, state: dict) -> None:
        self.root = state.get("root", self.root)
        self.finished = list(state.get("finished", []))
        self.failed = list(state.get("failed", []))
        self.next_id = int(state.get("next_id", (max(self.finished + self.failed) + 1) if self.finished or self.failed else 0))
        self.in_progress = []


class UploadManager:
    def __init__(self, registry: UploadRegistry, fs_backend=None):
        self.registry = registry
        self.fs_backend = fs_backend or DiskFS()
        self._hashers: Dict[int, "hashlib._Hash"] = {}
        self._sizes: Dict[int, int] = {}

    def prepare_upload(self, metadata: dict) -> int:
        sid = self.registry.allocate()
        loc = self.registry.get_location(sid)
        self.fs_backend.make_dirs(os.path.dirname(loc.upload_path))
        with self.fs_backend.open_file(loc.metadata_path, "w") as f:
            f.write(json.dumps(metadata))
        with self.fs_backend.open_file(loc.upload_path, "wb") as f:
            f.write(b"")
        self._hashers[sid] = hashlib.sha256()
        self._sizes[sid] = 0
        return sid

    def upload_part(self, submission_id: int, data: bytes) -> int:
        if submission_id not in self.registry.in_progress or submission_id not in self._hashers:
            raise UploadNotFoundError(submission_id)
        loc = self.registry.get_location(submission_id)
        with self.fs_backend.open_file(loc.upload_path, "ab") as f:
            f.write(data)
        self._hashers[submission_id].update(data)
        self._sizes[submission_id] += len(data)
        return len(data)

    def finilize_upload(self, submission_id: int, checksum: str) -> int:
        if submission_id not in self.registry.in_progress or submission_id not in self._hashers:
            raise FileNotFoundError(submission_id)
        ok = self._hashers.pop(submission_id).hexdigest() == checksum
        size = self._sizes.pop(submission_id)
        self.registry.mark(submission_id, ok)
        if not ok:
            raise FailedUploadError(submission_id)
        self._persist()
        return size

    def _persist(self) -> None:
        try:
            self.fs_backend.make_dirs(self.registry.root)
            with self.fs_backend.open_file(self.registry.registry_data_path(), "w") as f:
                f.write(json.dumps(self.registry.state_dict()))
        except OSError:
            pass
