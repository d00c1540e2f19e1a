"""Sample-axis data parallelism: one process per GPU, each holding a contiguous block of
sample columns; the only exchange is one small NCCL allreduce of the packed N x N moment buffer per pass,
issued by libpicard_b200.so itself on its own stream.  `torch.distributed` is plumbing only: it carries the
NCCL unique id from rank 0 to the other ranks (works over gloo or nccl) and provides barriers for timing.
The reference has no distributed code at all; nothing here mirrors a reference interface.
"""
from __future__ import annotations

import ctypes as C

from . import _ffi

__all__ = ["shard_range", "Communicator", "broadcast_unique_id"]


def shard_range(n_samples: int, rank: int, world_size: int) -> tuple[int, int]:
    """[begin, end) of the sample columns rank `rank` owns: contiguous blocks, sizes differing by at most one,
    every begin a multiple of 2 when n_samples allows (TMA needs 16-byte aligned row starts)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    if n_samples < 0:
        raise ValueError("n_samples must be >= 0")
    pairs = n_samples // 2  # distribute pairs of samples so every shard start is even
    base, rem = divmod(pairs, world_size)
    begin = 2 * (rank * base + min(rank, rem))
    end = 2 * ((rank + 1) * base + min(rank + 1, rem))
    if rank == world_size - 1:
        end = n_samples  # odd leftover sample goes to the last rank
    return begin, end


def broadcast_unique_id(unique_id: bytes | None, src: int = 0) -> bytes:
    """Broadcast the 128-byte NCCL unique id from rank `src` through torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(_ffi.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        assert unique_id is not None and len(unique_id) == _ffi.UNIQUE_ID_BYTES
        buf.copy_(torch.frombuffer(bytearray(unique_id), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


class Communicator:
    """Owns a `picard_comm_t*` (NCCL communicator over the ranks of the current torch.distributed group)."""

    def __init__(self, handle, rank: int, size: int):
        self.handle = handle
        self.rank = rank
        self.size = size

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_ffi.UNIQUE_ID_BYTES)
        st = _ffi.lib().picard_comm_unique_id(buf)
        if st != 0:
            raise RuntimeError("picard_comm_unique_id failed (is NCCL loadable?)")
        return buf.raw

    @classmethod
    def create(cls, unique_id: bytes, rank: int, size: int, device: int) -> "Communicator":
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        st = _ffi.lib().picard_comm_create(C.byref(h), unique_id, C.c_int32(rank), C.c_int32(size), C.c_int32(device), err,
                                           C.c_size_t(1024))
        if st != 0:
            raise RuntimeError(err.value.decode() or "picard_comm_create failed")
        return cls(h, rank, size)

    @classmethod
    def from_torch_distributed(cls, device: int) -> "Communicator":
        import torch.distributed as dist
        rank, size = dist.get_rank(), dist.get_world_size()
        uid = cls.unique_id() if rank == 0 else None
        uid = broadcast_unique_id(uid, 0)
        return cls.create(uid, rank, size, device)

    def close(self):
        if self.handle:
            _ffi.lib().picard_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
