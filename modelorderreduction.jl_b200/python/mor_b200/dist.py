"""Row-sharded multi-GPU mode: one process per GPU.

`shard_rows` is the partition every rank must agree on (contiguous row ranges in rank
order -- that is what makes "lowest global index" reproduce Julia's first-index tie-break
across shard boundaries).  `init_from_torch` wires the library's NCCL communicator using
an already-initialised torch.distributed process group to broadcast the NCCL unique id;
in a Julia host the same 128 bytes travel over Distributed/MPI."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from . import _lib


def shard_rows(m_global: int, nranks: int, rank: int, align: int = 4) -> Tuple[int, int]:
    """Rows [row0, row0 + m_local) of rank `rank`.  Shards are contiguous, ordered by rank,
    cover [0, m_global) exactly and start on multiples of `align` rows (16-byte aligned
    columns, whole Philox blocks)."""
    if nranks < 1 or not (0 <= rank < nranks) or m_global < 0:
        raise ValueError("bad shard request")
    blocks = (m_global + align - 1) // align
    base, extra = divmod(blocks, nranks)
    b0 = rank * base + min(rank, extra)
    nb = base + (1 if rank < extra else 0)
    row0 = min(b0 * align, m_global)
    row1 = min((b0 + nb) * align, m_global)
    return row0, row1 - row0


def init_multi(ngpus: int) -> None:
    """Single-process multi-GPU mode: this process drives `ngpus` GPUs, the library shards the rows of every host
    matrix handed to the host-pointer entry points (mor_b200_init_multi; MOR_B200_NGPUS does the same lazily)."""
    _lib.check(_lib.load().mor_b200_init_multi(ngpus))


def multi_info() -> int:
    n = C.c_int32(0)
    _lib.check(_lib.load().mor_b200_multi_info(C.byref(n)))
    return n.value


def init(device: int) -> None:
    _lib.check(_lib.load().mor_b200_init(device))


def init_from_torch(device: int) -> Tuple[int, int]:
    """Call on every rank after torch.distributed.init_process_group(...)."""
    import torch
    import torch.distributed as dist

    init(device)
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _lib.load()
    buf = (C.c_ubyte * 128)()
    if world > 1 and rank == 0:
        _lib.check(lib.mor_b200_comm_unique_id(buf))
    if world > 1:
        backend = dist.get_backend()
        dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        data = bytes(t.cpu().tolist())
        buf = (C.c_ubyte * 128).from_buffer_copy(data)
    _lib.check(lib.mor_b200_comm_init(rank, world, buf))
    return rank, world


def comm_info() -> Tuple[int, int]:
    r, n = C.c_int32(0), C.c_int32(1)
    _lib.check(_lib.load().mor_b200_comm_info(C.byref(r), C.byref(n)))
    return r.value, n.value
