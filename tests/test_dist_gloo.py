"""World-size-2 CPU (gloo) test of the row-sharded algorithm: the exchange steps the library
performs with NCCL -- the per-step (value, global index, basis row) all-gather of DEIM and the
Gram all-reduce of the POD passes -- restated on the oracle and driven
through torch.distributed, must reproduce the unsharded reference result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    for p in (os.path.join(ROOT, "modelorderreduction.jl_b200", "python"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import mor_oracle as O
    from mor_b200.dist import shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allgather(obj):
        res = [None] * world
        dist.all_gather_object(res, obj)
        return res

    def allreduce_sum(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    rng = np.random.Generator(np.random.PCG64(42))           # same data on every rank
    Q, _ = np.linalg.qr(rng.standard_normal((1003, 16)))
    Q[700] = Q[3]                                             # exact tie across the shard boundary
    row0, m_loc = shard_rows(Q.shape[0], world, rank)
    idx = O.deim_interpolation_indices_sharded(Q[row0:row0 + m_loc], row0, allgather)
    X = rng.standard_normal((2001, 24)) * np.logspace(0, -4, 24)
    r0, ml = shard_rows(X.shape[0], world, rank)
    s, V = O.pod_svd_sharded(X[r0:r0 + ml], allreduce_sum)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx, s=s, V=V, full_idx=O.deim_interpolation_indices(Q),
             s_ref=np.linalg.svd(X, compute_uv=False))
    dist.destroy_process_group()


def test_sharded_deim_and_pod_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for o in outs:
        assert o["idx"].tolist() == o["full_idx"].tolist()          # bit-exact, tie-break included
        assert np.max(np.abs(o["s"] - o["s_ref"]) / o["s_ref"][0]) < 1e-12
    assert outs[0]["idx"].tolist() == outs[1]["idx"].tolist()        # replicated result
    assert np.array_equal(outs[0]["s"], outs[1]["s"])                # bitwise identical on both ranks
