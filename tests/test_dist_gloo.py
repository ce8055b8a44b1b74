"""World-size-2 CPU (gloo) test of the row-sharded algorithm: the exchange steps the library
performs with NCCL -- the per-step (value, global index, basis row) all-gather of DEIM and the
Gram all-reduce of the POD passes -- restated on the oracle and driven
through torch.distributed, must reproduce the unsharded reference result.  DEIM is covered in its
streaming form and in the blocked form the library runs for k > 16 (panels / sub-panels local to
each rank, one record all-gather per step), the latter also with pipelined ingest, where the
deferred columns of the chosen rows are caught up by an all-reduce at every panel boundary."""
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
    # the blocked path (panels of 16, sub-panels of 4) with k > 16, plain and with pipelined ingest (deferred columns
    # caught up by an all-reduce at every panel boundary), incl. an exact tie across the shard boundary
    Qb, _ = np.linalg.qr(rng.standard_normal((1501, 40)))
    Qb[1200] = Qb[7]
    rb0, mb = shard_rows(Qb.shape[0], world, rank)
    idx_b = O.deim_interpolation_indices_blocked_sharded(Qb[rb0:rb0 + mb], rb0, allgather, allreduce_sum)
    idx_p = O.deim_interpolation_indices_blocked_sharded(Qb[rb0:rb0 + mb], rb0, allgather, allreduce_sum,
                                                         cols_ready=lambda l: min(40, (l // 16 + 1) * 16))
    X = rng.standard_normal((2001, 24)) * np.logspace(0, -4, 24)
    r0, ml = shard_rows(X.shape[0], world, rank)
    s, V = O.pod_svd_sharded(X[r0:r0 + ml], allreduce_sum)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx, s=s, V=V, full_idx=O.deim_interpolation_indices(Q),
             idx_b=idx_b, idx_p=idx_p, full_idx_b=O.deim_interpolation_indices(Qb),
             s_ref=np.linalg.svd(X, compute_uv=False))
    dist.destroy_process_group()


def test_sharded_deim_and_pod_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for o in outs:
        assert o["idx"].tolist() == o["full_idx"].tolist()          # bit-exact, tie-break included
        assert np.max(np.abs(o["s"] - o["s_ref"]) / o["s_ref"][0]) < 1e-12
        assert o["idx_b"].tolist() == o["full_idx_b"].tolist()      # blocked path, sharded
        assert o["idx_p"].tolist() == o["full_idx_b"].tolist()      # ... with pipelined ingest / deferred columns
    assert outs[0]["idx"].tolist() == outs[1]["idx"].tolist()        # replicated result
    assert np.array_equal(outs[0]["s"], outs[1]["s"])                # bitwise identical on both ranks
