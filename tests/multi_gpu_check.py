"""Row-sharded POD / RSVD / DEIM on N GPUs (one process per GPU, NCCL) against the unsharded CPU
oracle.  Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
               --master-port P tests/multi_gpu_check.py
Every rank generates the same host matrix from a seed, uploads only its row shard, and checks the
replicated results (spectrum, indices) and its shard of the basis."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "modelorderreduction.jl_b200", "python"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import mor_oracle as O
    from mor_b200 import _lib
    from mor_b200 import dist as mdist
    from mor_b200.device import DeviceMatrix, deim_indices_dev, pod_factor_dev

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mdist.init_from_torch(local)
    assert mdist.comm_info() == (rank, world)

    def gather_rows(local_rows):
        t = torch.from_numpy(np.ascontiguousarray(local_rows)).cuda()
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device="cuda"))
        mx = max(int(s.item()) for s in sizes)
        pad = torch.zeros((mx, t.shape[1]), dtype=t.dtype, device="cuda"); pad[:t.shape[0]] = t
        outs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        return np.concatenate([o[:int(s.item())].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)

    rng = np.random.default_rng(123)
    # ---- POD SVD: one-pass (column-scaled) and multi-pass (correlated columns) ------------------
    for name, X in (("graded", rng.standard_normal((300_000, 128)) * np.logspace(0, -3, 128)),
                    ("correlated", (np.linalg.qr(rng.standard_normal((40_001, 40)))[0] * np.logspace(0, -7, 40))
                     @ np.linalg.qr(rng.standard_normal((40, 40)))[0].T)):
        M, n = X.shape
        k = 12
        row0, m_loc = mdist.shard_rows(M, world, rank)
        dX = DeviceMatrix.from_host(X[row0:row0 + m_loc])
        f = pod_factor_dev(dX, _lib.MOR_SVD, k, flags=_lib.POD_MAY_OVERWRITE, m_global=M)
        s_ref = np.linalg.svd(X, compute_uv=False)
        assert f.s.shape == (n,), f.s.shape
        assert O.sigma_error(s_ref, f.s) <= 1e-10, (name, O.sigma_error(s_ref, f.s))
        U_loc = f.basis(k, DeviceMatrix.alloc(m_loc, k)).to_host()
        U = gather_rows(U_loc)
        assert O.subspace_sine(O._svd(X)[0][:, :k], U) <= 1e-8, name
        # replicated results are bit-identical on every rank
        s_all = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(s_all, torch.from_numpy(f.s).cuda())
        assert all(torch.equal(s_all[0], t) for t in s_all)
        if rank == 0:
            print(f"POD {name}: passes={f.info()[0]} sigma_err={O.sigma_error(s_ref, f.s):.2e} ok", flush=True)
        f.free(); dX.free()

    # ---- RSVD with a shared Omega -----------------------------------------------------------------
    X = rng.standard_normal((100_003, 96)) * np.logspace(0, -3, 96); X[:, :8] *= 20
    omega = rng.standard_normal((96, 10))
    M = X.shape[0]
    row0, m_loc = mdist.shard_rows(M, world, rank)
    dX, dO = DeviceMatrix.from_host(X[row0:row0 + m_loc]), DeviceMatrix.from_host(omega)
    f = pod_factor_dev(dX, _lib.MOR_RSVD, 8, p=2, omega=dO, m_global=M)
    po = O.POD(X, 8); O.reduce(po, O.RSVD(2), omega=omega)
    assert O.sigma_error(po.spectrum, f.s) <= 1e-10
    U = gather_rows(f.basis(8, DeviceMatrix.alloc(m_loc, 8)).to_host())
    assert O.subspace_sine(po.rbasis, U) <= 1e-8
    if rank == 0:
        print("RSVD ok", flush=True)

    # ---- known-answer matrix at 6M x 128, generated shard by shard on the devices -------------------
    M, n, k = 6_000_002, 128, 16
    row0, m_loc = mdist.shard_rows(M, world, rank)
    Xk = DeviceMatrix.alloc(m_loc, n).synth(_lib.SYNTH_KAT, seed=11, global_row0=row0, param=3.0)
    f = pod_factor_dev(Xk, _lib.MOR_SVD, k, flags=_lib.POD_MAY_OVERWRITE, m_global=M)
    s_exact = 10.0 ** (-3.0 * np.arange(n) / (n - 1))
    assert O.sigma_error(s_exact, f.s) <= 1e-10, O.sigma_error(s_exact, f.s)
    err = f.basis(k, DeviceMatrix.alloc(m_loc, k)).kat_check(k, n, seed=11, global_row0=row0, param=3.0)
    assert err <= 1e-8, err
    if rank == 0:
        print(f"KAT 6Mx128: passes={f.info()[0]} sigma_err={O.sigma_error(s_exact, f.s):.2e} left-vector err={err:.2e} ok", flush=True)
    f.free(); Xk.free()

    # ---- DEIM: exact ties across the shard boundary, bit-exact indices ----------------------------
    Q, _ = np.linalg.qr(rng.standard_normal((200_001, 40)))
    r_a, r_b = 5, 150_000
    Q[r_b] = Q[r_a]
    row0, m_loc = mdist.shard_rows(Q.shape[0], world, rank)
    idx = deim_indices_dev(DeviceMatrix.from_host(Q[row0:row0 + m_loc]))
    want = O.deim_interpolation_indices(Q)
    assert idx.tolist() == want.tolist(), (idx.tolist(), want.tolist())
    T = np.zeros((100_000, 2)); T[:, 0] = 1.0; T[70_000, 0] = -4.0; T[20_000, 0] = 4.0; T[:, 1] = np.arange(100_000) % 5
    row0, m_loc = mdist.shard_rows(T.shape[0], world, rank)
    assert deim_indices_dev(DeviceMatrix.from_host(T[row0:row0 + m_loc])).tolist() == O.deim_interpolation_indices(T).tolist()
    # k = 256 (the headline dimension) sharded: blocked and streaming paths, both against the oracle
    Q, _ = np.linalg.qr(rng.standard_normal((131_072, 256)))
    row0, m_loc = mdist.shard_rows(Q.shape[0], world, rank)
    dQ = DeviceMatrix.from_host(Q[row0:row0 + m_loc])
    want, margins = O.deim_interpolation_indices(Q, return_margins=True)
    got = deim_indices_dev(dQ)
    t = _lib.last_timings()
    assert got.tolist() == want.tolist(), "k=256 blocked path differs from the oracle"
    assert t["deim_blocked"] == 1.0 or t["deim_fallback"] == 1.0
    want_exchange = 2.0 if os.environ.get("MOR_B200_DEIM_EXCHANGE") == "nccl" else 1.0
    assert t["deim_exchange"] == want_exchange, f"per-step exchange mode {t['deim_exchange']} (1 = peer mailboxes, 2 = NCCL)"
    import mor_b200
    mg = mor_b200.deim_last_margins()[0]
    if t["deim_fallback"] == 0.0:
        np.testing.assert_allclose(mg[1:], margins[1:], rtol=1e-6, atol=1e-9)
    os.environ["MOR_B200_DEIM_BLOCK"] = "0"
    got_s = deim_indices_dev(dQ)
    del os.environ["MOR_B200_DEIM_BLOCK"]
    assert got_s.tolist() == want.tolist(), "k=256 streaming path differs from the oracle"
    if rank == 0:
        print(f"DEIM ok on {world} GPUs (k=256 blocked + streaming vs oracle; per-step exchange: "
              f"{'peer mailboxes over NVLink' if want_exchange == 1.0 else 'ncclAllGather'})", flush=True)

    # ---- BASELINE config 5 at oracle size: Burgers snapshots generated shard by shard, POD basis, DEIM ----
    nx, n = 192, 48
    M = nx * nx
    row0, m_loc = mdist.shard_rows(M, world, rank)
    S = DeviceMatrix.alloc(m_loc, n).synth(_lib.SYNTH_BURGERS, seed=2, global_row0=row0, param=1e-3, iparam=nx)
    S_all = gather_rows(S.to_host())
    assert np.abs(S_all - O.burgers_snapshots(M, n, nx, 1e-3, seed=2)).max() <= 1e-11
    f = pod_factor_dev(S, _lib.MOR_SVD, n, m_global=M)
    assert O.sigma_error(np.linalg.svd(S_all, compute_uv=False), f.s) <= 1e-10
    Ub = f.basis(n, DeviceMatrix.alloc(m_loc, n))
    idx = deim_indices_dev(Ub)
    assert _lib.last_timings()["deim_blocked"] == 1.0 or _lib.last_timings()["deim_fallback"] == 1.0
    assert idx.tolist() == O.deim_interpolation_indices(gather_rows(Ub.to_host())).tolist()
    if rank == 0:
        print(f"Burgers config 5 (sharded generator, POD basis, blocked DEIM) ok on {world} GPUs", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
