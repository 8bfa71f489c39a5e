"""World-size-2 gloo tests (CPU) of the multi-GPU decomposition of the matvec: the sharding arithmetic is
checked with the CPU oracle standing in for the per-rank device work.
  mode 1, by feature window : rank r computes the windows w = r (mod G); ONE allreduce(sum) of the partial output vector.
  mode 2, by point block    : rank r spreads its block of points; ONE allreduce(sum) of the oversampled grids; every rank
                              interpolates at its own block.  (Checked on the spread stage with the NumPy restatement.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_api as O
    import np_fastsum as NP

    rng = np.random.default_rng(7)
    n, d, P, dw = 600, 9, 3, 3
    X = rng.random((d, n)) * n ** (1.0 / d)
    x = rng.random(n) - 0.5
    y0 = rng.random(n) - 0.5
    f, l, mu, alpha, beta = 1.1, 0.5, 0.02, 0.7, 0.3
    win = np.arange(9, dtype=np.int32)
    full = O.OracleAdditiveNFFT(X, win, P, dw, "gaussian").build(f, l, mu).matvec(x, alpha, beta, y0)

    # ---- mode 1: window sharding.  partial = sum over owned windows of f^2 (K_w a + mu a), a = (alpha / P) x
    part = np.zeros(n)
    for w in range(P):
        if w % world != rank:
            continue
        one = O.OracleAdditiveNFFT(X, win[w * dw:(w + 1) * dw].copy(), 1, dw, "gaussian").build(f, l, mu)
        part += one.matvec(x, alpha / P, 0.0)  # single-window additive handle: f^2 (K_w + mu I) (alpha/P) x
    t = torch.from_numpy(part)
    dist.all_reduce(t)  # the ONE collective of the matvec
    y1 = beta * y0 + t.numpy()
    err1 = float(np.linalg.norm(y1 - full) / np.linalg.norm(full))

    # ---- mode 2: point-block sharding of the spread: sum of per-rank grids == grid of all points
    pts, _ = NP.scale_points(X[:3])
    lo, hi = rank * n // world, (rank + 1) * n // world
    c = np.floor(pts * 64).astype(int)
    lj = np.arange(10)
    psi = NP.kb_phi(pts[:, :, None] - (c[:, :, None] - 4 + lj) / 64)
    idx = (c[:, :, None] - 4 + lj) % 64

    def spread(j0, j1):
        g = np.zeros((64, 64, 64))
        for j in range(j0, j1):
            np.add.at(g, np.ix_(idx[j, 0], idx[j, 1], idx[j, 2]), x[j] * np.einsum("a,b,c->abc", psi[j, 0], psi[j, 1], psi[j, 2]))
        return g

    gl = torch.from_numpy(spread(lo, hi))
    dist.all_reduce(gl)
    err2 = float(np.abs(gl.numpy() - spread(0, n)).max() / np.abs(gl.numpy()).max()) if rank == 0 else 0.0
    ret[rank] = (err1, err2)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_window_and_block_sharding():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r][0] <= 1e-13, ret[r]
    assert ret[0][1] <= 1e-13


def _sharded_precond_worker(rank, world, port, q):
    """NumPy restatement of the point-block-sharded AFN and Nystrom applies (precond_host.c: afn_solve_sharded,
    nys_solve_sharded) with gloo collectives, against the unsharded oracle applies (oracle/np_setup.py)."""
    sys.path.insert(0, ROOT)
    os.environ["OMP_NUM_THREADS"] = "2"
    from oracle import np_setup as S

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def allreduce(v):
        t = torch.from_numpy(np.ascontiguousarray(v))
        dist.all_reduce(t)
        return t.numpy()

    n, d, k, lfil = 400, 6, 30, 12
    rng = np.random.default_rng(3)
    X = rng.random((d, n)) * n ** (1.0 / d)
    wins = [[0, 1, 2], [3, 4, 5]]
    f, l, mu = 1.1, 0.9, 0.05
    r = rng.random(n) - 0.5
    lo, hi = rank * n // world, (rank + 1) * n // world
    r_loc = r[lo:hi]
    # ---- AFN
    F = S.afn_setup(X, k, lfil, f, l, mu, "gaussian", wins)
    perm, Lc, K12, G = F["perm"], F["L"], F["K12"], F["G"]
    n2 = n - k
    a11 = lambda v: np.linalg.solve(Lc.T, np.linalg.solve(Lc, v))
    lm = np.array([p - lo if lo <= p < hi else -1 for p in perm[:k]])
    sch = perm[k:]
    plo = int(np.searchsorted(sch, lo))
    phi = int(np.searchsorted(sch, hi))
    sc = sch[plo:phi] - lo
    rp1 = allreduce(np.where(lm >= 0, r_loc[np.maximum(lm, 0)], 0.0))
    y1 = a11(rp1)
    rp2l = r_loc[sc] - K12[:, plo:phi].T @ y1
    full = np.zeros(n2)
    full[plo:phi] = rp2l
    full = allreduce(full)
    w = np.zeros(n2)
    w[plo:phi] = G[plo:phi, :] @ full
    w = allreduce(w)
    y2l = G.T[plo:phi, :] @ w
    u = allreduce(K12[:, plo:phi] @ y2l)
    y1 = a11(rp1 - u)
    out = np.zeros(hi - lo)
    out[lm[lm >= 0]] = y1[lm >= 0]
    out[sc] = y2l
    ref = S.afn_apply(F, r)[lo:hi]
    e_afn = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
    # ---- Nystrom
    permn = np.random.default_rng(4).permutation(n).astype(np.int32)
    U, s, eta = S.nys_setup(X, permn, k, f, l, mu, "gaussian", wins)
    pos = np.zeros(hi - lo, dtype=int)
    for i, p in enumerate(permn):
        if lo <= p < hi:
            pos[p - lo] = i
    UT = U[pos, :].T
    z = allreduce(UT @ r_loc)
    outn = r_loc / eta + UT.T @ ((s - 1.0 / eta) * z)
    refn = S.nys_apply(permn, U, s, eta, r)[lo:hi]
    e_nys = float(np.linalg.norm(outn - refn) / np.linalg.norm(refn))
    q.put((rank, e_afn, e_nys))
    dist.destroy_process_group()


def test_sharded_preconditioner_algebra_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_precond_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, e_afn, e_nys in res:
        assert e_afn <= 1e-10, (rank, e_afn)
        assert e_nys <= 1e-10, (rank, e_nys)


def _probe_sharding_worker(rank, world, port, q):
    """Probe sharding of the stochastic Lanczos quadrature (lanczos_host.c: Nfft4GPB200SetProbeSharding; BASELINE configs[4]): probe i
    on rank i (mod R), every rank keeps the whole operator, ONE allreduce of the four sums.  NumPy restatement of the quadrature
    (oracle/np_solvers.py) with gloo, against the single-process run over all probes."""
    sys.path.insert(0, ROOT)
    os.environ["OMP_NUM_THREADS"] = "2"
    from oracle import np_setup as S
    from oracle import np_solvers as NS

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, d, nv, maxits = 200, 6, 6, 30
    rng = np.random.default_rng(9)
    X = rng.random((d, n)) * n ** (1.0 / d)
    wins = [[0, 1, 2], [3, 4, 5]]
    f, l, mu = 1.1, 0.9, 0.05
    K = S.dense_matrix(X, f, l, mu, "gaussian", wins)
    h = 1e-6
    dKl = (S.dense_matrix(X, f, l + h, mu, "gaussian", wins) - S.dense_matrix(X, f, l - h, mu, "gaussian", wins)) / (2 * h)
    mv = lambda v: K @ v
    gmv = lambda v: np.concatenate([2.0 / f * (K @ v), dKl @ v, f * f * v])
    probes = np.where(np.random.default_rng(10).random((nv, n)) < 0.5, -1.0, 1.0)      # every rank draws every probe
    mine = [probes[i] for i in range(nv) if i % world == rank]
    val, dval = NS.slq_logdet(mv, gmv, n, maxits, mine)
    sums = np.concatenate([[val * len(mine)], dval * len(mine) * n])                   # un-normalised local sums
    t = torch.from_numpy(sums)
    dist.all_reduce(t)                                                                  # the one collective
    val_sh, dval_sh = t.numpy()[0] / nv, t.numpy()[1:] / nv / n
    val_1, dval_1 = NS.slq_logdet(mv, gmv, n, maxits, probes)
    q.put((rank, abs(val_sh - val_1) / abs(val_1), float(np.abs(dval_sh - dval_1).max() / np.abs(dval_1).max())))
    dist.destroy_process_group()


def test_probe_sharded_lanczos_quadrature_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_probe_sharding_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, e_val, e_dval in res:
        assert e_val <= 1e-12 and e_dval <= 1e-12, (rank, e_val, e_dval)
