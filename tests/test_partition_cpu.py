"""Host logic of the fused spread / gather launches (no GPU needed): the guided range partition of a window's chunk list
(nfft_host.c:b200_partition_chunks) -- coverage, order, cost shares of the rounds and balance inside a round."""
import ctypes as C

import numpy as np
import pytest

import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg


def _chunks(rng, ncells, mean_points):
    """chunk descriptors {first point, count <= 8, cell key, 0} of a cell-sorted point list"""
    counts = rng.poisson(mean_points, ncells)
    ch, first = [], 0
    for key, cnt in enumerate(counts):
        while cnt > 0:
            c = min(8, cnt)
            ch.append((first, c, key, 0))
            first += c
            cnt -= c
    return np.array(ch, dtype=np.int32).reshape(-1, 4)


def _cost(ch, flush):
    keys = ch[:, 2]
    change = np.ones(len(ch), dtype=bool)
    change[1:] = keys[1:] != keys[:-1]
    return ch[:, 1] + 1.0 + flush * change


@pytest.mark.parametrize("weights", [[1.0], [8.0, 4.0, 2.0, 1.0, 1.0], [3.0, 1.0]])
@pytest.mark.parametrize("mean_points", [3.0, 40.0, 125.0])
def test_partition_covers_and_balances(weights, mean_points):
    L = pkg.lib()
    rng = np.random.default_rng(5)
    ch = _chunks(rng, 4000, mean_points)
    nch, wcnt, rounds, flush = len(ch), 37, len(weights), 60.0
    w = np.array(weights, dtype=np.float64)
    first = np.zeros(rounds * wcnt + 1, dtype=np.int32)
    assert L.Nfft4GPB200DebugPartitionChunks(pkg.capi.iptr(ch), nch, wcnt, rounds, pkg.capi.dptr(w), flush, pkg.capi.iptr(first)) == 0
    assert first[0] == 0 and first[-1] == nch and (np.diff(first) >= 0).all()          # every chunk exactly once, in order
    cost = _cost(ch, flush)
    csum = np.concatenate([[0.0], np.cumsum(cost)])
    rc = csum[first[1:]] - csum[first[:-1]]                                              # cost of every range
    biggest = cost.max()
    for k in range(rounds):
        share = rc[k * wcnt:(k + 1) * wcnt]
        assert abs(share.sum() - cost.sum() * w[k] / w.sum()) <= 2 * biggest             # round k gets its weight's share
        assert share.max() - share.min() <= 2 * biggest                                 # ranges of a round: equal cost up to one chunk


def test_partition_degenerate_inputs():
    L = pkg.lib()
    w = np.array([2.0, 1.0])
    ch = np.array([[0, 5, 7, 0]], dtype=np.int32)                                        # one chunk, 6 ranges: the rest are empty
    first = np.full(7, -1, dtype=np.int32)
    assert L.Nfft4GPB200DebugPartitionChunks(pkg.capi.iptr(ch), 1, 3, 2, pkg.capi.dptr(w), 24.0, pkg.capi.iptr(first)) == 0
    assert first[0] == 0 and first[-1] == 1 and (np.diff(first) >= 0).all() and np.diff(first).sum() == 1
    assert L.Nfft4GPB200DebugPartitionChunks(pkg.capi.iptr(ch), 0, 3, 2, pkg.capi.dptr(w), 24.0, pkg.capi.iptr(first)) == 0
    assert (first == 0).all()
    assert L.Nfft4GPB200DebugPartitionChunks(None, 1, 3, 2, pkg.capi.dptr(w), 24.0, pkg.capi.iptr(first)) == -1


@pytest.mark.parametrize("n", [1, 2, 7, 40, 120])
def test_host_tridiagonal_eigensolver(n):
    """dense_small.c:b200_stev (the dstev of the stochastic Lanczos quadrature, lanczos.c:539) against NumPy."""
    L = pkg.lib()
    rng = np.random.default_rng(n)
    d = rng.standard_normal(n) + 3.0
    e = rng.standard_normal(max(n - 1, 1))
    T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
    w = d.copy()
    ee = e.copy()
    Z = np.zeros((n, n))
    assert L.Nfft4GPB200HostStev(n, pkg.capi.dptr(w), pkg.capi.dptr(ee), pkg.capi.dptr(Z)) == 0
    Z = Z.T                                               # column-major eigenvectors
    wr = np.linalg.eigvalsh(T)
    assert np.abs(w - wr).max() <= 1e-13 * max(1.0, np.abs(wr).max())
    assert np.abs(Z.T @ Z - np.eye(n)).max() <= 1e-12
    assert np.abs(T @ Z - Z * w).max() <= 1e-12 * max(1.0, np.abs(wr).max())
