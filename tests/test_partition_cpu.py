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


def test_poly1d_block_local_cell_sort():
    """Plan build of the table-free 1-D path (nfft_host.c:b200_poly1d_block_sort, host only): inside every block of T consecutive points
    the points are ordered by grid cell (inside a cell: dealt out by bank class); lp is a permutation of the block, rs the run starts per cell key, and the stored
    coordinate is the offset 2 u - 1 of the point inside its cell (exact)."""
    import ctypes as C
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    L = pkg.lib()
    rng = np.random.default_rng(3)
    for n, T in ((1000, 256), (8192 * 2 + 77, 8192), (50, 64), (6784 * 3, 6784)):
        xt = np.clip((rng.random(n) - 0.5) * 0.5, -0.25, 0.25)
        xt[0], xt[-1] = 0.25, -0.25                      # the two extreme cells (keys 48 and 16)
        nblk = (n + T - 1) // T
        oxt, olp, ors = np.zeros(nblk * T), np.zeros(nblk * T, dtype=np.uint16), np.zeros(nblk * 80, dtype=np.uint16)
        assert L.Nfft4GPB200DebugPoly1dBlockSort(pkg.capi.dptr(xt), n, T, pkg.capi.dptr(oxt), olp.ctypes.data_as(C.POINTER(C.c_ushort)),
                                                 ors.ctypes.data_as(C.POINTER(C.c_ushort))) == 0
        key = (np.floor(xt * 64).astype(int) + 32) & 63
        for b in range(nblk):
            base, cnt = b * T, min(T, n - b * T)
            lp, rs = olp[base:base + cnt].astype(int), ors[b * 80:b * 80 + 65].astype(int)
            assert sorted(lp.tolist()) == list(range(cnt))                       # a permutation of the block
            t = xt[base + lp] * 64
            assert (oxt[base:base + cnt] == 2.0 * (t - np.floor(t)) - 1.0).all()  # in-cell offsets follow it
            kb = key[base + lp]
            assert (np.diff(kb) >= 0).all()                                      # ordered by cell
            assert rs[0] == 0 and rs[64] == cnt and (np.diff(rs) >= 0).all()
            for c in np.unique(kb):
                seg = lp[rs[c]:rs[c + 1]]
                assert (kb[rs[c]:rs[c + 1]] == c).all()                          # the run of cell c
                # dealt out by shared-memory bank class (lp mod 16), every class in original order: while all classes have points
                # left, each group of 16 consecutive entries of the run holds 16 different classes
                full = 16 * np.bincount(seg % 16, minlength=16).min()
                assert all(len(set((seg[g:g + 16] % 16).tolist())) == 16 for g in range(0, full, 16))
                for q in range(16):
                    assert (np.diff(seg[seg % 16 == q]) > 0).all()
    assert L.Nfft4GPB200DebugPoly1dBlockSort(pkg.capi.dptr(np.zeros(4)), 4, 16, None, None, None) == -1
