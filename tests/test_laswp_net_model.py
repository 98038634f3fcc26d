"""CPU model of the node-level row interchange (csrc/laswp.cu, laswp_compose_kernel + laswp_net_kernel).

The device code composes the per-panel exchange lists K1 emits into the net permutation of a pivot range and applies
it in one pass per column with three kinds of moves only (block <- block, block <- below, below <- block).  This
model restates exactly that data flow in numpy and checks it against plain sequential swaps
(apply_permutation!, src/lu.jl:164-188) -- including the claim that a below-block row never receives content from
another below-block row, chunking of long ranges, repeated pivot targets and identity steps."""
import numpy as np
import pytest


def panel_list(ipiv, k, w, m):
    """What K1 emits for the panel with pivots ipiv[k:k+w] (0-based absolute rows): (dst, src) pairs, new[dst] = old[src]."""
    cur = np.arange(m)
    for i in range(k, k + w):
        p = ipiv[i]
        cur[i], cur[p] = cur[p], cur[i]
    moved = np.nonzero(cur != np.arange(m))[0]
    return [(int(d), int(cur[d])) for d in moved]


def compose_and_apply(a, lists, k0, k1, cap):
    """lists: {panel start: (width, [(dst, src)...])}.  Returns the matrix after the node-level algorithm."""
    m = a.shape[0]
    a = a.copy()
    start = k0
    while start < k1:
        lim = min(k1, start + cap)
        starts = [c for c in range(start, lim) if c in lists and c + lists[c][0] <= lim]
        end = max(c + lists[c][0] for c in starts)
        assert sum(lists[c][0] for c in starts) == end - start and starts[0] == start
        cur = np.arange(m)                                 # cur[r] = original row now at r (rows >= start only matter)
        for c in starts:
            pairs = lists[c][1]
            vals = [cur[s] for (_, s) in pairs]            # all reads ...
            for (d, _), v in zip(pairs, vals):             # ... before all writes
                cur[d] = v
        n1 = end - start
        srcmap = cur[start:end]
        clist = [(r, int(cur[r])) for r in range(end, m) if cur[r] != r]
        assert all(start <= s < end for _, s in clist), "below <- below move"
        assert len(clist) <= n1
        xs = a[start:end].copy()                           # staged block
        new_block = np.where((srcmap < end)[:, None], xs[np.clip(srcmap - start, 0, n1 - 1)], a[np.clip(srcmap, 0, m - 1)])
        a[start:end] = new_block
        for d, s in clist:
            a[d] = xs[s - start]
        start = end
    return a


@pytest.mark.parametrize("m,k0,np_,cap,leaf,seed", [(300, 0, 128, 8192, 64, 0), (500, 40, 256, 96, 32, 1), (257, 0, 257, 64, 16, 2),
                                                    (1000, 100, 512, 200, 64, 3), (64, 0, 64, 8192, 8, 4)])
def test_net_permutation_equals_sequential_swaps(m, k0, np_, cap, leaf, seed):
    rng = np.random.default_rng(seed)
    k1 = k0 + np_
    ipiv = np.arange(m)
    for i in range(k0, k1):
        r = rng.random()
        ipiv[i] = i if r < 0.15 else (int(rng.integers(i, min(m, i + 3))) if r < 0.3 else int(rng.integers(i, m)))
    lists, c = {}, k0
    while c < k1:
        w = min(leaf, k1 - c) if rng.random() < 0.8 else min(max(1, leaf // 2), k1 - c)
        lists[c] = (w, panel_list(ipiv, c, w, m))
        c += w
    a = rng.random((m, 5))
    want = a.copy()
    for i in range(k0, k1):
        p = ipiv[i]
        if p != i:
            want[[i, p]] = want[[p, i]]
    got = compose_and_apply(a, lists, k0, k1, cap)
    assert np.array_equal(got, want)
