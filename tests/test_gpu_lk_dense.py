"""GPU parity of the tiled dense-grid tracker (k_lk_dense, config C4): bit-identical to the thread-per-feature kernel
(k_lk_track_small, itself checked against the oracle in test_gpu_lk.py) on the same points, for every supported
window / level count / criteria, odd frame sizes, batches, flags; the automatic switch inside lkf_pyrlk_sparse
(device-side grid check) and the oracle on a small frame."""
import numpy as np
import pytest

from conftest import processed_pair

pytestmark = pytest.mark.gpu


def _grid(W, H):
    ys, xs = np.mgrid[0:H, 0:W]
    return np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)


def _pyramids(prev, nxt, win, max_level):
    import torch
    from lk_optical_flow_method_b200 import ops
    H, W = prev.shape[-2:]
    nl = ops.effective_levels(H, W, win, max_level)
    p = torch.from_numpy(np.ascontiguousarray(prev)).cuda()
    n = torch.from_numpy(np.ascontiguousarray(nxt)).cuda()
    if p.dim() == 2:
        p, n = p[None], n[None]
    return ops.build_pyramid(p, nl), ops.build_pyramid(n, nl)


@pytest.mark.parametrize("W,H,win,ml,crit,flags", [
    (200, 120, 5, 2, (3, 10, 0.03), 0),
    (203, 117, 5, 2, (3, 10, 0.03), 0),        # partial tiles on both edges
    (203, 117, 3, 2, (3, 10, 0.03), 0),
    (203, 117, 7, 2, (3, 10, 0.03), 0),
    (331, 75, 5, 3, (3, 30, 0.01), 0),         # four levels, cv2 default criteria
    (160, 96, 5, 0, (3, 10, 0.03), 0),         # a single level
    (160, 96, 5, 1, (1, 1, 0.0), 0),           # "pure stencil" variant: one iteration
    (160, 96, 5, 2, (1, 0, 0.0), 0),           # no iteration at all
    (203, 117, 5, 2, (3, 10, 0.03), 8),        # OPTFLOW_LK_GET_MIN_EIGENVALS
])
def test_dense_kernel_equals_small_kernel(W, H, win, ml, crit, flags):
    import torch
    from lk_optical_flow_method_b200 import ops
    prev, nxt = processed_pair(W, H)
    ppyr, npyr = _pyramids(prev, nxt, (win, win), ml)
    params = ops.lk_params_struct((win, win), ml, crit, flags)
    dn, ds, de, di = [t.cpu().numpy() for t in ops.pyrlk_dense(ppyr, npyr, params, want_iters=True)]
    # the same points in reversed order are not "the grid": lkf_pyrlk_sparse runs k_lk_track_small on them
    g = _grid(W, H)[::-1].copy()
    sn, ss, se, si = [t.cpu().numpy() for t in ops.pyrlk(ppyr, npyr, torch.from_numpy(g).cuda()[None], params, want_iters=True)]
    assert np.array_equal(dn[0], sn[0][::-1]) and np.array_equal(ds[0], ss[0][::-1])
    assert np.array_equal(de[0].view(np.uint32), se[0][::-1].view(np.uint32)) and np.array_equal(di[0], si[0][::-1])
    assert (ds == 1).any()


def test_sparse_entry_switches_to_dense_on_the_grid():
    """lkf_pyrlk_sparse handed the raster grid runs k_lk_dense (flag decided on the device); one moved point and the
    thread-per-feature kernel runs instead -- both give the per-point results of the explicit dense entry."""
    import torch
    from lk_optical_flow_method_b200 import ops
    W, H = 203, 117
    prev, nxt = processed_pair(W, H)
    ppyr, npyr = _pyramids(prev, nxt, (5, 5), 2)
    params = ops.lk_params_struct((5, 5), 2, (3, 10, 0.03))
    dense = [t.cpu().numpy() for t in ops.pyrlk_dense(ppyr, npyr, params)]
    g = torch.from_numpy(_grid(W, H)).cuda()[None].contiguous()
    auto = [t.cpu().numpy() for t in ops.pyrlk(ppyr, npyr, g, params)]
    for a, b in zip(dense, auto):
        assert np.array_equal(a, b)
    g2 = g.clone()
    g2[0, 1000, 0] += 0.25
    moved = [t.cpu().numpy() for t in ops.pyrlk(ppyr, npyr, g2, params)]
    keep = np.ones(W * H, bool)
    keep[1000] = False
    for a, b in zip(dense, moved):
        assert np.array_equal(a[0][keep], b[0][keep])
    assert not np.array_equal(dense[0][0][1000], moved[0][0][1000])


def test_dense_batch_and_oracle(oracle):
    """two frame pairs in one launch; every pixel of the small frames against the oracle"""
    from lk_optical_flow_method_b200 import ops
    W, H = 168, 100
    a0, a1 = processed_pair(W, H)
    b0, b1 = processed_pair(W, H, seed=1, t0=7)
    ppyr, npyr = _pyramids(np.stack([a0, b0]), np.stack([a1, b1]), (5, 5), 2)
    params = ops.lk_params_struct((5, 5), 2, (3, 10, 0.03))
    nxt, st, err = [t.cpu().numpy() for t in ops.pyrlk_dense(ppyr, npyr, params)]
    pts = _grid(W, H).reshape(-1, 1, 2)
    for k, (p, n) in enumerate(((a0, a1), (b0, b1))):
        q1, sq, eq = oracle.pyrlk(p, n, pts, (5, 5), 2, (3, 10, 0.03))
        assert (st[k].reshape(-1, 1) == sq).mean() >= 0.995
        both = ((st[k].reshape(-1, 1) == 1) & (sq == 1)).ravel()
        assert np.abs(nxt[k] - q1.reshape(-1, 2))[both].max() <= 1e-2
        assert np.abs(err[k].reshape(-1, 1) - eq)[both].max() <= 1e-2


def test_dense_rejects_what_it_does_not_cover():
    from lk_optical_flow_method_b200 import ops
    from lk_optical_flow_method_b200._lib import LkfError
    prev, nxt = processed_pair(160, 96)
    ppyr, npyr = _pyramids(prev, nxt, (9, 9), 2)
    with pytest.raises(LkfError):
        ops.pyrlk_dense(ppyr, npyr, ops.lk_params_struct((9, 9), 2, (3, 10, 0.03)))
    ppyr, npyr = _pyramids(prev, nxt, (5, 5), 2)
    with pytest.raises(LkfError):
        ops.pyrlk_dense(ppyr, npyr, ops.lk_params_struct((5, 5), 2, (3, 10, 0.03), 4))      # OPTFLOW_USE_INITIAL_FLOW


@pytest.mark.parametrize("seed", range(12))
def test_dense_kernel_random_sweep(seed):
    """seeded sweep of frame sizes (tiny, odd, not multiples of the 32 x 8 tile), windows, level counts, criteria and batch
    sizes: the tiled kernel against the thread-per-feature kernel on the same grid points (reversed order), bit for bit"""
    import torch
    from lk_optical_flow_method_b200 import ops
    rng = np.random.default_rng(1000 + seed)
    W = int(rng.choice([17, 33, 64, 95, 130, 257])) + int(rng.integers(0, 4))
    H = int(rng.choice([9, 16, 31, 50, 77])) + int(rng.integers(0, 4))
    win = int(rng.choice([3, 5, 7]))
    ml = int(rng.integers(0, 4))
    crit = [(3, 10, 0.03), (3, 30, 0.01), (1, 3, 0.0), (2, 0, 0.2), (3, 1, 0.5)][int(rng.integers(0, 5))]
    B = int(rng.integers(1, 4))
    flags = 8 if rng.random() < 0.25 else 0
    prevs, nxts = [], []
    for b in range(B):
        p, n = processed_pair(W, H, seed=int(rng.integers(0, 3)), t0=int(rng.integers(1, 9)))
        prevs.append(p); nxts.append(n)
    ppyr, npyr = _pyramids(np.stack(prevs), np.stack(nxts), (win, win), ml)
    params = ops.lk_params_struct((win, win), ml, crit, flags)
    dn, ds, de, di = [t.cpu().numpy() for t in ops.pyrlk_dense(ppyr, npyr, params, want_iters=True)]
    g = torch.from_numpy(_grid(W, H)[::-1].copy()).cuda()[None].repeat(B, 1, 1).contiguous()
    sn, ss, se, si = [t.cpu().numpy() for t in ops.pyrlk(ppyr, npyr, g, params, want_iters=True)]
    for b in range(B):
        assert np.array_equal(dn[b], sn[b][::-1]), (W, H, win, ml, crit, B)
        assert np.array_equal(ds[b], ss[b][::-1]) and np.array_equal(di[b], si[b][::-1])
        assert np.array_equal(de[b].view(np.uint32), se[b][::-1].view(np.uint32))
