"""GPU parity: FDS / LDS / loss kernels (through the C ABI via the host mirrors)
against the numpy oracle and the reference-generated golden fixtures.
Tolerances: bins / histograms bit-exact; fp32 statistics and losses 1e-5 rel."""
import ast
import numpy as np
import pytest
import torch

from util import golden, assert_close
from oracle import dir_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).to(DEV)


# ------------------------------------------------------------------- FDS
@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_fds_state_machine_vs_reference_golden(name):
    from fds import FDS
    g = golden("fds")
    bn, bs, ks, sg, mom, D, N = g[f"{name}_cfg"]
    m = FDS(int(D), int(bn), int(bs), 0, 1, str(g[f"{name}_kernel"]), int(ks), int(sg),
            None if mom < 0 else float(mom)).to(DEV)
    for ep in range(4):
        sm = m.smooth(T(g[f"{name}_e{ep}_bx"]), T(g[f"{name}_e{ep}_bl"]), ep)
        assert_close(sm.cpu().numpy(), g[f"{name}_e{ep}_smooth"], rtol=1e-5, atol=1e-5, what=f"smooth e{ep}")
        m.update_last_epoch_stats(ep)
        m.update_running_stats(T(g[f"{name}_e{ep}_feats"]), T(g[f"{name}_e{ep}_labels"]), ep)
        sd = m.state_dict()
        for k in ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                  "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked", "epoch"):
            assert_close(sd[k].cpu().numpy(), g[f"{name}_e{ep}_{k}"], rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")


def test_fds_state_dict_keys_and_alias():
    from fds import FDS
    m = FDS(8, 100, 3).to(DEV)
    assert list(m.state_dict().keys()) == ['epoch', 'running_mean', 'running_var', 'running_mean_last_epoch',
                                           'running_var_last_epoch', 'smoothed_mean_last_epoch',
                                           'smoothed_var_last_epoch', 'num_samples_tracked']
    assert m.running_mean.shape == (97, 8)
    m.update_last_epoch_stats(1)
    assert m.running_mean_last_epoch.data_ptr() == m.running_mean.data_ptr()   # fds.py:55 alias


@pytest.mark.parametrize("bn,bs,n,d", [(100, 3, 1000, 2048), (101, 0, 256, 2048), (100, 0, 5000, 36), (50, 10, 77, 7)])
def test_fds_bins_and_stats_vs_oracle(bn, bs, n, d):
    import _lib
    rng = np.random.RandomState(n + d)
    labels = rng.randint(0, 130, size=n).astype(np.float32)
    feats = np.maximum(rng.randn(n, d).astype(np.float32) * 0.7 + 3.0, 0)   # mean >> std
    feats[:, d // 2] = 0
    from fds import FDS
    m = FDS(d, bn, bs, momentum=0.9).to(DEV)
    m.begin_epoch_stats(T(labels))
    # bins bit-exact
    bins = torch.empty(n, dtype=torch.int32, device=DEV)
    _lib.call("dirb200_fds_bin_rows", _lib.ptr(T(labels)), n, bn, bs, 0, _lib.ptr(m._acc["flags"]), _lib.ptr(bins),
              _lib.stream_ptr())
    assert np.array_equal(bins.cpu().numpy(), O.fds_bin_index(labels, bn, bs))
    m.accumulate_batch(T(feats), T(labels))
    cnt, mean, var = O.fds_batch_stats(feats, labels, bn, bs)
    assert np.array_equal(m._acc["counts"].cpu().numpy(), cnt)
    m.finish_epoch_stats(0)
    has = cnt > 0
    assert_close(m.running_mean.cpu().numpy()[has], mean[has], rtol=1e-5, atol=1e-7, what="mean")
    assert_close(m.running_var.cpu().numpy()[has], var[has], rtol=1e-5, atol=1e-7, what="var")
    assert (m.running_var.cpu().numpy()[~has] == 1).all() and (m.running_mean.cpu().numpy()[~has] == 0).all()
    assert (m.running_var[:, d // 2].cpu().numpy()[has] == 0).all()       # dead channel -> exactly 0


def test_fds_streamed_equals_one_shot_full_size():
    """BASELINE config-2 size (N = 12 208 x 2048, 101 bins): streamed batches of
    256 == one call; counts add up; constant column has zero variance."""
    from fds import FDS
    torch.manual_seed(0)
    n, d = 12208, 2048
    labels = torch.randint(0, 101, (n,), device=DEV).float()
    feats = torch.relu(torch.randn(n, d, device=DEV) + 0.5)
    feats[:, 5] = 2.5
    a = FDS(d, 101, 0).to(DEV)
    b = FDS(d, 101, 0).to(DEV)
    a.update_running_stats(feats, labels, 0)
    b.begin_epoch_stats(labels)
    for i in range(0, n, 256):
        b.accumulate_batch(feats[i:i + 256], labels[i:i + 256])
    assert int(b._acc["counts"].sum()) == n
    b.finish_epoch_stats(0)
    assert_close(a.running_mean.cpu().numpy(), b.running_mean.cpu().numpy(), rtol=1e-6, atol=1e-7)
    assert_close(a.running_var.cpu().numpy(), b.running_var.cpu().numpy(), rtol=1e-6, atol=1e-7)
    assert float(a.running_var[:, 5].abs().max()) < 1e-6
    # against torch on device (independent implementation), one bin
    rows = feats[labels == 40]
    assert_close(a.running_mean[40].cpu().numpy(), rows.mean(0).cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert_close(a.running_var[40].cpu().numpy(), rows.var(0, unbiased=True).cpu().numpy(), rtol=2e-5, atol=1e-6)


def test_fds_calibrate_forward_backward_vs_oracle():
    from fds import FDS
    rng = np.random.RandomState(3)
    d, bn, bs = 64, 100, 3
    m = FDS(d, bn, bs).to(DEV)
    nb = bn - bs
    m1, m2 = rng.randn(nb, d).astype(np.float32), rng.randn(nb, d).astype(np.float32)
    v1 = (rng.rand(nb, d).astype(np.float32) + 0.05)
    v2 = (rng.rand(nb, d).astype(np.float32) * 4)
    v1[:, 7] = 0                      # dead channel everywhere
    v1[10] = 0                        # a whole row with sum(v1) < 1e-10 -> identity
    v1[20, 3] = 1e-7                  # clamp at 10
    v2[21, 4] = 1e-9                  # clamp at 0.1
    m.running_mean_last_epoch, m.running_var_last_epoch = T(m1), T(v1)
    m.smoothed_mean_last_epoch, m.smoothed_var_last_epoch = T(m2), T(v2)
    labels = np.concatenate([np.arange(0, 120), [13, 23, 24, 24, 99, 3]]).astype(np.float32)
    x = rng.randn(len(labels), d).astype(np.float32)
    xt = T(x).requires_grad_(True)
    y = m.smooth(xt * 1.0, T(labels).reshape(-1, 1), 5)
    ref = O.fds_calibrate(x, labels, bn, bs, m1, v1, m2, v2)
    assert_close(y.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-6, what="calibrate fwd")
    gy = rng.randn(*x.shape).astype(np.float32)
    y.backward(T(gy))
    assert_close(xt.grad.cpu().numpy(), gy * O.fds_calibrate_scale(labels, bn, bs, v1, v2), rtol=1e-5, atol=1e-7,
                 what="calibrate bwd")
    # edge value absent from the batch -> out-of-range rows untouched
    lab2 = np.array([1, 2, 50, 120], dtype=np.float32)
    x2 = rng.randn(4, d).astype(np.float32)
    y2 = m.smooth(T(x2), T(lab2).reshape(-1, 1), 5)
    assert_close(y2.cpu().numpy(), O.fds_calibrate(x2, lab2, bn, bs, m1, v1, m2, v2), rtol=1e-5, atol=1e-6)
    assert np.array_equal(y2.cpu().numpy()[[0, 1, 3]], x2[[0, 1, 3]])


def test_calibrate_mean_var_function_vs_golden():
    from utils import calibrate_mean_var
    g = golden("calibrate")
    for tag in ("plain", "zeros", "allzero", "clip_hi", "clip_lo"):
        y = calibrate_mean_var(T(g["x"]), T(g["m1"]), T(g[f"{tag}_v1"]), T(g["m2"]), T(g[f"{tag}_v2"]))
        assert_close(y.cpu().numpy(), g[f"{tag}_y"], rtol=1e-5, atol=1e-6, what=tag)


def test_smooth_tables_vs_oracle():
    from fds import FDS
    rng = np.random.RandomState(5)
    for kernel, ks, sg, nb in (("gaussian", 5, 2, 97), ("gaussian", 9, 1, 100), ("triang", 9, 1, 101), ("laplace", 3, 2, 5)):
        m = FDS(2048, nb, 0, kernel=kernel, ks=ks, sigma=sg).to(DEV)
        tab = rng.rand(nb, 2048).astype(np.float32)
        out = m._smooth_table(T(tab))
        assert_close(out.cpu().numpy(), O.smooth_bins(tab, O.fds_kernel_window(kernel, ks, sg)), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ losses
def test_losses_vs_reference_golden():
    import loss as L
    g = golden("loss")
    i = 0
    while f"case{i}_kind" in g.files:
        kind = str(g[f"case{i}_kind"])
        kw = ast.literal_eval(str(g[f"case{i}_kw"]))
        for use_w in (0, 1):
            x = T(g["x"]).requires_grad_(True)
            l = getattr(L, f"weighted_{kind}_loss")(x, T(g["t"]), T(g["w"]) if use_w else None, **kw)
            l.backward()
            assert_close(l.item(), g[f"case{i}_w{use_w}_loss"], rtol=1e-5, what=f"{kind}{kw} loss")
            assert_close(x.grad.cpu().numpy(), g[f"case{i}_w{use_w}_grad"], rtol=2e-5, atol=1e-7, what=f"{kind}{kw} grad")
        i += 1
    assert i == 10


@pytest.mark.parametrize("n", [1, 256, 100003, 32 * 240 * 320])
def test_loss_sizes_vs_oracle(n):
    import loss as L
    rng = np.random.RandomState(n % 1000)
    x = (rng.randn(n, 1) * 5 + 30).astype(np.float32)
    t = rng.randint(0, 100, size=(n, 1)).astype(np.float32)
    w = (rng.rand(n, 1) + 0.5).astype(np.float32)
    for kind in ("l1", "mse", "focal_l1", "huber"):
        xt = T(x).requires_grad_(True)
        l = getattr(L, f"weighted_{kind}_loss")(xt, T(t), T(w))
        (l * 2).backward()
        lo, go = O.weighted_loss(kind, x, t, w)
        assert_close(l.item(), lo, rtol=1e-5)
        assert_close(xt.grad.cpu().numpy(), 2 * go, rtol=2e-5, atol=1e-9 / max(1, n) ** 0)


# --------------------------------------------------------------------- LDS
@pytest.mark.parametrize("tag", ["agedb", "imdb_wiki"])
def test_lds_weights_vs_reference_golden(tag):
    from datasets import lds_prepare_weights
    g = golden("lds")
    labels = g[f"{tag}_labels"]
    n = 0
    for key in g.files:
        if not key.startswith(f"{tag}_w_"):
            continue
        rest = key[len(tag) + 3:]
        rw = "sqrt_inv" if rest.startswith("sqrt_inv") else "inverse"
        lds_on, k, ks, sg = rest[len(rw) + 1:].split("_")
        w, hist = lds_prepare_weights(labels, rw, lds=bool(int(lds_on)), lds_kernel=k, lds_ks=int(ks),
                                      lds_sigma=int(sg), return_hist=True)
        assert np.array_equal(hist.cpu().numpy(), O.lds_histogram(labels))         # bit exact
        assert_close(w.cpu().numpy(), g[key], rtol=2e-6, atol=0, what=key)
        n += 1
    assert n >= 4


@pytest.mark.parametrize("world", [2, 8])
def test_lds_weights_of_shards_equal_unsharded(world):
    """A label column split rank::world: histograms summed (the int64 all-reduce of datasets.lds_prepare_weights with
    sharded=True) + the sharded weights entry give every sample bit-for-bit the weight of the unsharded call."""
    import _lib
    from datasets import lds_prepare_weights
    from utils import get_lds_kernel_window
    labels = golden("lds")["imdb_wiki_labels"].astype(np.float32)
    for rw, lds_on in (("sqrt_inv", True), ("inverse", True), ("inverse", False)):
        whole = lds_prepare_weights(labels, rw, lds=lds_on, lds_kernel="gaussian", lds_ks=5, lds_sigma=2).cpu().numpy()
        hist = torch.zeros(121, dtype=torch.int64, device=DEV)
        shards = [T(labels[r::world]) for r in range(world)]
        for sh in shards:                    # == SUM all-reduce of the per-rank histograms
            _lib.call("dirb200_lds_histogram", _lib.ptr(sh), sh.numel(), 121, _lib.ptr(hist), _lib.stream_ptr())
        assert np.array_equal(hist.cpu().numpy(), O.lds_histogram(labels))
        window = np.ascontiguousarray(get_lds_kernel_window("gaussian", 5, 2), dtype=np.float64) if lds_on else None
        for r, sh in enumerate(shards):
            out = torch.empty_like(sh)
            scratch = torch.empty(2 * 121 + 2, dtype=torch.float64, device=DEV)
            _lib.call("dirb200_lds_weights_sharded", _lib.ptr(sh), sh.numel(), len(labels), 121, _lib.REWEIGHT[rw],
                      None if window is None else window.ctypes.data_as(_lib.P), 0 if window is None else len(window),
                      _lib.ptr(hist), _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr())
            assert np.array_equal(out.cpu().numpy(), whole[r::world]), (rw, lds_on, r)


def test_lds_histogram_clamps_and_accumulates():
    import _lib
    lab = T([0, 0.9, 1, 119.5, 120, 121, 186, 500])
    hist = torch.zeros(121, dtype=torch.int64, device=DEV)
    for _ in range(2):
        _lib.call("dirb200_lds_histogram", _lib.ptr(lab), lab.numel(), 121, _lib.ptr(hist), _lib.stream_ptr())
    h = hist.cpu().numpy()
    assert h[0] == 4 and h[1] == 2 and h[119] == 2 and h[120] == 8 and h.sum() == 16


# ------------------------------------------------- FDS variants (nyud2-dir, sts-b-dir)
def test_fds_variant_nyud2_vs_reference_golden():
    from fds_variants import FDSDepth
    g = golden("fds_nyud2")
    bn, bs, ks, sg, mom, C, B, H, W = g["cfg"]
    m = FDSDepth(int(C), int(bn), int(bs), 0, 1, "gaussian", int(ks), int(sg), float(mom)).to(DEV)
    for ep in range(4):
        sm = m.smooth(T(g[f"e{ep}_bx"]), T(g[f"e{ep}_bd"]), ep)
        assert_close(sm.cpu().numpy(), g[f"e{ep}_smooth"], rtol=1e-5, atol=1e-5, what=f"smooth e{ep}")
        m.update_last_epoch_stats(ep)
        m.update_running_stats(T(g[f"e{ep}_feats"]), T(g[f"e{ep}_depth"]), ep)
        sd = m.state_dict()
        for k in ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                  "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked", "epoch"):
            assert_close(sd[k].cpu().numpy(), g[f"e{ep}_{k}"], rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")


def test_fds_variant_stsb_vs_reference_golden():
    import _lib
    from fds_variants import FDSSTSB
    g = golden("fds_stsb")
    bn, bs, ks, sg, mom, D, N = g["cfg"]
    # bucket index of a label sweep incl. the float32 edges: bit exact
    sweep = T(g["sweep"])
    bins = torch.empty(sweep.numel(), dtype=torch.int32, device=DEV)
    flags = torch.zeros(2, dtype=torch.int32, device=DEV)
    _lib.call("dirb200_fds_bin_rows", _lib.ptr(sweep), sweep.numel(), int(bn), int(bs), _lib.BIN_EDGES5, _lib.ptr(flags),
              _lib.ptr(bins), _lib.stream_ptr())
    assert np.array_equal(bins.cpu().numpy(), g["sweep_bucket"] - int(bs))
    m = FDSSTSB(int(D), int(bn), int(bs), 0, 1, "gaussian", int(ks), int(sg), float(mom)).to(DEV)
    for ep in range(4):
        sm = m.smooth(T(g[f"e{ep}_bx"]), T(g[f"e{ep}_bl"]), ep)
        assert_close(sm.cpu().numpy(), g[f"e{ep}_smooth"], rtol=1e-5, atol=1e-5, what=f"smooth e{ep}")
        m.update_last_epoch_stats(ep)
        m.update_running_stats(T(g[f"e{ep}_feats"]), T(g[f"e{ep}_labels"]), ep)
        sd = m.state_dict()
        for k in ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                  "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked", "epoch"):
            assert_close(sd[k].cpu().numpy(), g[f"e{ep}_{k}"], rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")


def test_fds_depth_dense_size_vs_oracle():
    """BASELINE config-4-like dense map (reduced: 4 x 128 x 60 x 80 = 19 200 pixel rows): bins bit-exact, stats 1e-5."""
    import _lib
    from fds_variants import FDSDepth
    rng = np.random.RandomState(11)
    B, C, H, W = 4, 128, 60, 80
    feats = np.maximum(rng.randn(B, C, H, W).astype(np.float32) + 0.4, 0)
    depth = (rng.rand(B, 1, H, W).astype(np.float32) * 9.3 + 0.7)
    m = FDSDepth(C).to(DEV)
    m.update_running_stats(T(feats), T(depth), 0)
    ref = O.FDSVariantState("nyud2", C, 100, 7)
    ref.update_running_stats(feats, depth, 0)
    assert_close(m.running_mean.cpu().numpy(), ref.running_mean, rtol=1e-5, atol=1e-6)
    assert_close(m.running_var.cpu().numpy(), ref.running_var, rtol=1e-5, atol=1e-6)
    assert_close(m.num_samples_tracked.cpu().numpy(), ref.num_samples_tracked, rtol=0, atol=0)


def test_nyud2_pixel_weights_and_dense_loss_vs_reference():
    """Per-pixel LDS weights (table lookup, bit exact) and the dense weighted MSE of nyud2-dir/train.py:200."""
    import datasets
    import loss as L
    g = golden("lds_nyud2")
    depth = T(g["depth"])
    for rw in ("inverse", "sqrt_inv"):
        for lds_on in (0, 1):
            w = datasets.depth_pixel_weights(depth, g[f"bw_{rw}_{lds_on}"])
            assert np.array_equal(w.cpu().numpy(), g[f"w_{rw}_{lds_on}"]), (rw, lds_on)
    pred = (depth + torch.randn_like(depth) * 0.3).requires_grad_(True)
    w = datasets.depth_pixel_weights(depth, g["bw_inverse_1"])
    l = L.weighted_mse_loss(pred, depth, w)
    l.backward()
    ref = (((pred.detach() - depth) ** 2) * w).mean()
    assert_close(l.item(), ref.item(), rtol=1e-5)
    assert_close(pred.grad.cpu().numpy(), (2 * (pred.detach() - depth) * w / depth.numel()).cpu().numpy(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("tag,kw", [
    ("inv", dict(reweight="inverse")),
    ("sqrt_lds_gau_5_2", dict(reweight="sqrt_inv", lds=True, lds_kernel="gaussian", lds_ks=5, lds_sigma=2)),
    ("inv_lds_tri_5", dict(reweight="inverse", lds=True, lds_kernel="triang", lds_ks=5, lds_sigma=2)),
])
def test_stsb_lds_weights_vs_reference_golden(tag, kw):
    """sts-b-dir/tasks.py:44-73 on the real STS-B training scores: GPU bucket rule + host table vs the fixture made
    by the reference's own lines (tests/golden/make_golden_stsb_lds.py)."""
    from datasets import stsb_prepare_weights
    g = golden("lds_stsb")
    w = stsb_prepare_weights(g["scores"], **kw)
    assert w.is_cuda and w.dtype == torch.float32 and w.numel() == g["scores"].size
    assert_close(w.cpu().numpy(), g[f"w_{tag}"], rtol=2e-6, atol=0, what=tag)
    assert stsb_prepare_weights(g["scores"], "none") is None


# ------------------------------------------------- BASELINE configs 4 and 5 at their own size
def test_fds_depth_config4_full_size_vs_oracle():
    """BASELINE config 4: the refinement input [32, 128, 240, 320] -> 2 457 600 pixel rows x 128 channels, depth bins
    clamp(int(d*10), 7, 99).  Bins bit-exact, per-bin statistics within 1e-5 of the oracle, and the calibration of the
    whole map (FDS.smooth, 1.26 GB in place) against the oracle on a strided sample of rows."""
    import _lib
    from fds_variants import FDSDepth
    torch.manual_seed(4)
    B, C, H, W = 32, 128, 240, 320
    feats = torch.relu(torch.randn(B, C, H, W, device=DEV) + 0.4)
    depth = torch.rand(B, 1, H, W, device=DEV) * 9.3 + 0.7
    m = FDSDepth(C).to(DEV)
    lab = depth.reshape(-1)
    bins = torch.empty(lab.numel(), dtype=torch.int32, device=DEV)
    flags = torch.zeros(2, dtype=torch.int32, device=DEV)
    _lib.call("dirb200_fds_bin_rows", _lib.ptr(lab), lab.numel(), 100, 7, _lib.BIN_DEPTH10, _lib.ptr(flags), _lib.ptr(bins),
              _lib.stream_ptr())
    lab_np = lab.cpu().numpy()
    want_bins = O.bin_index_depth10(lab_np, 100, 7)
    assert np.array_equal(bins.cpu().numpy(), want_bins)
    m.update_running_stats(feats, depth, 0)
    rows = feats.permute(0, 2, 3, 1).reshape(-1, C).cpu().numpy()
    cnt, mean, var = O.fds_stats_from_bins(rows, want_bins, 93)
    assert np.array_equal(m.num_samples_tracked.cpu().numpy().astype(np.int64), cnt) and cnt.sum() == B * H * W
    has = cnt > 0
    assert_close(m.running_mean.cpu().numpy()[has], mean[has], rtol=1e-5, atol=1e-6, what="mean")
    assert_close(m.running_var.cpu().numpy()[has], var[has], rtol=1e-5, atol=1e-6, what="var")
    # second epoch: tables move, smoothing of the full map against the oracle on every 4 801st pixel row
    m.update_last_epoch_stats(1)
    m.update_running_stats(feats, depth, 1)
    sel = np.arange(0, B * H * W, 4801)
    before = rows[sel].copy()
    out = m.smooth(feats, depth, 1)
    got = out.permute(0, 2, 3, 1).reshape(-1, C)[torch.from_numpy(sel).to(DEV)].cpu().numpy()
    tabs = [getattr(m, k).cpu().numpy() for k in ("running_mean_last_epoch", "running_var_last_epoch",
                                                   "smoothed_mean_last_epoch", "smoothed_var_last_epoch")]
    want = before.copy()
    sb = want_bins[sel]
    for b in np.unique(sb):
        r = sb == b
        want[r] = O.calibrate_mean_var_v2(before[r], tabs[0][b], tabs[1][b], tabs[2][b], tabs[3][b], 0.2, 5.0)
    assert_close(got, want, rtol=1e-5, atol=1e-5, what="smooth")


def test_fds_stsb_config5_full_size_vs_oracle():
    """BASELINE config 5: feature_dim 12 000 (sts-b-dir/models.py:46-49), 50 score buckets over [0, 5], one epoch of
    the STS-B training set (5 749 sentence pairs) + a batch-128 smooth: statistics, empty-bucket fill and calibration
    against the oracle state machine."""
    from fds_variants import FDSSTSB
    rng = np.random.RandomState(9)
    N, D, Bs = 5749, 12000, 128
    feats = (rng.randn(N, D).astype(np.float32) * 0.5 + 0.3)
    scores = np.round(rng.beta(2.0, 1.5, size=N) * 5 * 4) / 4          # quarter-point scores as in STS-B
    scores = scores.astype(np.float32)
    m = FDSSTSB(D).to(DEV)
    ref = O.FDSVariantState("stsb", D, 50, 0)
    for ep in (0, 1):
        m.update_last_epoch_stats(ep)
        ref.update_last_epoch_stats(ep)
        m.update_running_stats(T(feats), T(scores), ep)
        ref.update_running_stats(feats, scores, ep)
        for k in ("running_mean", "running_var", "smoothed_mean_last_epoch", "smoothed_var_last_epoch",
                  "num_samples_tracked"):
            assert_close(getattr(m, k).cpu().numpy(), getattr(ref, k), rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")
    xb, lb = feats[:Bs].copy(), scores[:Bs]
    got = m.smooth(T(xb), T(lb).reshape(-1, 1), 1).cpu().numpy()
    want = ref.smooth(xb.copy(), lb, 1)
    assert_close(got, want, rtol=1e-5, atol=1e-5, what="smooth")
