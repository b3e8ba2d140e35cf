"""The numpy oracle (oracle/dir_oracle.py) against fixtures produced by the
reference's own code (tests/golden/make_golden.py).  CPU only."""
import ast
import numpy as np
import pytest
from util import golden, assert_close
from oracle import dir_oracle as O


def test_windows_match_reference():
    g = golden("windows")
    for key in g.files:
        which, kernel, ks, sigma = key.split("_")
        fn = O.fds_kernel_window if which == "fds" else O.lds_kernel_window
        w = fn(kernel, int(ks), int(sigma))
        assert w.dtype == g[key].dtype
        assert_close(w, g[key], rtol=2e-7 if which == "fds" else 1e-14, atol=0, what=key)


def test_survey_window_vectors():
    # SURVEY.md §4 golden vectors (recorded from the reference at survey time)
    assert_close(O.fds_kernel_window("gaussian", 5, 2),
                 [0.18625069, 0.20524769, 0.21700326, 0.20524769, 0.18625069], rtol=1e-6)
    assert_close(O.lds_kernel_window("gaussian", 5, 2),
                 [0.858285238, 0.945827649, 1, 0.945827649, 0.858285238], rtol=1e-8)
    assert_close(O.fds_kernel_window("laplace", 5, 2),
                 [0.12475479, 0.20568587, 0.3391187, 0.20568587, 0.12475479], rtol=1e-6)
    assert_close(O.lds_kernel_window("laplace", 9, 1)[:5],
                 [0.018315639, 0.049787068, 0.135335283, 0.367879441, 1], rtol=1e-8)


def test_lds_histogram_agedb_bit_exact():
    g = golden("lds")
    hist = O.lds_histogram(g["agedb_labels"])
    assert hist.sum() == 12208 and hist.max() == 353 and int(np.argmax(hist)) == 35
    assert list(hist[:30]) == [0, 1, 0, 4, 2, 2, 4, 5, 4, 2, 5, 8, 6, 8, 9, 15, 18, 28, 60, 50,
                               101, 103, 123, 152, 199, 224, 215, 244, 287, 253]
    assert (hist > 0).sum() == 100


@pytest.mark.parametrize("tag", ["agedb", "imdb_wiki"])
def test_lds_weights_match_reference(tag):
    g = golden("lds")
    labels = g[f"{tag}_labels"]
    n = 0
    for key in g.files:
        if not key.startswith(f"{tag}_w_"):
            continue
        rest = key[len(tag) + 3:]
        rw = "sqrt_inv" if rest.startswith("sqrt_inv") else "inverse"
        lds_on, k, ks, sg = rest[len(rw) + 1:].split("_")
        _, w = O.lds_weights(labels, rw, lds=bool(int(lds_on)), lds_kernel=k, lds_ks=int(ks), lds_sigma=int(sg))
        assert_close(w, g[key], rtol=2e-6, atol=0, what=key)
        assert abs(float(w.astype(np.float64).mean()) - 1) < 1e-5
        n += 1
    assert n >= 4


def test_survey_agedb_weight_vectors():
    g = golden("lds")
    labels = g["agedb_labels"]
    assert list(labels[:5]) == [31, 44, 34, 74, 62]
    _, w = O.lds_weights(labels, "inverse", lds=True)
    assert_close(w[:5], [0.46940538, 0.52160543, 0.43663308, 1.5401634, 0.8378155], rtol=2e-6)
    _, w = O.lds_weights(labels, "sqrt_inv")
    assert_close(w[:5], [0.81728405, 0.87962353, 0.7524706, 1.4806272, 0.98012847], rtol=2e-6)


def test_calibrate_matches_reference():
    g = golden("calibrate")
    for tag in ("plain", "zeros", "allzero", "clip_hi", "clip_lo"):
        y = O.calibrate_mean_var(g["x"], g["m1"], g[f"{tag}_v1"], g["m2"], g[f"{tag}_v2"])
        assert_close(y, g[f"{tag}_y"], rtol=1e-6, atol=1e-6, what=tag)


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_fds_state_machine_matches_reference(name):
    g = golden("fds")
    bn, bs, ks, sg, mom, D, N = g[f"{name}_cfg"]
    st = O.FDSState(int(D), int(bn), int(bs), 0, 1, str(g[f"{name}_kernel"]), int(ks), int(sg),
                    None if mom < 0 else float(mom))
    for ep in range(4):
        sm = st.smooth(g[f"{name}_e{ep}_bx"], g[f"{name}_e{ep}_bl"], ep)
        assert_close(sm, g[f"{name}_e{ep}_smooth"], rtol=1e-5, atol=1e-5, what=f"smooth e{ep}")
        st.update_last_epoch_stats(ep)
        st.update_running_stats(g[f"{name}_e{ep}_feats"], g[f"{name}_e{ep}_labels"], ep)
        for k in ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                  "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked"):
            assert_close(getattr(st, k), g[f"{name}_e{ep}_{k}"], rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")
        assert st.epoch == int(g[f"{name}_e{ep}_epoch"][0])


def test_fds_bin_index_edge_rules():
    # edge value present -> out-of-range labels fold; absent -> dropped (fds.py:92-99)
    b = O.fds_bin_index([1, 3, 50, 99, 120], 100, 3)
    assert list(b) == [0, 0, 47, 96, 96]
    b = O.fds_bin_index([1, 4, 50, 98, 120], 100, 3)
    assert list(b) == [-1, 1, 47, 95, -1]


def test_losses_match_reference():
    g = golden("loss")
    i = 0
    while f"case{i}_kind" in g.files:
        kind = str(g[f"case{i}_kind"])
        kw = ast.literal_eval(str(g[f"case{i}_kw"]))
        for use_w in (0, 1):
            l, gr = O.weighted_loss(kind, g["x"], g["t"], g["w"] if use_w else None, **kw)
            assert_close(l, g[f"case{i}_w{use_w}_loss"], rtol=2e-6, what=f"{kind}{kw} loss")
            assert_close(gr, g[f"case{i}_w{use_w}_grad"], rtol=2e-5, atol=1e-7, what=f"{kind}{kw} grad")
        i += 1
    assert i == 10


def test_resnet_ref_matches_reference():
    """oracle/resnet_ref.py (fp32) vs the reference resnet50 fwd+bwd fixture."""
    import torch
    from util import det_param
    from oracle import resnet_ref as R
    g = golden("resnet")
    shapes = R.param_shapes()
    assert sum(int(np.prod(s)) for _, s in shapes) == int(g["n_params"]) == 23510081
    p = {}
    for n, s in shapes:
        if len(s) == 4:
            v = det_param(n, s, (2.0 / (s[0] * s[2] * s[3])) ** 0.5)
        elif n == "linear.weight":
            v = det_param(n, s, 0.02)
        elif n == "linear.bias":
            v = torch.full(s, 0.1)
        elif n.endswith("weight"):
            v = 1.0 + 0.1 * det_param(n, s, 1.0)
        else:
            v = 0.1 * det_param(n, s, 1.0)
        p[n] = v.requires_grad_(True)
    torch.set_num_threads(8)
    x = det_param("input", (2, 3, 224, 224), 1.0)
    t = torch.tensor([[31.0], [44.0]])
    stats = {}
    pred, enc = R.forward(p, x, stats=stats)
    loss = ((pred - t).abs() * torch.tensor([[0.5], [1.5]])).mean()
    loss.backward()
    assert_close(pred.detach().numpy(), g["pred"], rtol=2e-4, atol=1e-4, what="pred")
    assert_close(enc.detach().numpy(), g["enc"], rtol=2e-3, atol=2e-4, what="enc")
    assert_close(loss.item(), g["loss"], rtol=1e-4)
    assert_close(stats["bn1.running_mean"].numpy(), g["bn1_running_mean"], rtol=1e-4, atol=1e-6)
    assert_close(stats["bn1.running_var"].numpy(), g["bn1_running_var"], rtol=1e-4, atol=1e-6)
    for key in g.files:
        if key.startswith("grad_abs/"):
            n = key.split("/", 1)[1]
            got = p[n].grad.double().abs().sum().item()
            assert abs(got - float(g[key])) <= 3e-2 * float(g[key]) + 1e-6, (n, got, float(g[key]))  # batch-2 BN: fp32 noise flips ReLU masks


def test_fds_variant_nyud2_matches_reference():
    g = golden("fds_nyud2")
    bn, bs, ks, sg, mom, C, B, H, W = g["cfg"]
    st = O.FDSVariantState("nyud2", int(C), int(bn), int(bs), kernel="gaussian", ks=int(ks), sigma=int(sg),
                           momentum=float(mom))
    for ep in range(4):
        sm = st.smooth(g[f"e{ep}_bx"], g[f"e{ep}_bd"], ep)
        assert_close(sm, g[f"e{ep}_smooth"], rtol=1e-5, atol=1e-5, what=f"smooth e{ep}")
        st.update_last_epoch_stats(ep)
        st.update_running_stats(g[f"e{ep}_feats"], g[f"e{ep}_depth"], ep)
        for k in ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                  "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked"):
            assert_close(getattr(st, k), g[f"e{ep}_{k}"], rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")


def test_fds_variant_stsb_matches_reference():
    g = golden("fds_stsb")
    bn, bs, ks, sg, mom, D, N = g["cfg"]
    assert np.array_equal(O.bin_index_edges5(g["sweep"], int(bn), int(bs)), g["sweep_bucket"] - int(bs))
    st = O.FDSVariantState("stsb", int(D), int(bn), int(bs), kernel="gaussian", ks=int(ks), sigma=int(sg),
                           momentum=float(mom))
    for ep in range(4):
        sm = st.smooth(g[f"e{ep}_bx"], g[f"e{ep}_bl"], ep)
        assert_close(sm, g[f"e{ep}_smooth"], rtol=1e-5, atol=1e-5, what=f"smooth e{ep}")
        st.update_last_epoch_stats(ep)
        st.update_running_stats(g[f"e{ep}_feats"], g[f"e{ep}_labels"], ep)
        for k in ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                  "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked"):
            assert_close(getattr(st, k), g[f"e{ep}_{k}"], rtol=1e-5, atol=1e-6, what=f"{k} e{ep}")


def test_nyud2_bucket_weights_host_matches_reference():
    """datasets.depth_bucket_weights (host code of the product, numpy/scipy like the reference) vs loaddata.py."""
    import datasets
    g = golden("lds_nyud2")
    for rw in ("inverse", "sqrt_inv"):
        for lds_on in (0, 1):
            bw = datasets.depth_bucket_weights([int(v) for v in g["train_bucket_num"]], rw, lds=bool(lds_on))
            assert_close(bw, g[f"bw_{rw}_{lds_on}"], rtol=1e-6, atol=0, what=f"{rw} {lds_on}")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_shot_metrics_match_reference(tag):
    # fixture: the reference's own shot_metrics on the AgeDB test labels (tests/golden/make_golden_metrics.py)
    g = golden("metrics")
    sd = O.shot_metrics(g[f"{tag}_preds"], g[f"{tag}_labels"], g["train_labels"])
    got = [[sd[k][m] for m in ("mse", "l1", "gmean")] for k in ("many", "median", "low")]
    assert_close(got, g[f"{tag}_ref"], rtol=1e-6, atol=1e-12, what=f"shot metrics {tag}")
    assert_close([sd["overall"][m] for m in ("mse", "l1", "gmean")], g[f"{tag}_overall"], rtol=1e-6, atol=1e-12)
    assert sd["many"]["count"] + sd["median"]["count"] + sd["low"]["count"] == sd["overall"]["count"] == 2140


STSB_LDS = {"inv": dict(reweight="inverse"), "sqrt": dict(reweight="sqrt_inv"),
            "inv_lds_gau_5_2": dict(reweight="inverse", lds=True, lds_kernel="gaussian", lds_ks=5, lds_sigma=2),
            "sqrt_lds_gau_5_2": dict(reweight="sqrt_inv", lds=True, lds_kernel="gaussian", lds_ks=5, lds_sigma=2),
            "sqrt_lds_lap_9_1": dict(reweight="sqrt_inv", lds=True, lds_kernel="laplace", lds_ks=9, lds_sigma=1),
            "inv_lds_tri_5": dict(reweight="inverse", lds=True, lds_kernel="triang", lds_ks=5, lds_sigma=2)}


@pytest.mark.parametrize("tag", sorted(STSB_LDS))
def test_stsb_lds_weights_match_reference(tag):
    # fixture: the reference's own weighting lines (sts-b-dir/tasks.py:44-73) on the real STS-B training scores
    g = golden("lds_stsb")
    hist, w = O.stsb_lds_weights(g["scores"], **STSB_LDS[tag])
    assert np.array_equal(hist, g["hist"])                                   # integer work: bit-exact
    assert_close(w, g[f"w_{tag}"], rtol=2e-6, atol=0, what=tag)


@pytest.mark.parametrize("tag", sorted(STSB_LDS))
def test_stsb_host_mirror_from_oracle_bins(tag):
    # the host half of datasets.stsb_prepare_weights (everything except the GPU binning kernel), fed with the
    # oracle's bucket indices
    import datasets
    g = golden("lds_stsb")
    bins = O.bin_index_edges5(g["scores"], 50, 0)
    w, hist = datasets.stsb_weights_from_bins(bins, **STSB_LDS[tag])
    assert np.array_equal(hist, g["hist"])
    assert_close(w, g[f"w_{tag}"], rtol=2e-6, atol=0, what=tag)


def test_fds_stats_from_bins_equals_per_bin_loop():
    """The vectorised statistics used by the BASELINE-size GPU tests == the per-bin loop restatement of fds.py:100-102."""
    rng = np.random.RandomState(5)
    for (bn, bs, n, d) in ((100, 3, 3000, 40), (101, 0, 500, 700), (50, 10, 64, 5)):
        labels = rng.randint(0, 125, size=n).astype(np.float32)
        feats = np.maximum(rng.randn(n, d).astype(np.float32) * 0.7 + 2.0, 0)
        cnt, mean, var = O.fds_batch_stats(feats, labels, bn, bs)
        cnt2, mean2, var2 = O.fds_stats_from_bins(feats, O.fds_bin_index(labels, bn, bs), bn - bs)
        assert np.array_equal(cnt, cnt2)
        np.testing.assert_allclose(mean2, mean, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(var2, var, rtol=1e-5, atol=1e-7)
