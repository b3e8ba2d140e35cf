"""GPU parity of the evaluation reductions (dirb200_int_label_histogram / dirb200_shot_metrics through the
train.py mirror) against the reference-generated fixture and the numpy oracle.  Counts and the histogram are
integers (bit-exact); the error sums are fp64 on both sides (1e-6 relative, the float32 difference is shared)."""
import numpy as np
import pytest
import torch
from util import golden, assert_close
from oracle import dir_oracle as O

pytestmark = pytest.mark.gpu


def _rows(sd):
    return [[sd[k][m] for m in ("mse", "l1", "gmean")] for k in ("many", "median", "low")]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_shot_metrics_vs_reference_golden(tag):
    import train as T
    g = golden("metrics")
    sd = T.shot_metrics(torch.from_numpy(g[f"{tag}_preds"]).cuda(), torch.from_numpy(g[f"{tag}_labels"]).cuda(),
                        g["train_labels"])
    assert_close(_rows(sd), g[f"{tag}_ref"], rtol=1e-6, atol=1e-12, what=f"shot metrics {tag}")
    assert_close([sd["overall"][m] for m in ("mse", "l1", "gmean")], g[f"{tag}_overall"], rtol=1e-6, atol=1e-12)


def test_shot_metrics_large_random_vs_oracle_and_exact_counts():
    import _lib
    rng = np.random.RandomState(5)
    train = np.clip(np.round(rng.gamma(6.0, 6.5, size=191509)), 0, 140).astype(np.float32)   # IMDB-WIKI-sized column
    labels = np.clip(np.round(rng.gamma(6.0, 6.5, size=50000)), 0, 150).astype(np.float32)
    labels[::501] += 0.5                                   # non-integer label values: training count 0 -> low shot
    preds = (labels + rng.randn(labels.size) * 5).astype(np.float32)
    dev = torch.device("cuda")
    t, l, p = (torch.from_numpy(a).to(dev) for a in (train, labels, preds))
    nbins = 160
    hist = torch.zeros(nbins, dtype=torch.int64, device=dev)
    _lib.call("dirb200_int_label_histogram", _lib.ptr(t), t.numel(), nbins, _lib.ptr(hist), _lib.stream_ptr())
    assert np.array_equal(hist.cpu().numpy(), np.bincount(train.astype(int), minlength=nbins))   # bit-exact
    out = torch.empty(4, 4, dtype=torch.float64, device=dev)
    _lib.call("dirb200_shot_metrics", _lib.ptr(p), _lib.ptr(l), p.numel(), _lib.ptr(hist), nbins, 100, 20,
              _lib.ptr(out), _lib.stream_ptr())
    o = out.cpu().numpy()
    want = O.shot_metrics(preds, labels, train)
    for row, name in enumerate(("overall", "many", "median", "low")):
        assert int(o[row, 0]) == want[name]["count"], name                                       # bit-exact
        assert_close([o[row, 1] / o[row, 0], o[row, 2] / o[row, 0], np.exp(o[row, 3] / o[row, 0])],
                     [want[name][m] for m in ("mse", "l1", "gmean")], rtol=1e-6, atol=1e-12, what=name)
    assert int(o[0, 0]) == labels.size and int(o[1, 0] + o[2, 0] + o[3, 0]) == labels.size


def test_validate_mirror_reports_overall_and_shots():
    import train as T

    class Const(torch.nn.Module):
        def forward(self, x):
            return x[:, :1] * 0 + 30.0

    labels = torch.tensor([[25.], [30.], [41.], [30.]])
    loader = [(torch.zeros(2, 3), labels[:2], None), (torch.zeros(2, 3), labels[2:], None)]
    mse, l1, gm = T.validate(loader, Const().cuda(), train_labels=np.asarray([30] * 150 + [25] * 50 + [41] * 3))
    assert abs(mse - (25 + 0 + 121 + 0) / 4) < 1e-9 and abs(l1 - (5 + 0 + 11 + 0) / 4) < 1e-9 and gm == 0.0
