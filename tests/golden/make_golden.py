"""Generate the committed golden fixtures by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference (YyzHarry/imbalanced-regression @ a6fdc45) is imported from
/root/reference/agedb-dir with one shim: `fds.FDS` calls
`.cuda()` unconditionally (fds.py:52), so `torch.Tensor.cuda` is patched to
the identity on this CPU-only host.  Nothing is copied from the reference;
only its *outputs* on seeded inputs are stored (as .npz), together with the
label columns of its two age CSVs (real label histograms).
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/agedb-dir"
torch.Tensor.cuda = lambda self, *a, **k: self        # shim (2) of SURVEY §8c
sys.path.insert(0, REF)
import fds as ref_fds            # noqa: E402
import loss as ref_loss          # noqa: E402
import utils as ref_utils        # noqa: E402
import datasets as ref_datasets  # noqa: E402
import resnet as ref_resnet      # noqa: E402
import pandas as pd              # noqa: E402


def det_param(name: str, shape, scale: float) -> torch.Tensor:
    """Deterministic parameter generator shared with tests (see tests/util.py)."""
    seed = sum((i + 1) * ord(c) for i, c in enumerate(name)) % (2 ** 31)
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def windows():
    out = {}
    for kernel in ("gaussian", "triang", "laplace"):
        for ks, sigma in ((5, 2), (9, 1), (5, 1), (3, 2), (9, 2)):
            out[f"fds_{kernel}_{ks}_{sigma}"] = ref_fds.FDS._get_kernel_window(kernel, ks, sigma).numpy()
            out[f"lds_{kernel}_{ks}_{sigma}"] = np.asarray(ref_utils.get_lds_kernel_window(kernel, ks, sigma), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "windows.npz"), **out)


class _DS(ref_datasets.AgeDB):
    def __init__(self, df):      # bypass image plumbing; only _prepare_weights is used
        self.df = df


def lds():
    out = {}
    for tag, csv in (("agedb", "/root/reference/agedb-dir/data/agedb.csv"),
                     ("imdb_wiki", "/root/reference/imdb-wiki-dir/data/imdb_wiki.csv")):
        df = pd.read_csv(csv)
        df = df[df["split"] == "train"]
        labels = df["age"].values
        out[f"{tag}_labels"] = labels.astype(np.uint8 if labels.max() < 256 else np.int32)
        ds = _DS(df)
        for rw in ("sqrt_inv", "inverse"):
            for lds_on, k, ks, sg in ((False, "gaussian", 5, 2), (True, "gaussian", 5, 2),
                                      (True, "gaussian", 9, 1), (True, "triang", 9, 1),
                                      (True, "laplace", 5, 2)):
                if tag == "imdb_wiki" and (k != "gaussian" or ks != 5):
                    continue
                w = ds._prepare_weights(reweight=rw, lds=lds_on, lds_kernel=k, lds_ks=ks, lds_sigma=sg)
                out[f"{tag}_w_{rw}_{int(lds_on)}_{k}_{ks}_{sg}"] = np.asarray(w, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "lds.npz"), **out)


def fds_buffers(m):
    return {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}


def fds_cases():
    out = {}
    rng = np.random.RandomState(0)
    cases = {
        # name: (bucket_num, bucket_start, kernel, ks, sigma, momentum, D, N, label_lo, label_hi)
        "a": (100, 3, "gaussian", 9, 1, 0.9, 24, 700, 0, 110),     # agedb defaults, labels spill both edges
        "b": (100, 0, "gaussian", 5, 2, 0.9, 16, 500, 0, 130),     # imdb-wiki defaults
        "c": (101, 0, "triang", 5, 2, None, 8, 400, 0, 100),       # momentum None
        "d": (20, 5, "laplace", 3, 2, 0.5, 8, 64, 7, 15),          # edges never present -> nothing folds
    }
    for name, (bn, bs, k, ks, sg, mom, D, N, lo, hi) in cases.items():
        m = ref_fds.FDS(D, bucket_num=bn, bucket_start=bs, start_update=0, start_smooth=1,
                        kernel=k, ks=ks, sigma=sg, momentum=mom)
        out[f"{name}_cfg"] = np.array([bn, bs, ks, sg, -1.0 if mom is None else mom, D, N], dtype=np.float64)
        out[f"{name}_kernel"] = np.array(k)
        for ep in range(4):
            feats = np.maximum(rng.randn(N, D).astype(np.float32) + 0.5, 0)
            feats[:, 0] = 0.0                      # a dead (zero-variance) channel
            labels = rng.randint(lo, hi + 1, size=N).astype(np.float32)
            if name == "a" and ep == 2:
                labels[labels == 3] = 4            # lower edge value absent this epoch
            bx = np.maximum(rng.randn(48, D).astype(np.float32) + 0.5, 0)
            bl = rng.randint(lo, hi + 1, size=(48, 1)).astype(np.float32)
            out[f"{name}_e{ep}_feats"], out[f"{name}_e{ep}_labels"] = feats, labels
            out[f"{name}_e{ep}_bx"], out[f"{name}_e{ep}_bl"] = bx, bl
            # a training-step smooth() with this epoch's tables (in place on a copy)
            sm = m.smooth(torch.from_numpy(bx.copy()), torch.from_numpy(bl), ep)
            out[f"{name}_e{ep}_smooth"] = sm.numpy().copy()
            m.update_last_epoch_stats(ep)
            m.update_running_stats(torch.from_numpy(feats), torch.from_numpy(labels), ep)
            for kk, vv in fds_buffers(m).items():
                out[f"{name}_e{ep}_{kk}"] = vv
    np.savez_compressed(os.path.join(HERE, "fds.npz"), **out)


def calibrate_cases():
    out = {}
    rng = np.random.RandomState(1)
    x = rng.randn(6, 10).astype(np.float32)
    m1, m2 = rng.randn(10).astype(np.float32), rng.randn(10).astype(np.float32)
    v1 = (rng.rand(10).astype(np.float32) + 0.1)
    v2 = (rng.rand(10).astype(np.float32) * 3)
    v1z = v1.copy(); v1z[[1, 4]] = 0
    v1big = v1.copy(); v1big[2] = 1e-6            # ratio clamps at 10
    v2small = v2.copy(); v2small[3] = 1e-9        # ratio clamps at 0.1
    for tag, (a, b) in {"plain": (v1, v2), "zeros": (v1z, v2), "allzero": (np.zeros_like(v1), v2),
                        "clip_hi": (v1big, v2), "clip_lo": (v1, v2small)}.items():
        y = ref_utils.calibrate_mean_var(torch.from_numpy(x.copy()), torch.from_numpy(m1), torch.from_numpy(a),
                                         torch.from_numpy(m2), torch.from_numpy(b))
        out[f"{tag}_v1"], out[f"{tag}_v2"], out[f"{tag}_y"] = a, b, y.numpy()
    out["x"], out["m1"], out["m2"] = x, m1, m2
    np.savez_compressed(os.path.join(HERE, "calibrate.npz"), **out)


def loss_cases():
    out = {}
    rng = np.random.RandomState(2)
    x = (rng.randn(37, 1) * 8 + 40).astype(np.float32)
    t = rng.randint(0, 100, size=(37, 1)).astype(np.float32)
    x[3] = t[3]                                  # an exact zero residual
    x[5] = t[5] + 0.5                            # inside huber's quadratic zone
    w = (rng.rand(37, 1) * 3 + 0.2).astype(np.float32)
    out["x"], out["t"], out["w"] = x, t, w
    specs = [("mse", {}), ("l1", {}), ("huber", {}), ("huber", {"beta": 0.3}),
             ("focal_mse", {}), ("focal_l1", {}), ("focal_l1", {"activate": "tanh"}),
             ("focal_mse", {"activate": "tanh", "beta": 0.05, "gamma": 2}),
             ("focal_l1", {"beta": 20.0, "gamma": 1}), ("focal_l1", {"beta": 0.1, "gamma": 3})]
    for i, (kind, kw) in enumerate(specs):
        for use_w in (0, 1):
            xi = torch.from_numpy(x.copy()).requires_grad_(True)
            fn = getattr(ref_loss, f"weighted_{kind}_loss")
            l = fn(xi, torch.from_numpy(t), torch.from_numpy(w) if use_w else None, **kw)
            l.backward()
            out[f"case{i}_w{use_w}_loss"] = l.detach().numpy()
            out[f"case{i}_w{use_w}_grad"] = xi.grad.numpy()
        out[f"case{i}_kind"] = np.array(kind)
        out[f"case{i}_kw"] = np.array(repr(kw))
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


def resnet_case():
    """One fwd+bwd of the reference resnet50 (fp32, CPU, train mode, FDS idle)
    with deterministically generated parameters; stores outputs and a few
    gradient summaries -- pins oracle/resnet_ref.py."""
    torch.manual_seed(0)
    torch.set_num_threads(8)
    m = ref_resnet.resnet50(fds=True, bucket_num=100, bucket_start=3, start_update=0, start_smooth=1,
                            kernel="gaussian", ks=9, sigma=1, momentum=0.9)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 4:
                fan = p.shape[0] * p.shape[2] * p.shape[3]
                p.copy_(det_param(n, p.shape, (2.0 / fan) ** 0.5))
            elif n.endswith("linear.weight"):
                p.copy_(det_param(n, p.shape, 0.02))
            elif n.endswith("linear.bias"):
                p.fill_(0.1)
            elif n.endswith("weight"):        # BN gamma
                p.copy_(1.0 + 0.1 * det_param(n, p.shape, 1.0))
            else:                             # BN beta
                p.copy_(0.1 * det_param(n, p.shape, 1.0))
    m.train()
    x = det_param("input", (2, 3, 224, 224), 1.0)
    t = torch.tensor([[31.0], [44.0]])
    pred, enc = m(x, t, 0)
    loss = ref_loss.weighted_l1_loss(pred, t, torch.tensor([[0.5], [1.5]]))
    loss.backward()
    out = {"pred": pred.detach().numpy(), "enc": enc.detach().numpy(), "loss": loss.detach().numpy()}
    for n in ("conv1.weight", "bn1.weight", "bn1.bias", "layer1.0.conv2.weight", "layer2.0.downsample.0.weight",
              "layer3.5.conv3.weight", "layer4.2.bn3.weight", "layer4.2.conv1.weight", "linear.weight", "linear.bias"):
        g = dict(m.named_parameters())[n].grad
        out[f"grad_sum/{n}"] = g.double().sum().numpy()
        out[f"grad_abs/{n}"] = g.double().abs().sum().numpy()
        out[f"grad_head/{n}"] = g.reshape(-1)[:16].numpy().copy()
    out["bn1_running_mean"] = m.bn1.running_mean.numpy().copy()
    out["bn1_running_var"] = m.bn1.running_var.numpy().copy()
    out["n_params"] = np.array(sum(p.numel() for p in m.parameters()))
    np.savez_compressed(os.path.join(HERE, "resnet.npz"), **out)


if __name__ == "__main__":
    import logging
    logging.disable(logging.CRITICAL)
    windows(); lds(); fds_cases(); calibrate_cases(); loss_cases(); resnet_case()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
