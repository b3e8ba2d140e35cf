"""Golden fixtures for the NYUD2 / STS-B FDS variants, produced by the reference's own modules.

    python tests/golden/make_golden_variants.py nyud2
    python tests/golden/make_golden_variants.py stsb

(one process per variant: both import a module called `util`.)  Shims: `Tensor.cuda` / `Tensor.cpu`
return a CLONE, which reproduces what the device hops of nyud2-dir/models/fds.py:88-96 do on a real GPU (new
tensors, so `running_*_last_epoch` keeps the values it had when it was bound); sts-b-dir/fds.py has no hops and
keeps the alias.
"""
import importlib.util
import logging
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
logging.disable(logging.CRITICAL)


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def buffers(m):
    return {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}


def nyud2():
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()
    torch.Tensor.cpu = lambda self, *a, **k: self.clone()
    sys.path.insert(0, "/root/reference/nyud2-dir")
    ref = load("/root/reference/nyud2-dir/models/fds.py", "ref_fds_nyud2")
    rng = np.random.RandomState(0)
    C, B, H, W = 16, 2, 6, 5
    m = ref.FDS(C, bucket_num=100, bucket_start=7, start_update=0, start_smooth=1, kernel="gaussian", ks=5, sigma=2,
                momentum=0.9)
    out = {"cfg": np.array([100, 7, 5, 2, 0.9, C, B, H, W], dtype=np.float64)}
    for ep in range(4):
        feats = np.maximum(rng.randn(B, C, H, W).astype(np.float32) + 0.3, 0)
        feats[:, 0] = 0.0
        depth = (rng.rand(B, 1, H, W).astype(np.float32) * 11.0)       # 0 .. 11 m: below bucket 7 and above 99
        bx = np.maximum(rng.randn(B, C, H, W).astype(np.float32) + 0.3, 0)
        bd = (rng.rand(B, 1, H, W).astype(np.float32) * 11.0)
        out[f"e{ep}_feats"], out[f"e{ep}_depth"], out[f"e{ep}_bx"], out[f"e{ep}_bd"] = feats, depth, bx, bd
        sm = m.smooth(torch.from_numpy(bx.copy()), torch.from_numpy(bd), ep)
        out[f"e{ep}_smooth"] = sm.contiguous().numpy().copy()
        m.update_last_epoch_stats(ep)
        m.update_running_stats(torch.from_numpy(feats), torch.from_numpy(depth), ep)
        for k, v in buffers(m).items():
            out[f"e{ep}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "fds_nyud2.npz"), **out)


def stsb():
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, "/root/reference/sts-b-dir")
    import warnings
    warnings.filterwarnings("ignore")
    ref = load("/root/reference/sts-b-dir/fds.py", "ref_fds_stsb")
    rng = np.random.RandomState(1)
    D, N = 24, 200
    m = ref.FDS(D, bucket_num=50, bucket_start=0, start_update=0, start_smooth=1, kernel="gaussian", ks=5, sigma=2,
                momentum=0.9)
    out = {"cfg": np.array([50, 0, 5, 2, 0.9, D, N], dtype=np.float64)}
    for ep in range(4):
        feats = (rng.randn(N, D).astype(np.float32) * 0.5 + 0.2)
        feats[:, 0] = 0.0
        # STS-B style scores: multiples of 0.2 plus some arbitrary values; leaves several buckets empty
        lab = np.round(rng.rand(N) * 25) / 5.0
        extra = rng.rand(20).astype(np.float32) * 5.0
        lab[:20] = extra
        lab[20] = 5.0
        lab[21] = 0.0
        lab = lab.astype(np.float32)
        lab[(lab > 1.0) & (lab < 1.7)] = 1.0                          # buckets 11..16 empty
        bx = (rng.randn(32, D).astype(np.float32) * 0.5 + 0.2)
        bl = (np.round(rng.rand(32, 1) * 25) / 5.0).astype(np.float32)
        out[f"e{ep}_feats"], out[f"e{ep}_labels"], out[f"e{ep}_bx"], out[f"e{ep}_bl"] = feats, lab, bx, bl
        sm = m.smooth(torch.from_numpy(bx.copy()), torch.from_numpy(bl), ep)
        out[f"e{ep}_smooth"] = sm.numpy().copy()
        m.update_last_epoch_stats(ep)
        m.update_running_stats(torch.from_numpy(feats), torch.from_numpy(lab), ep)
        for k, v in buffers(m).items():
            out[f"e{ep}_{k}"] = v
    # bucket index of a label sweep (every multiple of 0.01 and the float32 edges themselves)
    sweep = np.concatenate([np.arange(0, 501, dtype=np.float32) / np.float32(100.0),
                            np.linspace(0, 5, 51).astype(np.float32)])
    out["sweep"] = sweep
    out["sweep_bucket"] = np.array([m._get_bucket_idx(v) for v in sweep], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "fds_stsb.npz"), **out)


if __name__ == "__main__" and sys.argv[1] in ("nyud2", "stsb"):
    {"nyud2": nyud2, "stsb": stsb}[sys.argv[1]]()
    print("ok", sys.argv[1])


def nyud2_lds():
    """nyud2-dir/loaddata.py bucket weights + per-pixel lookup (run as: ... nyud2_lds)."""
    import types
    sys.path.insert(0, "/root/reference/nyud2-dir")
    for name in ("nyu_transform",):                # image transforms pull in accimage/PIL extras: not needed here
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nyu_transform"].__dict__.update({k: object for k in
        ("Scale", "RandomHorizontalFlip", "Rotate", "CenterCrop", "ToTensor", "Lighting", "ColorJitter", "Normalize")})
    # loaddata.py gets numpy / torch through `from nyu_transform import *`
    sys.modules["nyu_transform"].__dict__.update({"np": np, "torch": torch})
    ld = load("/root/reference/nyud2-dir/loaddata.py", "ref_loaddata")
    out = {"train_bucket_num": np.asarray(ld.TRAIN_BUCKET_NUM, dtype=np.float64)}
    ds = ld.depthDataset.__new__(ld.depthDataset)
    rng = np.random.RandomState(3)
    depth = torch.from_numpy((rng.rand(2, 1, 12, 9) * 10.5).astype(np.float32))
    out["depth"] = depth.numpy()
    for rw in ("inverse", "sqrt_inv"):
        for lds_on in (0, 1):
            args = types.SimpleNamespace(reweight=rw, lds=bool(lds_on), lds_kernel="gaussian", lds_ks=5, lds_sigma=2,
                                         bucket_num=100, bucket_start=7)
            bw = ds._get_bucket_weights(args)
            ds.bucket_weights = bw
            out[f"bw_{rw}_{lds_on}"] = np.asarray(bw, dtype=np.float32)
            out[f"w_{rw}_{lds_on}"] = ds._get_weights(depth).numpy()
    np.savez_compressed(os.path.join(HERE, "lds_nyud2.npz"), **out)


if len(sys.argv) > 1 and sys.argv[1] == "nyud2_lds":
    nyud2_lds()
    print("ok nyud2_lds")
