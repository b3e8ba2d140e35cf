"""GPU parity of the whole native ResNet-50 path (runner + regressor + loss +
optimizer) against the torch fp32 oracle (oracle/resnet_ref.py, itself pinned
to the reference).  The conv stack stores bf16, so it is compared with the
oracle's quant=True mode (same storage points, straight-through rounding):
tolerances are a few bf16 ulps of the tensor scale, stated per check."""
import numpy as np
import pytest
import torch

from util import det_param

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_model(fds=False, **kw):
    from resnet import resnet50
    args = dict(fds=fds, bucket_num=100, bucket_start=3, start_update=0, start_smooth=1, kernel="gaussian",
                ks=9, sigma=1, momentum=0.9)
    args.update(kw)
    torch.manual_seed(0)
    m = resnet50(**args)
    with torch.no_grad():                       # non-trivial BN affine so its gradients are exercised
        for n, p in m.named_parameters():
            if p.dim() == 1 and ("bn" in n or "downsample.1" in n):
                p.copy_((1.0 if n.endswith("weight") else 0.0) + 0.1 * det_param(n, p.shape, 1.0))
    return m.to(DEV)


def oracle_params(m):
    return {n: p.detach().clone().requires_grad_(True) for n, p in m.named_parameters() if not n.startswith("FDS")}


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.reshape(1, -1).double(), b.reshape(1, -1).double()).item()


def test_state_dict_keys_match_reference_layout():
    from oracle import resnet_ref as R
    m = make_model(fds=True)
    names = [n for n, _ in m.named_parameters()]
    assert names == [n for n, _ in R.param_shapes()]
    assert [tuple(p.shape) for _, p in m.named_parameters()] == [s for _, s in R.param_shapes()]
    sd = m.state_dict()
    for k in ("bn1.running_mean", "bn1.running_var", "bn1.num_batches_tracked", "layer1.0.downsample.1.running_var",
              "layer4.2.bn3.num_batches_tracked", "FDS.running_mean", "FDS.smoothed_var_last_epoch", "linear.bias"):
        assert k in sd, k
    assert sum(p.numel() for p in m.parameters()) == 23510081
    assert sum(b.numel() for n, b in m.named_buffers() if not n.startswith("FDS")) == 53173
    # parameters tile one flat buffer; loading a state_dict keeps that
    m2 = make_model(fds=True)
    m2.load_state_dict(sd)
    assert torch.equal(m2.flat_parameters(), m.flat_parameters())
    assert m2.conv1.weight.data_ptr() == m2.flat_parameters().data_ptr()


@pytest.mark.parametrize("n,hw", [(16, 64), (4, 224)])
def test_forward_backward_vs_oracle(n, hw):
    from oracle import resnet_ref as R
    import loss as L
    m = make_model()
    m.train()
    p = oracle_params(m)
    x = det_param(f"x{n}", (n, 3, hw, hw), 1.0).to(DEV)
    t = (torch.arange(n, dtype=torch.float32, device=DEV).reshape(n, 1) * 5 + 10)
    w = torch.linspace(0.5, 1.5, n, device=DEV).reshape(n, 1)

    pred = m(x, t, 0)
    loss = L.weighted_l1_loss(pred, t, w)
    loss.backward()

    stats = {}
    rpred, renc = R.forward(p, x, stats=stats, quant=True)
    rloss = ((rpred - t).abs() * w).mean()
    rloss.backward()

    enc = m._run_forward(x, training=True)          # encoding (second pass, same batch statistics)
    assert rel(enc, renc.detach()) < 2e-2, rel(enc, renc.detach())
    assert cos(enc, renc.detach()) > 0.9995
    assert abs(loss.item() - rloss.item()) < 2e-2 * abs(rloss.item()) + 1e-3
    # BN running statistics of the first layer (momentum 0.1 from 0 / 1); two training forwards ran
    rm = 0.9 * stats["bn1.running_mean"] + stats["bn1.running_mean"]
    assert rel(m.bn1.running_mean, rm) < 2e-2
    assert int(m.bn1.num_batches_tracked) == 2
    # gradients: every parameter tensor, relative L2 error and direction
    named = dict(m.named_parameters())
    worst = 0.0
    for name, rp in p.items():
        g, rg = named[name].grad, rp.grad
        assert g is not None and torch.isfinite(g).all(), name
        e = rel(g, rg)
        worst = max(worst, e)
        assert cos(g, rg) > 0.98, (name, cos(g, rg), e)
    assert worst < 0.2, worst
    # aggregate over the flat buffer is much tighter than the worst tensor
    flat_ref = torch.cat([p[nm].grad.reshape(-1) for nm, _ in R.param_shapes()])
    assert rel(m.flat_grads(), flat_ref) < 5e-2, rel(m.flat_grads(), flat_ref)


def test_eval_mode_and_no_grad_paths():
    m = make_model(fds=True)
    x = det_param("xe", (8, 3, 64, 64), 1.0).to(DEV)
    t = torch.full((8, 1), 30.0, device=DEV)
    m.train()
    with torch.no_grad():
        out, feat = m(x, t, 0)                       # collection-pass shape: (pred, encoding)
    assert out.shape == (8, 1) and feat.shape == (8, 2048)
    m.eval()
    y1 = m(x)
    y2 = m(x)
    assert y1.shape == (8, 1) and torch.equal(y1, y2)  # eval: running statistics, deterministic
    assert int(m.bn1.num_batches_tracked) == 1


def test_train_step_fused_adam_matches_torch_adam():
    """3 steps of (forward, weighted L1, backward, Adam) with our fused optimizer vs torch.optim.Adam driven by
    the same gradients: parameters stay equal to fp32 rounding."""
    import loss as L
    from optim import FusedAdam
    m = make_model()
    m.train()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    shadow = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref_opt = torch.optim.Adam(shadow, lr=1e-3)
    x = det_param("xa", (8, 3, 64, 64), 1.0).to(DEV)
    t = torch.linspace(5, 80, 8, device=DEV).reshape(8, 1)
    losses = []
    for step in range(3):
        pred = m(x, t, 0)
        loss = L.weighted_l1_loss(pred, t, None)
        opt.zero_grad()
        loss.backward()
        for s, p in zip(shadow, m.parameters()):
            s.grad = p.grad.detach().clone()
        opt.step()
        ref_opt.step()
        losses.append(loss.item())
        for s, p in zip(shadow, m.parameters()):
            assert torch.allclose(s, p, rtol=1e-5, atol=1e-7)
    assert all(np.isfinite(losses))


def test_fds_model_training_step_smooth_active():
    """epoch >= start_smooth with non-trivial FDS tables: forward returns (pred, smoothed encoding) and the
    backward reaches the backbone through the calibration."""
    import loss as L
    m = make_model(fds=True, bucket_num=100, bucket_start=0, ks=5, sigma=2)
    m.train()
    torch.manual_seed(1)
    nb = 100
    m.FDS.running_mean_last_epoch = torch.randn(nb, 2048, device=DEV) * 0.1 + 0.5
    m.FDS.running_var_last_epoch = torch.rand(nb, 2048, device=DEV) + 0.5
    m.FDS.smoothed_mean_last_epoch = torch.randn(nb, 2048, device=DEV) * 0.1 + 0.5
    m.FDS.smoothed_var_last_epoch = torch.rand(nb, 2048, device=DEV) + 0.5
    x = det_param("xf", (8, 3, 64, 64), 1.0).to(DEV)
    t = torch.tensor([[3.], [17.], [17.], [40.], [99.], [120.], [0.], [55.]], device=DEV)
    pred, enc = m(x, t, 2)
    raw = m._run_forward(x, training=True)
    assert not torch.allclose(enc, raw)              # encoding returned is the smoothed one (in-place alias)
    L.weighted_l1_loss(pred, t, torch.ones_like(t)).backward()
    assert m.conv1.weight.grad.abs().sum() > 0 and torch.isfinite(m.flat_grads()).all()
