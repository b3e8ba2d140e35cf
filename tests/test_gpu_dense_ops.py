"""GPU parity of the NYUD2-DIR dense-prediction operators (nyud2-dir/models/modules.py) against torch fp32
on the same bf16-rounded operands: 5x5 / 3x3 convolutions with their gradients through the autograd wrappers, bilinear
up-sampling forward / backward, the channel concat, and one _UpProjection block composed of them."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("n,h,w,c,ho,wo", [(2, 8, 10, 64, 15, 19), (1, 29, 38, 128, 57, 76), (3, 7, 7, 16, 7, 7),
                                           (2, 12, 9, 8, 5, 4)], ids=["up_odd", "decoder_up", "identity", "down"])
def test_upsample_bilinear_fwd_bwd(n, h, w, c, ho, wo):
    import dense_ops as D
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, c, h, w, generator=g).to(DEV)
    xb = nhwc(x).requires_grad_(True)
    xr = nchw(xb.detach()).requires_grad_(True)
    out = D.upsample_bilinear(xb, (ho, wo))
    ref = F.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=False)
    assert rel(nchw(out.detach()), ref.detach()) < 3e-3
    dy = torch.randn(n, c, ho, wo, generator=g).to(DEV)
    dyb = nhwc(dy)
    out.backward(dyb)
    ref.backward(nchw(dyb))
    assert rel(nchw(xb.grad), xr.grad) < 3e-3


def test_cat_channels_fwd_bwd():
    import dense_ops as D
    g = torch.Generator().manual_seed(1)
    parts = [nhwc(torch.randn(2, c, 6, 5, generator=g).to(DEV)).requires_grad_(True) for c in (16, 16, 64, 8)]
    out = D.cat_channels(parts)
    assert torch.equal(out, torch.cat([p.detach() for p in parts], 3))
    dy = nhwc(torch.randn(2, 104, 6, 5, generator=g).to(DEV))
    out.backward(dy)
    off = 0
    for p in parts:
        c = p.shape[3]
        assert torch.equal(p.grad, dy[..., off:off + c])
        off += c


@pytest.mark.parametrize("cin,cout,k", [(128, 128, 5), (64, 16, 5), (128, 1, 5), (64, 64, 3)],
                         ids=["r_conv0", "mff_branch_16", "depth_head_1", "conv1_2"])
def test_conv2d_nhwc_autograd(cin, cout, k):
    import dense_ops as D
    g = torch.Generator().manual_seed(2)
    n, h, w = 2, 12, 16
    x = torch.randn(n, cin, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV).requires_grad_(True)
    xb = nhwc(x).requires_grad_(True)
    xr = nchw(xb.detach()).requires_grad_(True)
    wr = wt.detach().to(torch.bfloat16).float().requires_grad_(True)
    y = D.conv2d_nhwc(xb, wt, 1, k // 2)
    ref = F.conv2d(xr, wr, padding=k // 2)
    assert y.shape == (n, h, w, cout)
    assert rel(nchw(y.detach()), ref.detach()) < 4e-3
    dy = nhwc(torch.randn(n, cout, h, w, generator=g).to(DEV))
    y.backward(dy)
    ref.backward(nchw(dy))
    assert rel(nchw(xb.grad), xr.grad) < 4e-3
    assert rel(wt.grad, wr.grad) < 4e-3      # (cuDNN's own 5x5 wgrad is only good to ~3e-3 at the corner taps)


def test_up_projection_block_composed():
    """_UpProjection.forward (modules.py:23-32) without its BatchNorms: upsample -> (5x5 conv -> relu -> 3x3 conv) + 5x5
    conv -> relu, forward and input gradient against the same graph in torch."""
    import dense_ops as D
    g = torch.Generator().manual_seed(3)
    n, cin, cout, h, w, ho, wo = 2, 128, 64, 8, 10, 15, 19
    x = torch.randn(n, cin, h, w, generator=g).to(DEV)
    mk = lambda co, ci, k: (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5).to(DEV)
    w1, w12, w2 = mk(cout, cin, 5), mk(cout, cout, 3), mk(cout, cin, 5)
    xb = nhwc(x).requires_grad_(True)
    up = D.upsample_bilinear(xb, (ho, wo))
    b1 = D.conv2d_nhwc(torch.relu(D.conv2d_nhwc(up, w1, 1, 2)), w12, 1, 1)
    out = torch.relu(b1 + D.conv2d_nhwc(up, w2, 1, 2))
    q = lambda t: t.to(torch.bfloat16).float()
    xr = nchw(xb.detach()).requires_grad_(True)
    upr = q(F.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=False))
    r1 = q(F.conv2d(q(F.relu(q(F.conv2d(upr, q(w1), padding=2)))), q(w12), padding=1))
    ref = F.relu(r1 + q(F.conv2d(upr, q(w2), padding=2)))
    assert rel(nchw(out.detach()), ref.detach()) < 1e-2
    dy = nhwc(torch.randn(n, cout, ho, wo, generator=g).to(DEV))
    out.backward(dy)
    ref.backward(nchw(dy))
    assert rel(nchw(xb.grad), xr.grad) < 3e-2


@pytest.mark.parametrize("relu", [True, False], ids=["bn_relu", "bn_only"])
def test_batch_norm_train_fwd_bwd(relu):
    import dense_ops as D
    g = torch.Generator().manual_seed(4)
    n, c, h, w = 3, 128, 10, 12
    x = (torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3).to(DEV)
    xb = nhwc(x).requires_grad_(True)
    xr = nchw(xb.detach()).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(c, generator=g).to(DEV))
        bn.bias.copy_(0.1 * torch.randn(c, generator=g).to(DEV))
    gam = bn.weight.detach().clone().requires_grad_(True)
    bet = bn.bias.detach().clone().requires_grad_(True)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    out = D.batch_norm_train(xb, gam, bet, rm, rv, 0.1, 1e-5, relu)
    bn.train()
    ref = bn(xr)
    if relu:
        ref = F.relu(ref)
    assert rel(nchw(out.detach()), ref.detach()) < 4e-3
    assert rel(rm, bn.running_mean) < 1e-4 and rel(rv, bn.running_var) < 1e-4
    dy = nhwc(torch.randn(n, c, h, w, generator=g).to(DEV))
    out.backward(dy)
    ref.backward(nchw(dy))
    assert rel(nchw(xb.grad), xr.grad) < 1e-2
    assert rel(gam.grad, bn.weight.grad) < 1e-2 and rel(bet.grad, bn.bias.grad) < 1e-2


def test_pool_entry_points():
    import _lib
    g = torch.Generator().manual_seed(5)
    n, c, h, w = 2, 64, 12, 16
    x = torch.randn(n, c, h, w, generator=g).to(DEV)
    xb = nhwc(x)
    ho, wo = 6, 8
    out = torch.empty(n, ho, wo, c, dtype=torch.bfloat16, device=DEV)
    idx = torch.empty(n, ho, wo, c, dtype=torch.uint8, device=DEV)
    st = _lib.stream_ptr()
    _lib.call("dirb200_maxpool3x3s2_fwd", _lib.ptr(xb), n, h, w, c, _lib.ptr(out), _lib.ptr(idx), st)
    xr = nchw(xb).requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(nchw(out), ref.detach())
    dy = nhwc(torch.randn(n, c, ho, wo, generator=g).to(DEV))
    dx = torch.empty_like(xb)
    _lib.call("dirb200_maxpool3x3s2_bwd", _lib.ptr(dy), _lib.ptr(idx), n, h, w, c, _lib.ptr(dx), st)
    ref.backward(nchw(dy))
    assert rel(nchw(dx), xr.grad) < 4e-3
    enc = torch.empty(n, c, dtype=torch.float32, device=DEV)
    _lib.call("dirb200_avgpool_fwd", _lib.ptr(xb), n, h * w, c, _lib.ptr(enc), st)
    assert rel(enc, nchw(xb).mean(dim=(2, 3))) < 1e-5
    gx = torch.empty_like(xb)
    genc = torch.randn(n, c, generator=g).to(DEV)
    _lib.call("dirb200_avgpool_bwd", _lib.ptr(genc), n, h * w, c, _lib.ptr(gx), st)
    assert rel(nchw(gx), (genc / (h * w))[:, :, None, None].expand(n, c, h, w)) < 4e-3


def test_refinement_module_forward_backward():
    """Module R of nyud2-dir (modules.py:128-174) assembled from the operators: forward and all parameter gradients
    against the same torch graph on bf16-rounded storage points; state_dict keys as the reference's."""
    import dense_ops as D
    g = torch.Generator().manual_seed(6)
    n, c, h, w = 2, 128, 12, 16
    m = D.RefinementR(c).to(DEV)
    m.train()
    assert set(m.state_dict()) >= {"conv0.weight", "bn0.weight", "bn0.running_mean", "conv1.weight", "bn1.bias",
                                   "conv2.weight", "conv2.bias"}
    x = nhwc(torch.randn(n, c, h, w, generator=g).to(DEV)).requires_grad_(True)
    out = m(x)
    q = lambda t: t.to(torch.bfloat16).float()
    xr = nchw(x.detach()).requires_grad_(True)
    ref_m = torch.nn.Sequential()
    p = {k: v.detach().clone().requires_grad_(True) for k, v in m.named_parameters()}
    def bn(t, pre):
        return F.batch_norm(t, None, None, p[pre + ".weight"], p[pre + ".bias"], True, 0.1, 1e-5)
    r0 = q(F.relu(bn(q(F.conv2d(xr, q(p["conv0.weight"]), padding=2)), "bn0")))
    r1 = q(F.relu(bn(q(F.conv2d(r0, q(p["conv1.weight"]), padding=2)), "bn1")))
    ref = F.conv2d(r1, q(p["conv2.weight"]), padding=2) + p["conv2.bias"][None, :, None, None]
    assert rel(nchw(out.detach()), ref.detach()) < 2e-2
    dy = nhwc(torch.randn(n, 1, h, w, generator=g).to(DEV))
    out.backward(dy)
    ref.backward(nchw(dy))
    assert rel(nchw(x.grad), xr.grad) < 5e-2
    for k, v in m.named_parameters():
        assert v.grad is not None and rel(v.grad.float(), p[k].grad) < 5e-2, (k, rel(v.grad.float(), p[k].grad))
