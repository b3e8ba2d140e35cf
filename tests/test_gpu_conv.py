"""GPU parity of the tcgen05 implicit-GEMM convolutions (fprop / dgrad / wgrad)
against torch fp32 conv (TF32 off) on the same bf16-rounded operands.  Tolerance: the
kernel accumulates in fp32 and rounds the result to bf16 once, so outputs must
match the fp32 reference to bf16 precision (rel 2^-8 of the tensor scale);
wgrad is fp32 end to end (1e-3 of scale for the long reductions)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False          # the reference convolutions must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False


def nhwc_bf16(t):      # NCHW fp32 -> NHWC bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def to_nchw_f32(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def ref_wgrad(xr, wshape, dyr, stride, pad):
    """fp32 weight gradient.  For filters larger than 3x3 cuDNN's wgrad picks transform-domain algorithms whose
    corner-tap entries are off by up to 3e-3 of the tensor scale (checked against float64: the kernel under test agreed
    to 1e-6, cuDNN did not), so those are computed as an explicit im2col GEMM instead."""
    if wshape[2] <= 3:
        return torch.nn.grad.conv2d_weight(xr, wshape, dyr, stride=stride, padding=pad)
    cout, cin, kh, kw = wshape
    dw = torch.zeros(cout, cin * kh * kw, device=xr.device)
    for b in range(xr.shape[0]):                       # per image: bounds the unfolded matrix
        cols = F.unfold(xr[b:b + 1], (kh, kw), padding=pad, stride=stride)[0]          # [cin*kh*kw, L]
        dw += dyr[b].reshape(cout, -1) @ cols.t()
    return dw.view(cout, cin, kh, kw)


def run_conv(n, h, w, cin, cout, k, stride, pad, seed=0, check_dgrad=True, device_rng=False):
    import _lib, _convlib  # noqa: F401
    g = torch.Generator(device=DEV if device_rng else "cpu").manual_seed(seed)   # device_rng: full-size tensors
    gdev = DEV if device_rng else "cpu"
    x = torch.randn(n, cin, h, w, generator=g, device=gdev).to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=g, device=gdev) / (cin * k * k) ** 0.5).to(DEV)
    xb = nhwc_bf16(x)
    xr = to_nchw_f32(xb)                                  # bf16-rounded x as fp32 NCHW
    wr = wt.to(torch.bfloat16).float()
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    st = _lib.stream_ptr()
    wf = torch.empty(cout, k, k, cin, dtype=torch.bfloat16, device=DEV)
    wd = torch.empty(cin, k, k, cout, dtype=torch.bfloat16, device=DEV)
    _lib.call("dirb200_conv_prep_weights", _lib.ptr(wt), cout, cin, k, k, 0, _lib.ptr(wf), _lib.ptr(wd), st)
    assert torch.equal(wf, wt.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    assert torch.equal(wd, wt.permute(1, 2, 3, 0).contiguous().to(torch.bfloat16))
    y = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    shape = (n, h, w, cin, cout, k, k, stride, pad)
    _lib.call("dirb200_conv_fprop", _lib.ptr(xb), _lib.ptr(wf), _lib.ptr(y), *shape, 0, st)
    ref = F.conv2d(xr, wr, stride=stride, padding=pad)
    err = (to_nchw_f32(y) - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2 ** -7 * scale + 1e-6, f"fprop err {err} scale {scale}"

    dy = torch.randn(n, cout, ho, wo, generator=g, device=gdev).to(DEV)
    dyb = nhwc_bf16(dy)
    dyr = to_nchw_f32(dyb)
    del dy, x
    if check_dgrad:
        dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
        _lib.call("dirb200_conv_dgrad", _lib.ptr(dyb), _lib.ptr(wd), _lib.ptr(dx), *shape, st)
        ref_dx = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, stride=stride, padding=pad)
        err = (to_nchw_f32(dx) - ref_dx).abs().max().item()
        scale = ref_dx.abs().max().item()
        assert err <= 2 ** -7 * scale + 1e-6, f"dgrad err {err} scale {scale}"

    nbytes = _lib.raw("dirb200_conv_wgrad_workspace_bytes")(*shape, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    dw = torch.full((cout, cin, k, k), float("nan"), dtype=torch.float32, device=DEV)
    _lib.call("dirb200_conv_wgrad", _lib.ptr(xb), _lib.ptr(dyb), _lib.ptr(dw), _lib.ptr(ws), nbytes, *shape, 0, 0, st)
    ref_dw = ref_wgrad(xr, wr.shape, dyr, stride, pad)
    err = (dw - ref_dw).abs().max().item()
    scale = ref_dw.abs().max().item()
    assert err <= 2e-3 * scale + 1e-5, f"wgrad err {err} scale {scale}"
    # accumulate mode
    _lib.call("dirb200_conv_wgrad", _lib.ptr(xb), _lib.ptr(dyb), _lib.ptr(dw), _lib.ptr(ws), nbytes, *shape, 0, 1, st)
    assert (dw - 2 * ref_dw).abs().max().item() <= 4e-3 * scale + 1e-5
    torch.cuda.synchronize()


@pytest.mark.parametrize("cfg", [
    # n, h, w, cin, cout, k, stride, pad      (every distinct conv form of ResNet-50, small spatial sizes)
    (2, 8, 8, 64, 64, 1, 1, 0),        # layer1 1x1
    (2, 8, 8, 64, 256, 1, 1, 0),       # expand 1x1, BN=128 tiles x2
    (2, 8, 8, 256, 64, 1, 1, 0),       # reduce 1x1, 4 k-blocks
    (2, 8, 8, 64, 64, 3, 1, 1),        # layer1 3x3
    (2, 8, 8, 128, 128, 3, 2, 1),      # strided 3x3
    (2, 8, 8, 256, 512, 1, 2, 0),      # strided 1x1 downsample
    (3, 7, 7, 512, 512, 3, 1, 1),      # ragged M (147 pixels), many k-blocks (72)
    (1, 14, 14, 1024, 256, 1, 1, 0),   # 16 k-blocks > ring depth
    (4, 7, 7, 512, 2048, 1, 1, 0),     # 16 n-tiles
    (5, 9, 11, 64, 128, 3, 2, 1),      # odd sizes, non-square
    (75, 14, 14, 256, 256, 3, 1, 1),   # CTA pairs + im2col TMA: 115 m-tiles (odd: last pair has an empty peer half)
    (64, 28, 28, 512, 256, 1, 1, 0),   # CTA pairs + tiled TMA: 392 m-tiles, 8 k-blocks
    (64, 14, 14, 256, 1024, 1, 1, 0),  # CTA pairs, 4 n-tiles per cluster walk, accumulator double buffering
    # patch-resident 3x3 form (64 -> 64, stride 1): tiles of r whole padded image rows, nine displaced descriptors
    (3, 12, 20, 64, 64, 3, 1, 1),      # r = 4 rows of 22 padded columns (88 of 128 tile rows live), non-square
    (5, 7, 9, 64, 64, 3, 1, 1),        # r = 7: one tile per image
    (4, 56, 56, 64, 64, 3, 1, 1),      # the layer1 shape: r = 2 rows of 58, 112 tiles on 112 CTAs
    (1, 6, 60, 64, 64, 3, 1, 1),       # widest supported row (62 padded columns): the last tap reads to the slot's end
    (2, 5, 100, 64, 64, 3, 1, 1),      # too wide for a patch slot: im2col form
    # 5x5 / stride 1 / pad 2: the NYUD2 decoder and refinement convolutions (nyud2-dir/models/modules.py:11-20,154-160)
    (2, 12, 16, 128, 128, 5, 1, 2),    # R.conv0 / conv1 form: 25 taps x 2 channel blocks = 50 k-blocks
    (3, 9, 11, 64, 128, 5, 1, 2),      # odd sizes
    (1, 24, 32, 128, 64, 5, 1, 2),     # up-projection form (Cout < Cin)
    (2, 8, 8, 256, 256, 5, 1, 2),      # CTA pairs with 100 k-blocks
])
def test_conv_forms(cfg):
    run_conv(*cfg)


def test_conv_layer1_full_batch_shape():
    # a BASELINE-size layer: batch 32 of the 56x56x64 3x3 (M = 100 352 rows, 784 tiles)
    run_conv(32, 56, 56, 64, 64, 3, 1, 1, seed=1)


def test_conv_nyud2_refinement_shape():
    # BASELINE config 4 geometry at batch 1: the 5x5 128 -> 128 conv of nyud2-dir's R module on a 240 x 320 map
    run_conv(1, 240, 320, 128, 128, 5, 1, 2, seed=2, device_rng=True)


def test_stem_conv_and_s2d():
    import _lib, _convlib  # noqa: F401
    g = torch.Generator(device="cpu").manual_seed(3)
    n, h, w, cout = 3, 32, 32, 64
    x = torch.randn(n, 3, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, 3, 7, 7, generator=g) / 147 ** 0.5).to(DEV)
    st = _lib.stream_ptr()
    xs = torch.empty(n, h // 2, w // 2, 16, dtype=torch.bfloat16, device=DEV)
    _lib.call("dirb200_input_to_s2d", _lib.ptr(x), n, h, w, _lib.ptr(xs), st)
    ref_s2d = torch.zeros(n, h // 2, w // 2, 16, device=DEV)
    for ph in range(2):
        for pw in range(2):
            for c in range(3):
                ref_s2d[..., (ph * 2 + pw) * 4 + c] = x[:, c, ph::2, pw::2]
    assert torch.equal(xs, ref_s2d.to(torch.bfloat16))
    wf = torch.empty(cout, 256, dtype=torch.bfloat16, device=DEV)
    _lib.call("dirb200_conv_prep_weights", _lib.ptr(wt), cout, 3, 7, 7, 1, _lib.ptr(wf), None, st)
    shape = (n, h, w, 3, cout, 7, 7, 2, 3)
    y = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.call("dirb200_conv_fprop", _lib.ptr(xs), _lib.ptr(wf), _lib.ptr(y), *shape, 1, st)
    xr, wr = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    ref = F.conv2d(xr, wr, stride=2, padding=3)
    err = (to_nchw_f32(y) - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item() + 1e-6, err
    dy = torch.randn(n, cout, h // 2, w // 2, generator=g).to(DEV)
    dyb = nhwc_bf16(dy)
    nbytes = _lib.raw("dirb200_conv_wgrad_workspace_bytes")(*shape, 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    dw = torch.full((cout, 3, 7, 7), float("nan"), dtype=torch.float32, device=DEV)
    _lib.call("dirb200_conv_wgrad", _lib.ptr(xs), _lib.ptr(dyb), _lib.ptr(dw), _lib.ptr(ws), nbytes, *shape, 1, 0, st)
    ref_dw = torch.nn.grad.conv2d_weight(xr, wr.shape, to_nchw_f32(dyb), stride=2, padding=3)
    err = (dw - ref_dw).abs().max().item()
    assert err <= 2e-3 * ref_dw.abs().max().item() + 1e-5, err
