"""dense_ops.py -- the operators of the NYUD2-DIR decoder / feature-fusion / refinement modules
(nyud2-dir/models/modules.py:6-174) on the B200 path, as autograd functions over NHWC bf16 tensors:

  conv2d_nhwc(x, weight, stride, padding)   nn.Conv2d(..., bias=False) with 1x1 / 3x3 / 5x5 filters (modules.py:11-20, 63,
                                            107, 134-141): tcgen05 implicit GEMM, forward + data / weight gradients
  upsample_bilinear(x, size)                F.upsample(x, size=size, mode='bilinear') (modules.py:24)
  cat_channels(tensors)                     torch.cat(tensors, 1) (modules.py:120)

Activations are channels-last bf16 ([N, H, W, C], C a multiple of 64 for the convolutions, of 8 elsewhere); weights stay
the reference's fp32 [Cout, Cin, KH, KW] parameters (state_dict compatible).  Convolutions with fewer than 64 output
channels (the 16-channel MFF branches, the final 1-channel depth conv) are run with the output channels zero-padded to 64.
No CPU path."""
import torch

import _lib
import _convlib  # noqa: F401  (registers the conv entry points)


def _shape(x, weight, stride, padding):
    n, h, w, cin = x.shape
    cout, cin_w, kh, kw = weight.shape
    assert cin == cin_w and kh == kw, (x.shape, weight.shape)
    return (n, h, w, cin, cout, kh, kw, stride, padding)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, padding):
        _lib.require_cuda(x, weight)
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and weight.dtype == torch.float32
        shape = _shape(x, weight, stride, padding)
        n, h, w, cin, cout, k, _, s, p = shape
        ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        st = _lib.stream_ptr()
        wf = torch.empty(cout, k, k, cin, dtype=torch.bfloat16, device=x.device)
        wd = torch.empty(cin, k, k, cout, dtype=torch.bfloat16, device=x.device)
        _lib.call("dirb200_conv_prep_weights", _lib.ptr(weight.contiguous()), cout, cin, k, k, 0, _lib.ptr(wf), _lib.ptr(wd), st)
        y = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=x.device)
        _lib.call("dirb200_conv_fprop", _lib.ptr(x), _lib.ptr(wf), _lib.ptr(y), *shape, 0, st)
        ctx.save_for_backward(x, wd)
        ctx.shape = shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wd = ctx.saved_tensors
        shape = ctx.shape
        n, h, w, cin, cout, k, _, s, p = shape
        dy = dy.contiguous()
        st = _lib.stream_ptr()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.call("dirb200_conv_dgrad", _lib.ptr(dy), _lib.ptr(wd), _lib.ptr(dx), *shape, st)
        if ctx.needs_input_grad[1]:
            nbytes = _lib.raw("dirb200_conv_wgrad_workspace_bytes")(*shape, 0)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            dw = torch.empty(cout, cin, k, k, dtype=torch.float32, device=x.device)
            _lib.call("dirb200_conv_wgrad", _lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(ws), nbytes, *shape, 0, 0, st)
        return dx, dw, None, None


def conv2d_nhwc(x, weight, stride=1, padding=0):
    """x bf16 [N, H, W, Cin] (Cin a multiple of 64), weight fp32 [Cout, Cin, K, K] -> bf16 [N, Ho, Wo, Cout]."""
    cout = weight.shape[0]
    if cout % 64 != 0:                       # narrow heads: zero-padded output channels, sliced off again
        pad = 64 - cout % 64
        wp = torch.cat([weight, weight.new_zeros(pad, *weight.shape[1:])], 0)
        return _ConvFn.apply(x, wp, stride, padding)[..., :cout]
    return _ConvFn.apply(x, weight, stride, padding)


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ho, wo):
        _lib.require_cuda(x)
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[3] % 8 == 0
        n, h, w, c = x.shape
        out = torch.empty(n, ho, wo, c, dtype=torch.bfloat16, device=x.device)
        _lib.call("dirb200_upsample_bilinear_fwd", _lib.ptr(x), n, h, w, c, ho, wo, _lib.ptr(out), _lib.stream_ptr())
        ctx.dims = (n, h, w, c, ho, wo)
        return out

    @staticmethod
    def backward(ctx, dy):
        n, h, w, c, ho, wo = ctx.dims
        dy = dy.contiguous()
        dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=dy.device)
        _lib.call("dirb200_upsample_bilinear_bwd", _lib.ptr(dy), n, h, w, c, ho, wo, _lib.ptr(dx), _lib.stream_ptr())
        return dx, None, None


def upsample_bilinear(x, size):
    """F.upsample(x, size=size, mode='bilinear') (align_corners=False) on an NHWC bf16 tensor."""
    return _UpsampleFn.apply(x, int(size[0]), int(size[1]))


class _CatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        n, h, w = xs[0].shape[:3]
        chans = [int(t.shape[3]) for t in xs]
        total = sum(chans)
        out = torch.empty(n, h, w, total, dtype=torch.bfloat16, device=xs[0].device)
        off = 0
        for t, c in zip(xs, chans):
            assert t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape[:3]) == (n, h, w)
            _lib.call("dirb200_copy_channels", _lib.ptr(t), c, 0, _lib.ptr(out), total, off, c, n * h * w, _lib.stream_ptr())
            off += c
        ctx.chans = chans
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        n, h, w, total = dy.shape
        outs, off = [], 0
        for c in ctx.chans:
            g = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=dy.device)
            _lib.call("dirb200_copy_channels", _lib.ptr(dy), total, off, _lib.ptr(g), c, 0, c, n * h * w, _lib.stream_ptr())
            outs.append(g)
            off += c
        return tuple(outs)


def cat_channels(tensors):
    """torch.cat(tensors, 1) for NHWC bf16 tensors (channel counts multiples of 8)."""
    return _CatFn.apply(*tensors)
