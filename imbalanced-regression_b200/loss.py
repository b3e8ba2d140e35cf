"""loss.py -- drop-in mirror of agedb-dir/loss.py (and sts-b-dir/loss.py's
defaults via keyword arguments): the five weighted regression losses, each ONE
fused forward+backward kernel of libdirb200 (dirb200_loss_fwd_bwd) instead of
4-8 elementwise launches plus a reduction.

Signatures, defaults and the function names (looked up by name at
agedb-dir/train.py:255) are the reference's.
"""
import torch

import _lib

_WS = {}


def _workspace(device):
    ws = _WS.get(device)
    if ws is None:
        ws = torch.empty(int(_lib.raw("dirb200_loss_workspace_bytes")(0)), dtype=torch.uint8, device=device)
        _WS[device] = ws
    return ws


class _WeightedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, weights, kind, activate, beta, gamma):
        _lib.require_cuda(inputs, targets, weights)
        x = inputs.detach().to(torch.float32).contiguous()
        t = targets.detach().to(torch.float32).expand_as(x).contiguous()
        w = None if weights is None else weights.detach().to(torch.float32).expand_as(x).contiguous()
        n = x.numel()
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        need_grad = inputs.requires_grad
        grad = torch.empty_like(x) if need_grad else None
        ws = _workspace(x.device)
        _lib.call("dirb200_loss_fwd_bwd", _lib.LOSS_KINDS[kind], _lib.ptr(x), _lib.ptr(t), _lib.ptr(w), n,
                  float(beta), float(gamma), _lib.ACTIVATE[activate], 1.0, _lib.ptr(loss), _lib.ptr(grad),
                  _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None


def _run(kind, inputs, targets, weights, activate='sigmoid', beta=0., gamma=1.):
    return _WeightedLossFn.apply(inputs, targets, weights, kind, activate, beta, gamma)


def weighted_mse_loss(inputs, targets, weights=None):
    return _run('mse', inputs, targets, weights)


def weighted_l1_loss(inputs, targets, weights=None):
    return _run('l1', inputs, targets, weights)


def weighted_focal_mse_loss(inputs, targets, weights=None, activate='sigmoid', beta=.2, gamma=1):
    return _run('focal_mse', inputs, targets, weights, 'tanh' if activate == 'tanh' else 'sigmoid', beta, gamma)


def weighted_focal_l1_loss(inputs, targets, weights=None, activate='sigmoid', beta=.2, gamma=1):
    return _run('focal_l1', inputs, targets, weights, 'tanh' if activate == 'tanh' else 'sigmoid', beta, gamma)


def weighted_huber_loss(inputs, targets, weights=None, beta=1.):
    return _run('huber', inputs, targets, weights, beta=beta)
