"""Fused optimizers over the model's flat parameter / gradient buffers
(SURVEY.md §8 f-1): ONE kernel per step instead of torch.optim's 161-tensor
foreach chain.  Semantics follow torch.optim.Adam / torch.optim.SGD as used at
agedb-dir/train.py:163-164 (Adam: no weight decay; SGD: momentum, weight decay).

Works on any set of parameters that tile one contiguous flat fp32 buffer with
their .grad tiling another (resnet.ResNet guarantees both); otherwise raises --
there is no per-tensor fallback.
"""
import torch

import _lib
import resnet  # noqa: F401  (registers dirb200_adam_step / dirb200_sgd_step)


def _flat_span(tensors):
    """(data_ptr, numel) if the tensors tile one contiguous fp32 region in order, else None."""
    base = tensors[0].data_ptr()
    off = 0
    for t in tensors:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.data_ptr() != base + 4 * off:
            return None
        off += t.numel()
    return base, off


class _FlatOptimizer(torch.optim.Optimizer):
    def _span(self, group):
        ps = [p for p in group['params'] if p.requires_grad]
        if not ps:
            return None
        if any(p.grad is None for p in ps):
            raise _lib.Dirb200Error("fused optimizer: a parameter has no .grad (call backward first)")
        sp, sg = _flat_span(ps), _flat_span([p.grad for p in ps])
        if sp is None or sg is None or sp[1] != sg[1]:
            raise _lib.Dirb200Error("fused optimizer needs parameters (and grads) that tile one flat fp32 buffer; "
                                    "pass model.parameters() of a dirb200 resnet in order")
        return ps, sp[0], sg[0], sp[1]

    def _clip_coef(self, group, ps, g_ptr, n):
        """Device scalar min(1, max_norm / ||grad||) of torch.nn.utils.clip_grad_norm_ (sts-b-dir/trainer.py:147-149)
        when the group has `max_grad_norm`; the step kernel multiplies the gradients by it as it reads them."""
        mx = group.get('max_grad_norm')
        if not mx:
            return None
        st = self.state[ps[0]]
        if 'clip_ws' not in st:
            st['clip_ws'] = torch.zeros(_lib.raw("dirb200_grad_clip_workspace_bytes")(), dtype=torch.uint8,
                                        device=ps[0].device)
            st['clip_out'] = torch.zeros(2, dtype=torch.float32, device=ps[0].device)
        _lib.call("dirb200_grad_clip_coef", g_ptr, n, float(group['grad_scale']), float(mx), _lib.ptr(st['clip_ws']),
                  st['clip_ws'].numel(), _lib.ptr(st['clip_out']), _lib.stream_ptr())
        return st['clip_out']

    def last_grad_norm(self):
        """Norm seen by the most recent clipped step (device tensor, no sync) or None."""
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.requires_grad]
            if ps and 'clip_out' in self.state[ps[0]]:
                return self.state[ps[0]]['clip_out'][1]
        return None

    def zero_grad(self, set_to_none=False):
        """Keeps the flat gradient views attached and zeroes them (one memset per group)."""
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            span = _flat_span([p.grad for p in ps])
            if span is not None:
                flat = torch.as_strided(ps[0].grad, (span[1],), (1,))
                flat.zero_()
            else:
                for p in ps:
                    p.grad.zero_()


class FusedAdam(_FlatOptimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, grad_scale=1.0,
                 max_grad_norm=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale,
                                      max_grad_norm=max_grad_norm))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            span = self._span(group)
            if span is None:
                continue
            ps, p_ptr, g_ptr, n = span
            st = self.state[ps[0]]
            if 'exp_avg' not in st or st['exp_avg'].numel() != n:
                st['step'] = 0
                st['exp_avg'] = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
                st['exp_avg_sq'] = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
            st['step'] += 1
            b1, b2 = group['betas']
            clip = self._clip_coef(group, ps, g_ptr, n)
            _lib.call("dirb200_adam_step", p_ptr, g_ptr, _lib.ptr(st['exp_avg']), _lib.ptr(st['exp_avg_sq']), n,
                      float(group['lr']), float(b1), float(b2), float(group['eps']), float(group['weight_decay']),
                      int(st['step']), float(group['grad_scale']), _lib.ptr(clip), _lib.stream_ptr())


class FusedSGD(_FlatOptimizer):
    def __init__(self, params, lr, momentum=0, weight_decay=0, grad_scale=1.0, max_grad_norm=None):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, grad_scale=grad_scale,
                                      max_grad_norm=max_grad_norm))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            span = self._span(group)
            if span is None:
                continue
            ps, p_ptr, g_ptr, n = span
            st = self.state[ps[0]]
            first = 'momentum_buffer' not in st or st['momentum_buffer'].numel() != n
            if first:
                st['momentum_buffer'] = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
            clip = self._clip_coef(group, ps, g_ptr, n)
            _lib.call("dirb200_sgd_step", p_ptr, g_ptr, _lib.ptr(st['momentum_buffer']), n, float(group['lr']),
                      float(group['momentum']), float(group['weight_decay']), int(first), float(group['grad_scale']),
                      _lib.ptr(clip), _lib.stream_ptr())
