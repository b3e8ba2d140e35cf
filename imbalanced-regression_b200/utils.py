"""utils.py -- mirror of agedb-dir/utils.py for the names the hot path uses
(`calibrate_mean_var` :97-107, `get_lds_kernel_window` :110-122) plus the
small host helpers train.py needs (:10-94).  The arithmetic of
calibrate_mean_var runs in libdirb200 (dirb200_fds_calibrate_fwd)."""
import os
import shutil

import numpy as np
import torch
from scipy.ndimage import gaussian_filter1d
from scipy.signal.windows import triang

import _lib


class AverageMeter(object):
    def __init__(self, name, fmt=':f'):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ('{name} {val' + self.fmt + '} ({avg' + self.fmt + '})').format(**self.__dict__)


class ProgressMeter(object):
    def __init__(self, num_batches, meters, prefix=""):
        digits = len(str(num_batches // 1))
        self.batch_fmtstr = '[{:' + str(digits) + 'd}/' + ('{:' + str(digits) + 'd}').format(num_batches) + ']'
        self.meters, self.prefix = meters, prefix

    def display(self, batch):
        import logging
        logging.info('\t'.join([self.prefix + self.batch_fmtstr.format(batch)] + [str(m) for m in self.meters]))


def prepare_folders(args):
    for folder in (args.store_root, os.path.join(args.store_root, args.store_name)):
        if not os.path.exists(folder):
            os.makedirs(folder, exist_ok=True)


def adjust_learning_rate(optimizer, epoch, args):
    lr = args.lr
    for milestone in args.schedule:
        lr *= 0.1 if epoch >= milestone else 1.
    for group in optimizer.param_groups:
        group['lr'] = 0. if group.get('name') == 'noupdate_params' else lr


def save_checkpoint(args, state, is_best, prefix=''):
    filename = f"{args.store_root}/{args.store_name}/{prefix}ckpt.pth.tar"
    torch.save(state, filename)
    if is_best:
        shutil.copyfile(filename, filename.replace('pth.tar', 'best.pth.tar'))


def calibrate_mean_var(matrix, m1, v1, m2, v2, clip_min=0.1, clip_max=10):
    """(matrix - m1) * sqrt(clamp(v2 / v1)) + m2 with the reference's two
    early-outs, on the GPU; `matrix` [n, d] is calibrated in place (all rows
    share one statistics row) and returned."""
    _lib.require_cuda(matrix, m1, v1, m2, v2)
    assert matrix.dtype == torch.float32 and matrix.dim() == 2 and matrix.is_contiguous()
    n, d = matrix.shape
    labels = torch.zeros(n, dtype=torch.float32, device=matrix.device)
    rowbin = torch.empty(n, dtype=torch.int32, device=matrix.device)
    scratch = torch.empty(2, dtype=torch.int32, device=matrix.device)
    tabs = [t.reshape(1, d).to(torch.float32).contiguous() for t in (m1, v1, m2, v2)]
    _lib.call("dirb200_fds_calibrate_fwd", _lib.ptr(matrix), _lib.ptr(labels), n, d, 1, 0, _lib.BIN_AGE,
              _lib.ptr(tabs[0]), _lib.ptr(tabs[1]), _lib.ptr(tabs[2]), _lib.ptr(tabs[3]),
              float(clip_min), float(clip_max), _lib.ptr(rowbin), _lib.ptr(scratch), _lib.stream_ptr())
    return matrix


def get_lds_kernel_window(kernel, ks, sigma):
    """float64 taps normalised to max 1 (agedb-dir/utils.py:110-122)."""
    assert kernel in ['gaussian', 'triang', 'laplace']
    half_ks = (ks - 1) // 2
    if kernel == 'gaussian':
        impulse = np.zeros(ks)
        impulse[half_ks] = 1.
        resp = gaussian_filter1d(impulse, sigma=sigma)
        return resp / max(resp)
    if kernel == 'triang':
        return triang(ks)
    taps = np.asarray([np.exp(-abs(x) / sigma) / (2. * sigma) for x in np.arange(-half_ks, half_ks + 1)])
    return taps / max(taps)


class nvtx_range:
    """NVTX range around a phase of the step (tracing), visible in nsys / ncu timelines.  Active only with
    DIRB200_NVTX=1 (torch.cuda.nvtx.range_push / range_pop cost a few microseconds of host time each)."""
    _on = None

    def __init__(self, name):
        self.name = name
        if nvtx_range._on is None:
            import os
            nvtx_range._on = os.environ.get("DIRB200_NVTX") == "1"

    def __enter__(self):
        if nvtx_range._on:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if nvtx_range._on:
            import torch
            torch.cuda.nvtx.range_pop()
        return False
