"""CPU-only: the C-ABI library loads and exports every symbol include/dirb200.h
declares; argument validation works without a GPU (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dirb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dirb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import _lib
    import _convlib, resnet  # noqa: F401  (register the conv-stack / runner bindings)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dirb200.h but not exported"
    # and the Python binding table covers the header
    assert set(syms) <= set(_lib.exported_symbols()), set(syms) - set(_lib.exported_symbols())


def test_version_and_error_channel():
    import _lib
    assert _lib.raw("dirb200_version")() >= 100
    # invalid argument is rejected before any CUDA call
    rc = _lib.raw("dirb200_fds_smooth_tables")(None, 10, 4, None, 5, None, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = _lib.raw("dirb200_loss_fwd_bwd")(99, None, None, None, 4, 0.0, 1.0, 0, 1.0, None, None, None, 0, None)
    assert rc == -1 and "kind" in _lib.last_error()


def test_no_cpu_fallback():
    import pytest
    import torch
    import _lib
    from loss import weighted_l1_loss
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Dirb200Error):
        weighted_l1_loss(torch.zeros(4, 1), torch.zeros(4, 1))


def test_wgrad_split_search_never_leaves_a_straggler_wave():
    """Host-side scheduling of the wgrad split-K GEMM (csrc/conv_igemm.cu: conv_wgrad_splits), observable without a
    GPU through dirb200_conv_wgrad_workspace_bytes = splits * K_total * Cout * 4 (num_sms() falls back to 148 when no
    device is present).  For every conv of the batch-256 ResNet-50: the work items (tiles x splits) must fill the
    persistent CTAs' waves to >= 85 % -- the earlier ceil(2*SMs/tiles) rule left e.g. 297 items (a third wave for one
    item) on the 3x3 layers -- and every split keeps >= 8 k-blocks."""
    import _lib, _convlib  # noqa: F401
    sms = 148
    convs = []          # (h, cin, cout, k, stride, pad)
    h, inpl = 56, 64
    for li, nb in enumerate((3, 4, 6, 3)):
        pl = 64 << li
        for b in range(nb):
            s = 2 if (b == 0 and li > 0) else 1
            convs += [(h, inpl, pl, 1, 1, 0), (h, pl, pl, 3, s, 1), (h // s, pl, pl * 4, 1, 1, 0)]
            if b == 0:
                convs.append((h, inpl, pl * 4, 1, s, 0))
            inpl, h = pl * 4, h // s
    assert len(convs) == 52
    worst = 1.0
    for (hh, cin, cout, k, s, p) in convs:
        nbytes = _lib.raw("dirb200_conv_wgrad_workspace_bytes")(256, hh, hh, cin, cout, k, k, s, p, 0)
        splits = nbytes // (k * k * cin * cout * 4)
        assert splits >= 1 and nbytes == splits * k * k * cin * cout * 4
        ho = (hh + 2 * p - k) // s + 1
        kblocks = (256 * ho * ho + 63) // 64
        bn = 256 if cout % 256 == 0 else (128 if cout % 128 == 0 else 64)
        tiles = ((k * k * cin // 64 + 1) // 2) * (cout // bn)
        items = tiles * splits
        waves = -(-items // sms)
        fill = items / (waves * sms)
        worst = min(worst, fill)
        assert fill >= 0.85, (hh, cin, cout, k, s, splits, items)
        assert kblocks // splits >= 8
    assert worst >= 0.85


def test_conv_plan_selects_the_intended_gemm_forms():
    """Host-only selection logic of the conv stage (dirb200_conv_plan; no device needed): the batch-256 ResNet-50 shapes
    get the forms DESIGN.md section 4.1 describes."""
    import ctypes
    import _lib, _convlib  # noqa: F401

    def plan(h, cin, cout, k, s, p, op, stem=0, hw_in=None):
        a = (ctypes.c_int * 7)()
        hh = hw_in or h
        rc = _lib.raw("dirb200_conv_plan")(256, hh, hh, cin, cout, k, k, s, p, stem, op, a)
        assert rc == 0, _lib.last_error()
        return dict(zip(("bn", "pairs", "feed", "patch_rows", "splits", "launches", "bn_moments"), a))

    GATHER, TILED, IM2COL, PATCH = 0, 1, 2, 3
    # layer1 conv2 (64 -> 64 3x3 at 56x56): patch-resident in all three passes, 2 padded rows per tile, one partial per CTA
    for op in (0, 1, 2):
        q = plan(56, 64, 64, 3, 1, 1, op)
        assert (q["feed"], q["patch_rows"], q["bn"]) == (PATCH, 2, 64), q
    assert plan(56, 64, 64, 3, 1, 1, 2)["splits"] == 148 and plan(56, 64, 64, 3, 1, 1, 1)["bn_moments"] == 1
    # 1x1 stride-1 GEMMs: tiled TMA; K-heavy 256-wide ones as CTA pairs
    assert plan(56, 64, 256, 1, 1, 0, 0) == dict(bn=256, pairs=0, feed=TILED, patch_rows=0, splits=1, launches=1, bn_moments=0)
    q = plan(14, 256, 1024, 1, 1, 0, 0)
    assert (q["bn"], q["pairs"], q["feed"]) == (256, 1, TILED)
    q = plan(7, 2048, 512, 1, 1, 0, 1)                   # dgrad: N = Cin = 2048 -> 256-wide pair tiles, carries BN moments
    assert (q["bn"], q["pairs"], q["bn_moments"]) == (256, 1, 1)
    # 3x3 layers: im2col TMA; pairs from layer3 on; stride-2 dgrad = 4 TMA-fed parity-class launches, no BN moments
    q = plan(28, 128, 128, 3, 1, 1, 0)
    assert (q["bn"], q["pairs"], q["feed"]) == (128, 0, IM2COL)
    q = plan(14, 256, 256, 3, 1, 1, 0)
    assert (q["bn"], q["pairs"], q["feed"]) == (256, 1, IM2COL)
    q = plan(56, 128, 128, 3, 2, 1, 1)
    assert (q["feed"], q["launches"], q["bn_moments"]) == (IM2COL, 4, 0)
    q = plan(56, 256, 512, 1, 2, 0, 1)                   # 1x1 stride-2 downsample: only the (even, even) class has a tap
    assert q["launches"] == 1
    # wgrad never runs as pairs; every split keeps work (checked in detail by the test above)
    assert plan(14, 256, 256, 3, 1, 1, 2)["pairs"] == 0
    # 5x5 (NYUD2 refinement conv) goes through im2col TMA as well; the stem keeps the cp.async gather
    assert plan(240, 128, 128, 5, 1, 2, 0)["feed"] == IM2COL
    assert plan(224, 3, 64, 7, 2, 3, 0, stem=1)["feed"] == GATHER
