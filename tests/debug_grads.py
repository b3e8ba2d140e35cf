"""Per-parameter gradient comparison with the oracle (manual debugging aid, GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from test_gpu_resnet import _fb, rel, cos

layers = tuple(int(c) for c in sys.argv[1].split(","))
n, hw = int(sys.argv[2]), int(sys.argv[3])
m, p, stats, enc, renc, loss, rloss = _fb(layers, n, hw)
named = dict(m.named_parameters())
for name, rp in p.items():
    g, rg = named[name].grad, rp.grad
    print(f"{name:34s} cos {cos(g, rg):.5f} rel {rel(g, rg):.4f} |g| {g.norm().item():.4e} |ref| {rg.norm().item():.4e}")
print("enc rel", rel(enc, renc), "loss", loss, rloss)
