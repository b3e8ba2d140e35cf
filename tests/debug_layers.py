"""Layer-by-layer comparison of the native runner with the oracle (manual debugging aid, GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from test_gpu_resnet import make_model, oracle_params, rel
from util import det_param
from oracle import resnet_ref as R

n, hw = int(sys.argv[1]), int(sys.argv[2])
m = make_model(); m.train()
p = oracle_params(m)
x = det_param(f"x{n}", (n, 3, hw, hw), 1.0).cuda()
enc = m._run_forward(x, training=True)
taps = {}
with torch.no_grad():
    renc = R.forward_encoding(p, x, quant=True, taps=taps)
names = [("stem.y", -1, 0), ("stem.a", -1, 1), ("stem.pool", -1, 6)]
for b in range(16):
    for w in range(7):
        if f"{b}.{w}" in taps:
            names.append((f"{b}.{w}", b, w))
for name, b, w in names:
    got = m.peek(x.shape, b, w)
    print(f"{name:10s} rel {rel(got, taps[name]):.4f}  shape {tuple(got.shape)}")
print("enc", rel(enc, renc))
