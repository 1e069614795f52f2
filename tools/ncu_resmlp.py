"""One large-batch ResMLP launch for ncu: python tools/ncu_resmlp.py [n]"""
import sys
import torch
from wtracker_b200.neural.engine import ResMLPEngine
from wtracker_b200.neural.mlp import load_worm_predictor
from wtracker_b200.paths import RESMLP_100
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
eng = ResMLPEngine(load_worm_predictor(RESMLP_100))
x = torch.randn((n, 28), device="cuda")
out = torch.empty((n, 2), device="cuda")
for _ in range(3):
    eng.forward(x, out)
torch.cuda.synchronize()
