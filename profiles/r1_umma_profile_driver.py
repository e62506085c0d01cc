import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import rxinfer_jl_b200 as rx
from bench import dense_model_f32
d, b, T = 64, int(os.environ.get("B", "4096")), 1000
ctx = rx.Context(0)
md = dense_model_f32(d)
y = torch.randn(T, d, b, device="cuda") * 3.3
mean = torch.empty(T, d, b, device="cuda")
for _ in range(3):
    ctx.lgssm(y, **md, smooth=True, out_mean=mean, cov_shared_out=True)
torch.cuda.synchronize()
