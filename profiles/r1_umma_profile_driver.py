import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, rxinfer_jl_b200 as rx
from oracle import lgssm
ctx = rx.Context(0)
mod = {k: np.asarray(v, np.float32) for k, v in lgssm.dense_model(64).items()}
y = torch.randn(1000, 64, 18944, device="cuda") * 3.3
for i in range(2):
    r = ctx.lgssm(y, **mod, smooth=True, cov_shared_out=True)
torch.cuda.synchronize()
