"""ncu driver (round 2): one launch each of the latent-AR kernel (order 5) and the d = 64 rule kernels.
    ncu --set full --clock-control none --import-source on -k regex:lar_vmp_kernel -c 1 -o gpurun_out/r2_lar python profiles/r2_lar_profile_driver.py
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_rules_large_launches.csv python profiles/r2_lar_profile_driver.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rxinfer_jl_b200 as rx  # noqa: E402

ctx = rx.Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
y = torch.randn(500, 16384, device="cuda", generator=g)
ctx.lar_vmp(y, 5, 5.0, iterations=15)
d, n = 64, 4096
mu = torch.randn(d, n, device="cuda", generator=g)
X = torch.randn(d, d, n, device="cuda", generator=g)
S = (torch.einsum("ikn,jkn->ijn", X, X) + d * torch.eye(d, device="cuda")[:, :, None]).contiguous()
A = np.linalg.qr(np.random.default_rng(0).standard_normal((d, d)))[0].astype(np.float32)
ctx.rule_add_cov(mu, S, np.eye(d, dtype=np.float32))
ctx.rule_mul_out(A, mu, S)
ctx.rule_mul_in(A, mu, S)
ctx.meancov_to_wmp(mu, S)
torch.cuda.synchronize()
print("ok")
