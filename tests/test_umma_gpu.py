"""tcgen05 / TMEM plumbing self-test: one 128 x 64 x 128 product on the tensor pipe with the 3xTF32
split must match the fp64 product to fp32-level accuracy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_umma_tf32x3_matches_fp64(ctx):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((128, 128)).astype(np.float32)
    B = rng.standard_normal((64, 128)).astype(np.float32)
    D = ctx.selftest_umma(torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")).cpu().numpy()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(D - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err
    # structured probe: each output depends on one (row, col) pair only
    A2 = np.zeros((128, 128), np.float32); B2 = np.zeros((64, 128), np.float32)
    A2[np.arange(128), np.arange(128)] = np.arange(1, 129)
    B2[np.arange(64), np.arange(64) * 2] = 1.0
    D2 = ctx.selftest_umma(torch.as_tensor(A2, device="cuda"), torch.as_tensor(B2, device="cuda")).cpu().numpy()
    assert np.array_equal(D2, A2.astype(np.float64) @ B2.astype(np.float64).T)
