import numpy as np


def f32_model(model):
    """Model matrices as the ABI sees them: rounded to fp32 once, upcast for the oracle."""
    return {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in model.items()}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# parity gate (BASELINE.json north_star / SURVEY.md 8d): fp32 GPU vs fp64 oracle
TOL_MEAN = 1e-5      # relative L2 over all (t, k, chain)
TOL_COV = 1e-4       # relative Frobenius
TOL_NLE = 1e-5       # relative, per chain
