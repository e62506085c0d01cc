"""Regenerates the committed fixtures under tests/golden/.

reference_goldens.json : fixed-data golden values copied from the reference's own tests (each with
                         its file:line) -- the oracle is pinned against these.
lgssm_*.npz            : small seeded LGSSM problems with the oracle's fp64 posteriors, so the GPU
                         parity tests also compare against committed numbers (guards oracle drift).
Run from the repo root:  python tests/golden/make_fixtures.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hgf, lgssm  # noqa: E402

GOLDENS = {
    "gamma_aliases_mean_s": {"value": 9.468846338832027, "ref": "test/models/aliases/aliases_gamma_tests.jl:43"},
    "gamma_aliases_bfe": {"value": 4.385584096993327, "ref": "test/models/aliases/aliases_gamma_tests.jl:44"},
    "two_node_mean_a": {"value": 1.5, "atol": 0.1, "ref": "test/models/models_tests.jl:254"},
    "two_node_bfe_a": {"value": 3.51551, "atol": 0.1, "ref": "test/models/models_tests.jl:255"},
    "two_node_mean_b": {"value": 1.0, "atol": 0.1, "ref": "test/models/models_tests.jl:308"},
    "two_node_bfe_b": {"value": 2.26551, "atol": 0.1, "ref": "test/models/models_tests.jl:309"},
    "entropy_normal_0_1": {"value": 1.4189385332046727, "ref": "test/score/diagnostics_tests.jl:24"},
    "published_smoothing_ms_d2_T1000": {"value": 77.231, "ref": "benchmarks/Linear Multivariate Gaussian State Space Model Benchmark.ipynb:799"},
    "published_filtering_ms_d2_T1000": {"value": 10.464, "ref": "benchmarks/Linear Multivariate Gaussian State Space Model Benchmark.ipynb:799"},
}


def f32(model):
    return {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in model.items()}


def main():
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(GOLDENS, f, indent=1, sort_keys=True)
    for d, T, batch in [(4, 64, 8), (2, 48, 6)]:
        mod = f32(lgssm.notebook_model(d))
        _, y = lgssm.generate_data(mod, T, batch, seed=42)
        r = lgssm.smooth_reference_schedule(y, **mod)
        np.savez_compressed(os.path.join(HERE, f"lgssm_d{d}_T{T}_b{batch}.npz"), y=y,
                            mean=r["mean"], cov=r["cov"], filt_mean=r["filt_mean"], filt_cov=r["filt_cov"],
                            neg_log_evidence=r["neg_log_evidence"], **{f"model_{k}": v for k, v in mod.items()})
    z, x, y = hgf.generate_data(40, 8)
    out = hgf.hgf_filter(y, iters=10)
    np.savez_compressed(os.path.join(HERE, "hgf_T40_b8.npz"), y=y, out=out)


if __name__ == "__main__":
    main()
