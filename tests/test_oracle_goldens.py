"""Pins the oracle against every fixed-data golden the reference's own tests hold for this path
 and cross-checks its independent restatements against each other."""
import json
import os

import numpy as np
import pytest

from oracle import c_twin, hgf, lgssm, rules as R, vmp

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def test_gamma_aliases_mean_bit_exact_and_bfe():
    means, fes = vmp.gamma_aliases_golden(iterations=100)
    assert means[-1] == G["gamma_aliases_mean_s"]["value"]          # bit-exact
    assert abs(fes[-1] - G["gamma_aliases_bfe"]["value"]) < 1e-12
    assert np.all(np.diff(fes) <= 1e-14)                             # aliases_gamma_tests.jl:45


def test_two_node_gaussian_goldens():
    m, v, bfe = vmp.two_node_gaussian(3.0, 1.0, 0.0, 1.0)
    assert abs(m - G["two_node_mean_a"]["value"]) < 1e-12 and abs(bfe - 3.5155121234846454) < 1e-12
    assert abs(bfe - G["two_node_bfe_a"]["value"]) < G["two_node_bfe_a"]["atol"]
    m, v, bfe = vmp.two_node_gaussian(2.0, 1.0, 0.0, 1.0)
    assert abs(m - G["two_node_mean_b"]["value"]) < 1e-12 and abs(bfe - 2.2655121234846454) < 1e-12


def test_two_node_through_lgssm_schedule():
    """The same golden through the LGSSM oracle (d = m = 1, T = 1): BFE == -log evidence on a tree."""
    one = np.array([[1.0]])
    r = lgssm.smooth_reference_schedule(np.zeros((1, 1, 1), np.float32), one, one, one, one, np.array([3.0]), one)
    assert abs(r["mean"][0, 0, 0] - 1.5) < 1e-14 and abs(r["cov"][0, 0, 0, 0] - 0.5) < 1e-14
    assert abs(r["neg_log_evidence"][0] - 3.5155121234846454) < 1e-12


def test_normal_entropy_golden():
    assert R.normal_entropy(1.0) == G["entropy_normal_0_1"]["value"]


@pytest.mark.parametrize("d,T,batch", [(2, 50, 3), (4, 120, 4)])
def test_schedule_equals_kalman_rts(d, T, batch):
    mod = lgssm.notebook_model(d)
    _, y = lgssm.generate_data(mod, T, batch)
    a = lgssm.smooth_reference_schedule(y, **mod)
    b = lgssm.kalman_rts(y, **mod)
    for k in ("mean", "cov", "filt_mean", "filt_cov"):
        assert np.abs(a[k] - b[k]).max() < 1e-11, k
    assert np.abs(a["neg_log_evidence"] - b["neg_log_evidence"]).max() < 1e-9
    # covariances SPD (mlgssm_test.jl:126)
    assert np.all(np.linalg.eigvalsh(np.moveaxis(a["cov"], -1, 1)) > 0)


def test_schedule_with_missing_data_equals_kalman():
    mod = lgssm.notebook_model(4)
    _, y = lgssm.generate_data(mod, 60, 5)
    rng = np.random.default_rng(3)
    mask = rng.random((60, 5)) > 0.3
    mask[-4:, 1] = False                       # trailing missing data: "no message" branch
    a = lgssm.smooth_reference_schedule(y, **mod, mask=mask)
    b = lgssm.kalman_rts(y, **mod, mask=mask)
    assert np.abs(a["mean"] - b["mean"]).max() < 1e-10
    assert np.abs(a["cov"] - b["cov"]).max() < 1e-10


def test_per_chain_models():
    rng = np.random.default_rng(0)
    base = lgssm.notebook_model(2)
    batch = 4
    mods = {k: np.stack([v * (1 + 0.1 * i) if k in ("P", "Q", "S0") else v for i in range(batch)]) for k, v in base.items()}
    _, y = lgssm.generate_data(base, 30, batch)
    a = lgssm.smooth_reference_schedule(y, **mods)
    for i in range(batch):
        one = {k: v[i] for k, v in mods.items()}
        b = lgssm.kalman_rts(y[:, :, i:i + 1], **one)
        assert np.abs(a["mean"][:, :, i] - b["mean"][:, :, 0]).max() < 1e-11


def test_c_twin_matches_numpy_oracle():
    mod = lgssm.notebook_model(4)
    _, y = lgssm.generate_data(mod, 100, 9)
    a = lgssm.smooth_reference_schedule(y, **mod)
    c = c_twin.smooth(y, **mod, nthreads=2)
    assert np.abs(a["mean"] - c["mean"]).max() < 1e-11
    assert np.abs(a["cov"] - c["cov"]).max() < 1e-11
    assert np.abs(a["neg_log_evidence"] - c["neg_log_evidence"]).max() < 1e-9


def test_committed_fixture_is_current():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lgssm_d4_T64_b8.npz"))
    mod = {k[6:]: z[k] for k in z.files if k.startswith("model_")}
    r = lgssm.smooth_reference_schedule(z["y"], **mod)
    assert np.abs(r["mean"] - z["mean"]).max() < 1e-12
    assert np.abs(r["cov"] - z["cov"]).max() < 1e-12


def test_streaming_filter_is_transition_first():
    mod = lgssm.notebook_model(2)
    _, y = lgssm.generate_data(mod, 20, 2)
    s = lgssm.filter_streaming(y, **mod)
    # equals the smoothing-graph filter started from the prior pushed through (A, P)
    m0p = mod["A"] @ mod["m0"]; S0p = mod["A"] @ mod["S0"] @ mod["A"].T + mod["P"]
    f = lgssm.filter_reference_schedule(y, mod["A"], mod["B"], mod["P"], mod["Q"], m0p, S0p)
    assert np.abs(s["mean"] - f["mean"]).max() < 1e-12


def test_hgf_oracle_sane_and_energy_formula():
    z, x, y = hgf.generate_data(200, 16)
    out, fe = hgf.hgf_filter(y, iters=10, return_free_energy=True)
    assert np.all(out[:, 1] > 0) and np.all(out[:, 3] > 0)            # hgf_tests.jl:131-132
    inside = np.abs(out[:, 0] - x) < 3 * np.sqrt(out[:, 1])
    assert inside.mean() > 0.95                                        # hgf_tests.jl:127-130
    # the GCV node energy restated at test/inference/inference_tests.jl:595-606
    m = np.array([[0.3, -0.2]]); V = np.array([[[0.5, 0.1], [0.1, 0.4]]]); qz = (np.array([0.2]), np.array([0.3]))
    psi = (0.3 + 0.2) ** 2 + 0.5 + 0.4 - 0.2
    want = 0.5 * (np.log(2 * np.pi) + (0.2 * 1.0 + 0.0) + psi * np.exp(-0.0) * np.exp(-0.2 + 0.5 * 0.3))
    assert abs(hgf._hgf_step_energy(m, V, qz, 1.0, 0.0)[0] - want) < 1e-14


def test_gauss_hermite_prod_recovers_gaussian_times_gaussian():
    """ELQ with b = 0 is exp(-a z / 2): the product with N(m, v) is N(m - a v / 2, v) exactly."""
    mz, vz = R.prod_normal_elq((np.array([0.7]), np.array([1.3])), (0.8, np.array([0.0]), -0.8, 0.0))
    assert abs(mz[0] - (0.7 - 0.4 * 1.3)) < 1e-10 and abs(vz[0] - 1.3) < 1e-9


def test_stream_vmp_gamma_free_energy_is_monotone():
    """The reference's streaming `test_model1` (test/inference/inference_tests.jl:752-860) asserts
    `all(<=(0), diff(engine.free_energy_history))` for 3 and 4 iterations (:846): the Bethe free energy, averaged over
    the observations, must not increase from one VMP iteration to the next.  Data as in the test (:778-788)."""
    rng = np.random.default_rng(7)
    n, batch = 10, 256
    x = np.cumsum(rng.standard_normal((n, batch)), axis=0)
    y = x + rng.standard_normal((n, batch)) / np.sqrt(10.0)
    for iters in (3, 4):
        out, fe = vmp.stream_vmp_gamma(y, iterations=iters, return_free_energy=True)
        hist = fe.mean(axis=0)                                   # free_energy_history: [iterations, chain]
        assert np.all(np.diff(hist, axis=0) <= 1e-12)
        assert np.all(out[:, 1] > 0) and np.all(out[:, 3] > 0) and np.allclose(out[:, 2], 1.0 + 0.5 * np.arange(1, n + 1)[:, None])
    # the rules are the pinned ones: one datum, one iteration, by hand
    o1 = vmp.stream_vmp_gamma(np.array([[2.0]]), iterations=1)
    mmin = (0.0 / 1e3 + 0.0 * 1.0) / (1 / 1e3 + 1.0); vx = 1.0 / (1.0 + 1.0); mx = vx * (mmin + 2.0 * 1.0)
    assert abs(o1[0, 0, 0] - mx) < 1e-15 and abs(o1[0, 1, 0] - vx) < 1e-15
    assert abs(o1[0, 2, 0] - 1.5) < 1e-15 and abs(o1[0, 3, 0] - (1.0 + 0.5 * ((2.0 - mx) ** 2 + vx))) < 1e-15


def test_mv_iid_wishart_oracle_meets_the_reference_assertions():
    """mv_iid_precision_tests.jl:44-66: n = 1500, d = 2, 10 iterations; `mean(q(m)) ~ m (atol 0.05)`,
    `mean(q(P)) ~ P (atol 0.07)`.  The reference's data come from a StableRNG stream through Distributions.jl's
    MvNormal sampler (not restated), so this is the reference's STATISTICAL assertion on data of the same model --
    parity for this model is unpinned beyond that (stated in DESIGN.md)."""
    from oracle import vmp
    rng = np.random.default_rng(11)      # like the reference's StableRNG(123) draw, a draw with a well-conditioned C: the
    n, d = 1500, 2                       # thresholds are tied to it (the Lambda = 100 I prior shrinks m along weak directions of P)
    m = rng.random(d)
    Lc = rng.standard_normal((d, d))
    C = Lc @ Lc.T
    P = np.linalg.inv(C)
    y = (m[None, :] + rng.standard_normal((n, d)) @ np.linalg.cholesky(C).T)[:, :, None]
    r = vmp.mv_iid_wishart(y, iterations=10)
    assert np.allclose(r["m_mean"][:, 0], m, atol=0.05)
    assert np.allclose(r["E_P"][:, :, 0], P, atol=0.07)


def test_vmp_gamma_smoother_free_energy_closed_form_equals_definition():
    """Bethe free energy of the Gamma-precision VMP around the smoother: the closed form the CUDA kernel evaluates (filter
    evidence under the old E[tau] + the tau terms) equals the dense evaluation of the definition, and decreases."""
    from oracle import vmp
    rng = np.random.default_rng(3)
    T, batch = 30, 4
    x = np.cumsum(rng.standard_normal((T, batch)), axis=0)
    y = x + rng.standard_normal((T, batch)) / np.sqrt(rng.gamma(2.0, 1.0, batch) + 0.3)
    r = vmp.lgssm_gamma_precision(y, iterations=6, return_free_energy=True)
    assert np.allclose(r["free_energy"], r["free_energy_closed_form"], rtol=0, atol=1e-8)
    assert np.all(np.diff(r["free_energy"], axis=0) < 1e-10)
