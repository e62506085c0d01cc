"""Size-independent properties of the oracle's rule restatements (hypothesis): the algebra the reference's rules obey
whatever the data -- products are commutative / associative and agree with the marginal rule, the two
parametrisations round-trip, the `*` rules are adjoint, `+` inverts, the smoothing schedule is affine in the
observations and equals the textbook Kalman / RTS smoother on random stable models, chunked streaming = one pass."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import lgssm, rules as R

SET = settings(max_examples=25, deadline=None)


def spd(rng, d, scale=1.0):
    X = rng.standard_normal((d, d))
    return scale * (X @ X.T + d * np.eye(d))


@SET
@given(st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_product_algebra_and_round_trips(d, seed):
    rng = np.random.default_rng(seed)
    ms = [(rng.standard_normal(d), spd(rng, d)) for _ in range(3)]
    ws = [R.meancov_to_wmp(*m) for m in ms]
    for (mu, S), w in zip(ms, ws):                                    # mean_cov(weightedmean_precision(.)) == id
        mu2, S2 = R.wmp_to_meancov(*w)
        assert np.allclose(mu2, mu, atol=1e-9) and np.allclose(S2, S, atol=1e-9)
    ab = R.prod_gaussian_wmp(ws[0], ws[1]); ba = R.prod_gaussian_wmp(ws[1], ws[0])
    assert np.array_equal(ab[0], ba[0]) and np.array_equal(ab[1], ba[1])          # commutative, exactly
    l = R.prod_gaussian_wmp(R.prod_gaussian_wmp(ws[0], ws[1]), ws[2])
    r = R.prod_gaussian_wmp(ws[0], R.prod_gaussian_wmp(ws[1], ws[2]))
    assert np.allclose(l[0], r[0], atol=1e-12) and np.allclose(l[1], r[1], atol=1e-12)
    m1 = R.marginal_from_messages(ws); m2 = R.wmp_to_meancov(*l)
    assert np.allclose(m1[0], m2[0], atol=1e-10) and np.allclose(m1[1], m2[1], atol=1e-10)
    # product of a message with a vanishing-precision (vague) message is the message itself
    vague = (np.zeros(d), 1e-12 * np.eye(d))
    mu3, S3 = R.wmp_to_meancov(*R.prod_gaussian_wmp(ws[0], vague))
    assert np.allclose(mu3, ms[0][0], atol=1e-6) and np.allclose(S3, ms[0][1], rtol=1e-6)


@SET
@given(st.integers(1, 5), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_multiplication_rules_are_adjoint_and_addition_inverts(d_in, d_out, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d_out, d_in))
    m_in = (rng.standard_normal(d_in), spd(rng, d_in))
    mu_o, S_o = R.multiplication_out(A, m_in)
    assert np.allclose(S_o, S_o.T) and np.all(np.linalg.eigvalsh(S_o) > -1e-10)
    # <backward(m_out), x> pairing: xi_in . x == xi_out . (A x),  x' W_in x == (A x)' W_out (A x)
    m_out = (rng.standard_normal(d_out), spd(rng, d_out))
    xi_o, W_o = R.meancov_to_wmp(*m_out)
    xi_i, W_i = R.multiplication_in((xi_o, W_o), A)
    x = rng.standard_normal(d_in)
    assert np.isclose(xi_i @ x, xi_o @ (A @ x)) and np.isclose(x @ W_i @ x, (A @ x) @ W_o @ (A @ x))
    # `+`: in1(out(a, b), b) recovers the mean of a and adds twice the covariance of b
    a = (rng.standard_normal(d_in), spd(rng, d_in)); b = (rng.standard_normal(d_in), spd(rng, d_in))
    o = R.addition_out(a, b)
    back = R.addition_in1(o, b)
    assert np.allclose(back[0], a[0]) and np.allclose(back[1], a[1] + 2 * b[1])
    back2 = R.addition_in2(o, a)
    assert np.allclose(back2[0], b[0]) and np.allclose(back2[1], b[1] + 2 * a[1])


def stable_model(rng, d, m):
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    return dict(A=0.95 * Q, B=rng.standard_normal((m, d)), P=spd(rng, d, 0.05), Q=spd(rng, m, 0.5),
                m0=rng.standard_normal(d), S0=spd(rng, d, 2.0))


@SET
@given(st.integers(1, 4), st.integers(1, 4), st.integers(1, 40), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_schedule_is_kalman_rts_and_affine_in_the_data(d, m, T, seed, tf):
    rng = np.random.default_rng(seed)
    mod = stable_model(rng, d, min(m, d))
    mm = mod["B"].shape[0]
    y1 = rng.standard_normal((T, mm, 2)); y2 = rng.standard_normal((T, mm, 2))
    r1 = lgssm.smooth_reference_schedule(y1, **mod, transition_first=tf)
    k1 = lgssm.kalman_rts(y1, **mod, transition_first=tf)
    assert np.allclose(r1["mean"], k1["mean"], atol=1e-8) and np.allclose(r1["cov"], k1["cov"], atol=1e-8)
    assert np.allclose(r1["neg_log_evidence"], k1["neg_log_evidence"], rtol=1e-9, atol=1e-9)
    # affine: mean(a y1 + (1 - a) y2) = a mean(y1) + (1 - a) mean(y2); covariances do not depend on the data
    a = 0.3
    r2 = lgssm.smooth_reference_schedule(y2, **mod, transition_first=tf)
    r3 = lgssm.smooth_reference_schedule(a * y1 + (1 - a) * y2, **mod, transition_first=tf)
    assert np.allclose(r3["mean"], a * r1["mean"] + (1 - a) * r2["mean"], atol=1e-9)
    assert np.allclose(r3["cov"], r1["cov"], atol=1e-12) and np.allclose(r1["cov"][..., 0], r1["cov"][..., 1], atol=1e-12)
    eig = np.linalg.eigvalsh(np.moveaxis(r1["cov"], -1, 1))
    assert np.all(eig > 0)                                                # mlgssm_test.jl:126 for every model


@SET
@given(st.integers(1, 3), st.integers(2, 30), st.integers(0, 2 ** 31 - 1))
def test_streaming_filter_chunks_compose(d, T, seed):
    """The streaming engine's carry: filtering [0, k) and then [k, T) from the carried posterior equals one pass."""
    rng = np.random.default_rng(seed)
    mod = stable_model(rng, d, d)
    y = rng.standard_normal((T, d, 3))
    whole = lgssm.filter_streaming(y, **mod)
    k = int(rng.integers(1, T))
    first = lgssm.filter_streaming(y[:k], **mod)
    for c in range(3):                      # the oracle takes one prior per call: continue chain by chain
        mod2 = dict(mod, m0=first["mean"][-1, :, c], S0=first["cov"][-1, :, :, c])
        second = lgssm.filter_streaming(y[k:, :, c:c + 1], **mod2)
        assert np.allclose(second["mean"][..., 0], whole["mean"][k:, :, c], atol=1e-9)
        assert np.allclose(second["cov"][..., 0], whole["cov"][k:, :, :, c], atol=1e-9)
