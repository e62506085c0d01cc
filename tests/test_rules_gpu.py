"""GPU parity of the per-rule kernels against the fp64 rule oracle, through
the C ABI, plus the call_rule / prod mirror of the reference's @call_rule interface."""
import numpy as np
import pytest
import torch

from oracle import rules as R
from util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-5


def spd(rng, n, d, scale=1.0):
    X = rng.standard_normal((n, d, d))
    return scale * (X @ np.swapaxes(X, -1, -2) + d * np.eye(d))


def soa_v(v):   # [n, d] -> [d, n] cuda
    return torch.as_tensor(np.ascontiguousarray(v.T), dtype=torch.float32, device="cuda")


def soa_m(M):   # [n, r, c] -> [r, c, n]
    return torch.as_tensor(np.ascontiguousarray(np.moveaxis(M, 0, -1)), dtype=torch.float32, device="cuda")


def back_v(t):
    return t.cpu().numpy().T.astype(np.float64)


def back_m(t):
    return np.moveaxis(t.cpu().numpy(), -1, 0).astype(np.float64)


def r32(a):
    return a.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 6, 8, 7, 16, 33, 64])
def test_gaussian_rules(ctx, d):
    """d = 7, 16, 33, 64: no register-resident instantiation -- csrc/rxg_rules_large.cu (element-wise pass, left-GEMM with the
    shared matrix, warp-per-message cholinv); same entry points, same tolerances."""
    rng = np.random.default_rng(d)
    n = (1000 if d <= 8 else 301) + d                   # ragged vs block size / vs the 8 messages per CTA of the large path
    mu, S = r32(rng.standard_normal((n, d))), r32(spd(rng, n, d))
    mu2, S2 = r32(rng.standard_normal((n, d))), r32(spd(rng, n, d))
    Sig = r32(spd(rng, 1, d)[0])
    A = r32(rng.standard_normal((d, d)))
    # rule #2 / #3: + Sigma
    for which in ("out", "mean"):
        mo, So = ctx.rule_add_cov(soa_v(mu), soa_m(S), Sig, which)
        ref = R.mvnormal_meancov_out((mu, S), Sig)
        assert rel_l2(back_v(mo), ref[0]) < TOL and rel_l2(back_m(So), ref[1]) < TOL
    # per-message Sigma
    mo, So = ctx.rule_add_cov(soa_v(mu), soa_m(S), soa_m(S2), "out")
    assert rel_l2(back_m(So), S + S2) < TOL
    # rule #3 from data
    mo, So = ctx.rule_mean_from_data(soa_v(mu), Sig)
    assert rel_l2(back_v(mo), mu) < 1e-7 and rel_l2(back_m(So), np.broadcast_to(Sig, S.shape)) < 1e-7
    # rule #1
    mo, So = ctx.rule_mul_out(A, soa_v(mu), soa_m(S))
    ref = R.multiplication_out(A, (mu, S))
    assert rel_l2(back_v(mo), ref[0]) < TOL and rel_l2(back_m(So), ref[1]) < TOL
    # rule #4 (includes the cholinv conversion)
    xi, W, st = ctx.rule_mul_in(A, soa_v(mu), soa_m(S))
    ref = R.multiplication_in(R.meancov_to_wmp(mu, S), A)
    assert rel_l2(back_v(xi), ref[0]) < 5 * TOL and rel_l2(back_m(W), ref[1]) < 5 * TOL
    assert int(st.abs().sum()) == 0
    # rule #5
    mo, So = ctx.rule_add_out(soa_v(mu), soa_m(S), soa_v(mu2), soa_m(S2))
    assert rel_l2(back_v(mo), mu + mu2) < TOL and rel_l2(back_m(So), S + S2) < TOL
    mo, So = ctx.rule_add_in(soa_v(mu), soa_m(S), soa_v(mu2), soa_m(S2))
    ref = R.addition_in1((mu, S), (mu2, S2))
    assert rel_l2(back_v(mo), ref[0]) < TOL and rel_l2(back_m(So), ref[1]) < TOL
    # prod + conversions + marginal
    xi1, W1, _ = ctx.meancov_to_wmp(soa_v(mu), soa_m(S))
    ref1 = R.meancov_to_wmp(mu, S)
    assert rel_l2(back_v(xi1), ref1[0]) < 5 * TOL and rel_l2(back_m(W1), ref1[1]) < 5 * TOL
    xi2, W2, _ = ctx.meancov_to_wmp(soa_v(mu2), soa_m(S2))
    xp, Wp = ctx.prod_gaussian(xi1, W1, xi2, W2)
    mm, Sm, st = ctx.marginal_gaussian([(xi1, W1), (xi2, W2)])
    refm = R.marginal_from_messages([R.meancov_to_wmp(mu, S), R.meancov_to_wmp(mu2, S2)])
    assert rel_l2(back_v(mm), refm[0]) < 10 * TOL and rel_l2(back_m(Sm), refm[1]) < 10 * TOL
    mb, Sb, _ = ctx.wmp_to_meancov(xp, Wp)
    assert rel_l2(back_m(Sb), refm[1]) < 10 * TOL


def test_non_spd_is_reported_per_message(ctx, rx):
    S = np.stack([np.eye(2), np.array([[1.0, 2.0], [2.0, 1.0]])])          # second one indefinite
    mu = np.zeros((2, 2))
    xi, W, st = ctx.meancov_to_wmp(soa_v(mu), soa_m(S))
    assert st.cpu().tolist() == [0, rx._lib.RXG_ERR_NOT_SPD]


def test_large_d_rules_report_non_spd_and_refuse_per_message_matrices(ctx, rx):
    rng = np.random.default_rng(3)
    d, n = 24, 19
    S = spd(rng, n, d)
    S[5] = -S[5]                                                            # one indefinite message
    S[11, 3, 3] = -1.0
    xi, W, st = ctx.meancov_to_wmp(soa_v(np.zeros((n, d))), soa_m(S))
    want = [0] * n; want[5] = rx._lib.RXG_ERR_NOT_SPD; want[11] = rx._lib.RXG_ERR_NOT_SPD
    assert st.cpu().tolist() == want
    ok = [i for i in range(n) if i not in (5, 11)]
    assert rel_l2(back_m(W)[ok], np.linalg.inv(r32(S)[ok])) < 1e-4
    with pytest.raises(rx.RxGaussError):                                    # per-message A only on the register-resident shapes
        ctx.rule_mul_out(soa_m(rng.standard_normal((n, d, d))), soa_v(np.zeros((n, d))), soa_m(S))
    with pytest.raises(rx.RxGaussError):
        ctx.meancov_to_wmp(soa_v(np.zeros((n, 65))), soa_m(np.tile(np.eye(65), (n, 1, 1))))


@pytest.mark.parametrize("shape", [(2, 4), (3, 5), (16, 64), (64, 16), (20, 20)])
def test_rectangular_multiplication(ctx, shape):
    rng = np.random.default_rng(0)
    n = 257
    do, di = shape
    B = r32(rng.standard_normal((do, di)))
    mu, S = r32(rng.standard_normal((n, di))), r32(spd(rng, n, di))
    mo, So = ctx.rule_mul_out(B, soa_v(mu), soa_m(S))
    ref = R.multiplication_out(B, (mu, S))
    assert rel_l2(back_v(mo), ref[0]) < TOL and rel_l2(back_m(So), ref[1]) < TOL
    my, Sy = r32(rng.standard_normal((n, do))), r32(spd(rng, n, do))
    xi, W, _ = ctx.rule_mul_in(B, soa_v(my), soa_m(Sy))
    ref = R.multiplication_in(R.meancov_to_wmp(my, Sy), B)
    assert rel_l2(back_v(xi), ref[0]) < 5 * TOL and rel_l2(back_m(W), ref[1]) < 5 * TOL


def test_scalar_and_gamma_rules(ctx):
    rng = np.random.default_rng(1)
    n = 4099
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda")
    mo, vo, mm, vm = r32(rng.standard_normal(n)), r32(rng.random(n) + .1), r32(rng.standard_normal(n)), r32(rng.random(n) + .1)
    a, b = ctx.rule_normal_precision_tau(t(mo), t(vo), t(mm), t(vm))
    ra, rb = R.normal_meanprec_tau((mo, vo), (mm, vm))
    assert rel_l2(a.cpu().numpy(), ra) < 1e-6 and rel_l2(b.cpu().numpy(), rb) < 1e-6
    sh, rt = r32(rng.random(n) + 1), r32(rng.random(n) + 1)
    m2, v2 = ctx.rule_normal_precision_out(t(mm), t(vm), t(sh), t(rt))
    ref = R.normal_meanprec_out_q_tau((mm, vm), sh / rt)
    assert rel_l2(v2.cpu().numpy(), ref[1]) < 1e-6
    pa, pb = ctx.prod_gamma(t(sh), t(rt), t(a.cpu().numpy()), t(b.cpu().numpy()))
    ref = R.prod_gamma((sh, rt), (a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64)))
    assert rel_l2(pa.cpu().numpy(), ref[0]) < 1e-6 and rel_l2(pb.cpu().numpy(), ref[1]) < 1e-6
    pm, pv = ctx.prod_normal(t(mo), t(vo), t(mm), t(vm))
    ref = R.prod_normal_mv((mo, vo), (mm, vm))
    assert rel_l2(pm.cpu().numpy(), ref[0]) < 1e-5 and rel_l2(pv.cpu().numpy(), ref[1]) < 1e-6


def test_gamma_aliases_golden_through_gpu_rules(ctx):
    """The reference's fixed-data golden (aliases_gamma_tests.jl:43: mean(q(s)) = 9.468846338832027)
    replayed with the batched GPU rule kernels (fp32 => ~1e-6 relative)."""
    t = lambda v: torch.full((64,), v, dtype=torch.float32, device="cuda")
    n = 6
    ga, gb = [t(1.0)] * n, [t(1e-12)] * n
    for _ in range(100):
        px = [ctx.rule_normal_precision_out(t(1.0), t(0.0), ga[i], gb[i]) for i in range(n)]
        fwd = [px[0]]
        for i in range(1, n):
            m, v = fwd[-1][0] + px[i][0], fwd[-1][1] + px[i][1]          # scalar +(:out)
            fwd.append((m, v))
        back = (t(10.0), t(1.0))
        q_s = ctx.prod_normal(*fwd[-1], *back)
        qx = [None] * n
        for i in range(n - 1, 0, -1):
            to_x = (back[0] - fwd[i - 1][0], back[1] + fwd[i - 1][1])
            qx[i] = ctx.prod_normal(*px[i], *to_x)
            back = (back[0] - px[i][0], back[1] + px[i][1])
        qx[0] = ctx.prod_normal(*px[0], *back)
        for i in range(n):
            a, b = ctx.rule_normal_precision_tau(qx[i][0], qx[i][1], t(1.0), t(0.0))
            ga[i], gb[i] = ctx.prod_gamma(t(1.0), t(1.0), a, b)
    assert abs(float(q_s[0][0]) - 9.468846338832027) < 2e-5


def test_gcv_rules(ctx):
    rng = np.random.default_rng(2)
    n = 2050
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda")
    my, vy = r32(rng.standard_normal(n)), r32(rng.random(n) * 0.1 + 0.01)
    mx, vx = r32(rng.standard_normal(n)), r32(rng.random(n) + 0.1)
    mz, vz = r32(rng.standard_normal(n) * 0.5), r32(rng.random(n) * 0.5 + 0.05)
    k, w = 1.0, 0.0
    mo, vo = ctx.rule_gcv_out(t(mx), t(vx), t(mz), t(vz), k, w)
    ref = R.gcv_y((mx, vx), (mz, vz), k, w)
    assert rel_l2(vo.cpu().numpy(), ref[1]) < 1e-5
    m, V = ctx.marginalrule_gcv_yx(t(my), t(vy), t(mx), t(vx), t(mz), t(vz), k, w)
    rm, rV = R.gcv_marginal_yx((my, vy), (mx, vx), (mz, vz), k, w)
    assert rel_l2(m.cpu().numpy().T, rm) < 1e-5
    assert rel_l2(np.moveaxis(V.cpu().numpy(), -1, 0), rV) < 1e-5
    zp_m, zp_v = r32(rng.standard_normal(n) * 0.3), r32(rng.random(n) * 2 + 0.1)
    gz, gv = ctx.rule_gcv_z_prod(m, V, t(zp_m), t(zp_v), k, w)
    elq = R.gcv_z_elq(rm, rV, k, w)
    rz, rv = R.prod_normal_elq((zp_m, zp_v), elq)
    assert np.abs(gz.cpu().numpy() - rz).max() < 2e-4 * max(1.0, np.abs(rz).max())
    assert rel_l2(gv.cpu().numpy(), rv) < 5e-4


def test_call_rule_mirror(rx, ctx):
    rng = np.random.default_rng(3)
    n, d = 300, 4
    mu, S = r32(rng.standard_normal((n, d))), r32(spd(rng, n, d))
    A = r32(rng.standard_normal((d, d)))
    P = r32(spd(rng, 1, d)[0])
    m_in = rx.MvNormalMeanCovariance(soa_v(mu), soa_m(S))
    out = rx.call_rule(ctx, "*", "out", m_A=rx.PointMass(A), m_in=m_in)
    out = rx.call_rule(ctx, "MvNormalMeanCovariance", "out", **{"m_μ": out, "q_Σ": rx.PointMass(P)})
    ref = R.mvnormal_meancov_out(R.multiplication_out(A, (mu, S)), P)
    assert rel_l2(back_v(out.mean()), ref[0]) < TOL and rel_l2(back_m(out.cov()), ref[1]) < TOL
    back = rx.call_rule(ctx, "*", "in", m_out=out, m_A=rx.PointMass(A))
    assert isinstance(back, rx.MvNormalWeightedMeanPrecision)
    q = rx.prod(ctx, m_in, back)
    refq = R.prod_gaussian_wmp(R.meancov_to_wmp(mu, S), R.multiplication_in(R.meancov_to_wmp(*ref), A))
    assert rel_l2(back_m(q.W), refq[1]) < 10 * TOL
    with pytest.raises(rx.RuleMethodError):
        rx.call_rule(ctx, "*", "in", m_out=out, m_A=rx.PointMass(A), meta=object())


def test_normal_precision_out_mean_field_variant(ctx, rx):
    """(q_mu, q_tau) is the mean-field rule NormalMeanPrecision(mean(q_mu), mean(q_tau)): variance 1/E[tau] only;
    (m_mu, q_tau) is the BP-message rule and adds var(m_mu) (the two were conflated)."""
    from rxinfer_jl_b200.distributions import GammaShapeRate, NormalMeanVariance
    from rxinfer_jl_b200.rules import call_rule
    rng = np.random.default_rng(5)
    n = 513
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda")
    mm, vm = r32(rng.standard_normal(n)), r32(rng.random(n) + .1)
    sh, rt = r32(rng.random(n) + 1), r32(rng.random(n) + 1)
    q = NormalMeanVariance(t(mm), t(vm))
    g = GammaShapeRate(t(sh), t(rt))
    bp = call_rule(ctx, "NormalMeanPrecision", "out", m_μ=q, q_τ=g)
    mf = call_rule(ctx, "NormalMeanPrecision", "out", q_μ=q, q_τ=g)
    ref_bp = R.normal_meanprec_out_q_tau((mm, vm), sh / rt)
    ref_mf = R.normal_meanprec_out_q_mu_q_tau((mm, vm), sh / rt)
    assert rel_l2(bp.v.cpu().numpy(), ref_bp[1]) < 1e-6 and rel_l2(mf.v.cpu().numpy(), ref_mf[1]) < 1e-6
    assert rel_l2(mf.v.cpu().numpy(), rt / sh) < 1e-6
    assert np.array_equal(mf.m.cpu().numpy(), mm) and np.array_equal(bp.m.cpu().numpy(), mm)


def test_structured_tau_and_wishart_rules(ctx, rx):
    """Structured NormalMeanPrecision(:tau)(q_out_mu) and the Wishart-precision rules against the oracle; then the
    fused IID Wishart VMP (rxg_mv_iid_wishart_vmp_f32) against the message-by-message oracle on several data sets."""
    from oracle import vmp
    rng = np.random.default_rng(8)
    n = 777
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    mj = r32(rng.standard_normal((n, 2)))
    Vj = r32(spd(rng, n, 2))
    sh, rt = ctx.rule_normal_precision_tau_joint(soa_v(mj), soa_m(Vj))
    ra, rb = R.normal_meanprec_tau_structured(mj.astype(np.float64), Vj.astype(np.float64))
    assert rel_l2(sh.cpu().numpy(), ra) < 1e-6 and rel_l2(rt.cpu().numpy(), rb) < 1e-5
    for d in (2, 3, 4):
        mo, mm = r32(rng.standard_normal((n, d))), r32(rng.standard_normal((n, d)))
        Vo, Vm = r32(spd(rng, n, d)), r32(spd(rng, n, d))
        df, iS = ctx.rule_mvnormal_precision_lambda(soa_v(mo), soa_m(Vo), soa_v(mm), soa_m(Vm))
        rdf, riS = R.mvnormal_meanprec_lambda((mo.astype(np.float64), Vo.astype(np.float64)), (mm.astype(np.float64), Vm.astype(np.float64)))
        assert rel_l2(df.cpu().numpy(), rdf) < 1e-6 and rel_l2(back_m(iS), riS) < 1e-5
        df2, iS2 = ctx.prod_wishart(df, iS, df, iS)
        pdf, piS = R.prod_wishart((rdf, riS), (rdf, riS))
        assert rel_l2(df2.cpu().numpy(), pdf) < 1e-6 and rel_l2(back_m(iS2), piS) < 1e-5
        EL, st = ctx.wishart_mean(df2, iS2)
        assert int(st.abs().sum()) == 0 and rel_l2(back_m(EL), R.wishart_mean((pdf, piS))) < 1e-4
    # fused VMP, d = 2 and 3, a few data sets of different size
    for d, N, batch in ((2, 300, 33), (3, 200, 10)):
        ys = []
        for b in range(batch):
            Lc = rng.standard_normal((d, d))
            C = Lc @ Lc.T + 0.1 * np.eye(d)
            ys.append(rng.random(d)[None, :] + rng.standard_normal((N, d)) @ np.linalg.cholesky(C).T)
        y = r32(np.stack(ys, axis=-1))
        ref = vmp.mv_iid_wishart(y, iterations=6)
        got = ctx.mv_iid_wishart_vmp(t(y), iterations=6)
        assert int(got["status"].abs().sum()) == 0
        assert rel_l2(got["m_mean"].cpu().numpy(), ref["m_mean"]) < 1e-5
        assert rel_l2(got["m_cov"].cpu().numpy(), ref["m_cov"]) < 1e-4
        assert rel_l2(got["df"].cpu().numpy(), ref["df"]) < 1e-6
        assert rel_l2(got["inv_scale"].cpu().numpy(), ref["inv_scale"]) < 1e-4
