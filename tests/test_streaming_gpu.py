"""Streaming engine in time-chunks: rxg_lgssm_filter_chunk_f32 / rxg_hgf_filter_chunk_f32
and the RxInferenceEngine mirror.  Oracle: the notebook's per-datum streaming filter
(oracle.lgssm.filter_streaming = /root/reference/benchmarks/...ipynb:107-113,199-216 through @autoupdates)."""
import numpy as np
import pytest
import torch

from oracle import hgf, lgssm
from util import TOL_COV, TOL_MEAN, TOL_NLE, f32_model, rel_l2

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")


def _chunks(T, sizes):
    out, t = [], 0
    i = 0
    while t < T:
        n = min(sizes[i % len(sizes)], T - t)
        out.append((t, t + n)); t += n; i += 1
    return out


@pytest.mark.parametrize("d,batch,sizes", [(4, 200, (1, 7, 64, 3)), (2, 66, (50,)), (4, 19000, (128,))])
def test_chunked_stream_equals_oracle_and_single_call(rx, ctx, d, batch, sizes):
    T = 300
    mod = f32_model(lgssm.notebook_model(d))
    _, y = lgssm.generate_data(mod, T, batch, seed=11)
    ref = lgssm.filter_streaming(y, **mod)
    yd = dev(y)
    kw = dict(A=mod["A"], B=mod["B"], P=mod["P"], Q=mod["Q"])
    whole = ctx.lgssm(yd, **kw, m0=mod["m0"], S0=mod["S0"], smooth=False, transition_first=True, want_evidence=True)
    prev = dev(np.repeat(mod["m0"][:, None], batch, 1))
    carry = np.ascontiguousarray(mod["S0"], dtype=np.float32).copy()
    means, covs, nle = [], [], 0.0
    for a, b in _chunks(T, sizes):
        r = ctx.lgssm_filter_chunk(yd[a:b].contiguous(), **kw, prev_mean=prev, carry_cov=carry, want_evidence=True)
        prev = r["mean"][-1]
        means.append(r["mean"]); covs.append(r["cov"]); nle = nle + r["neg_log_evidence"].double()
        # carry out == filtered covariance of the chunk's last step
        assert np.allclose(carry, r["cov"][-1, :, :, 0].cpu().numpy(), rtol=0, atol=0)
    mean, cov = torch.cat(means).cpu().numpy(), torch.cat(covs).cpu().numpy()
    assert rel_l2(mean, ref["mean"]) < TOL_MEAN and rel_l2(cov, ref["cov"]) < TOL_COV
    assert rel_l2(mean, whole["mean"].cpu().numpy()) < TOL_MEAN
    wn = whole["neg_log_evidence"].double().cpu().numpy()
    assert np.max(np.abs(nle.cpu().numpy() - wn) / np.abs(wn)) < TOL_NLE          # evidence is additive over chunks


def test_chunked_stream_with_offset_and_shared_cov_out(ctx):
    T, batch = 90, 130
    mod = f32_model(lgssm.notebook_model(2))
    u = np.array([0.5, -0.25], np.float32).astype(np.float64)
    _, y = lgssm.generate_data(mod, T, batch, seed=12)
    ref = lgssm.filter_streaming(y, **mod, u=u)
    yd = dev(y)
    prev = dev(np.repeat(mod["m0"][:, None], batch, 1))
    carry = np.ascontiguousarray(mod["S0"], dtype=np.float32).copy()
    means, covs = [], []
    for a, b in _chunks(T, (13, 40)):
        r = ctx.lgssm_filter_chunk(yd[a:b].contiguous(), mod["A"], mod["B"], mod["P"], mod["Q"], prev, carry, u=u,
                                   cov_shared_out=True)
        prev = r["mean"][-1]
        assert r["cov"].dim() == 3
        means.append(r["mean"]); covs.append(r["cov"])
    assert rel_l2(torch.cat(means).cpu().numpy(), ref["mean"]) < TOL_MEAN
    assert rel_l2(torch.cat(covs).cpu().numpy(), ref["cov"][..., 0]) < TOL_COV


@pytest.mark.parametrize("d", [16, 64])
def test_chunked_stream_large_state(ctx, d):
    T, batch = 48, 160
    mod = f32_model(lgssm.dense_model(d))
    _, y = lgssm.generate_data(mod, T, batch, seed=13)
    ref = lgssm.filter_streaming(y, **mod)
    yd = dev(y)
    prev = dev(np.repeat(mod["m0"][:, None], batch, 1))
    carry = np.ascontiguousarray(mod["S0"], dtype=np.float32).copy()
    means = []
    for a, b in _chunks(T, (20, 5)):
        r = ctx.lgssm_filter_chunk(yd[a:b].contiguous(), mod["A"], mod["B"], mod["P"], mod["Q"], prev, carry,
                                   cov_shared_out=True)
        prev = r["mean"][-1].contiguous()
        means.append(r["mean"])
    assert rel_l2(torch.cat(means).cpu().numpy(), ref["mean"]) < TOL_MEAN
    assert rel_l2(carry, ref["cov"][-1, :, :, 0]) < TOL_COV


def test_hgf_chunks_are_bitwise_the_single_call(ctx):
    _, _, y = hgf.generate_data(120, 96, seed=14)
    yd = dev(y)
    whole = ctx.hgf_filter(yd, iters=7)
    outs, prev = [], None
    for a, b in _chunks(120, (1, 30, 9)):
        o = ctx.hgf_filter(yd[a:b].contiguous(), iters=7) if prev is None else ctx.hgf_filter_chunk(yd[a:b].contiguous(), prev, iters=7)
        prev = o[-1].contiguous()
        outs.append(o)
    assert torch.equal(torch.cat(outs), whole)


def test_engine_mirror(rx, ctx):
    """infer(model=filtering, datastream=..., keephistory=...) -> RxInferenceEngine (streaming.jl:536-845)."""
    T, batch, d = 64, 80, 4
    mod = f32_model(lgssm.notebook_model(d))
    _, y = lgssm.generate_data(mod, T, batch, seed=15)
    ref = lgssm.filter_streaming(y, **mod)
    yd = dev(y)
    model = rx.linear_gaussian_ssm_filtering(mod["A"], mod["B"], mod["P"], mod["Q"], (mod["m0"], mod["S0"]))
    eng = rx.infer(model=model, datastream=(yd[a:b].contiguous() for a, b in _chunks(T, (10, 6))), batch=batch,
                   keephistory=T, free_energy=True, context=ctx)
    assert eng.is_completed and not eng.is_running and eng.ticks == T
    h = eng.history["x_t"]
    assert h.mean().shape == (T, d, batch) and rel_l2(h.mean().cpu().numpy(), ref["mean"]) < TOL_MEAN
    assert rel_l2(h.cov().cpu().numpy(), ref["cov"]) < TOL_COV
    whole = ctx.lgssm(yd, mod["A"], mod["B"], mod["P"], mod["Q"], mod["m0"], mod["S0"], smooth=False,
                      transition_first=True, want_evidence=True)["neg_log_evidence"].double()
    assert torch.max(torch.abs(eng.free_energy_history.double().sum(0) - whole) / whole.abs()) < TOL_NLE
    with pytest.raises(RuntimeError):
        eng.start()                                       # exhausted engine (streaming.jl:188-191)
    # push-driven engine with a bounded history (circular buffer, keephistory < ticks)
    eng2 = rx.infer(model=model, autoupdates="x_min_t_mean, x_min_t_cov = mean_cov(q(x_t))", batch=batch, keephistory=8,
                    context=ctx)
    for a, b in _chunks(T, (5,)):
        eng2.push(yd[a:b].contiguous())
    assert eng2.history["x_t"].mean().shape[0] == 8
    assert rel_l2(eng2.history["x_t"].mean().cpu().numpy(), ref["mean"][-8:]) < TOL_MEAN
    with pytest.raises(KeyError):
        rx.infer(model=model, autoupdates="...", batch=batch, keephistory=3, historyvars=("nope",), context=ctx)
    # HGF engine == the batch call of the reference test (data + autoupdates + keephistory)
    _, _, yh = hgf.generate_data(40, 64, seed=16)
    eh = rx.infer(model=rx.hgf(), datastream=[dev(yh[:15]), dev(yh[15:])], batch=64, iterations=5, keephistory=40, context=ctx)
    oh = ctx.hgf_filter(dev(yh), iters=5)
    assert torch.equal(eh.history["xt"].mean(), oh[:, 0]) and torch.equal(eh.history["zt"].var(), oh[:, 3])
