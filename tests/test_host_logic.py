"""Host-side logic that needs no GPU: sharding arithmetic, gathered-slab assembly, the infer()
keyword surface (unsupported features must raise, never be silently ignored)."""
import numpy as np
import pytest
import torch


def test_shard_bounds_cover_and_partition(rx):
    sh = rx.sharding
    for batch, world in [(10, 3), (65536, 8), (7, 8), (524288, 8)]:
        spans = [sh.shard_bounds(batch, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and b >= a
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_assemble_gathered_layout(rx):
    G, T, d, b = 3, 4, 2, 5
    full = torch.arange(T * d * G * b, dtype=torch.float32).reshape(T, d, G * b)
    slabs = torch.stack([full[..., g * b:(g + 1) * b].contiguous() for g in range(G)])
    assert torch.equal(rx.sharding.assemble_gathered(slabs), full)
    fullc = torch.arange(T * d * d * G * b, dtype=torch.float32).reshape(T, d, d, G * b)
    slabs = torch.stack([fullc[..., g * b:(g + 1) * b].contiguous() for g in range(G)])
    assert torch.equal(rx.sharding.assemble_gathered(slabs), fullc)


def test_infer_rejects_out_of_scope_keywords(rx):
    model = rx.linear_gaussian_ssm_smoothing(np.eye(2), np.eye(2), np.eye(2), np.eye(2), (np.zeros(2), np.eye(2)))
    y = torch.zeros(3, 2, 4)
    for kw in ("callbacks", "constraints", "meta", "predictvars", "annotations"):
        with pytest.raises(NotImplementedError):
            rx.infer(model=model, data={"y": y}, **{kw: object()})
    with pytest.raises(TypeError):
        rx.infer(model=model, data={"y": y}, not_a_keyword=1)
    with pytest.raises(NotImplementedError):
        rx.infer(model=model, data={"y": y}, options={"rulefallback": None})
    with pytest.raises(KeyError):
        rx.infer(model=model, data={"z": y})


def test_infer_fails_loudly_without_gpu_and_honours_catch_exception(rx):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model = rx.linear_gaussian_ssm_smoothing(np.eye(2), np.eye(2), np.eye(2), np.eye(2), (np.zeros(2), np.eye(2)))
    with pytest.raises(Exception):
        rx.infer(model=model, data={"y": torch.zeros(3, 2, 4)})


def test_call_rule_unknown_rule_raises(rx):
    with pytest.raises(rx.RuleMethodError):
        rx.call_rule(None, "Bernoulli", "out", m_p=None)


def test_streaming_engine_host_logic(rx):
    """RxInferenceEngine bookkeeping without a GPU (the device calls are stubbed): lifecycle flags, bounded history
    (circular buffer of `keephistory` ticks over uneven chunks), historyvars validation, exhausted engines."""
    from rxinfer_jl_b200.streaming import RxInferenceEngine

    class StubCtx:
        device = 0

        def __init__(self):
            self.t = 0

        def hgf_filter(self, chunk, init=None, **kw):
            n, b = chunk.shape
            out = (self.t + torch.arange(n, dtype=torch.float32))[:, None, None].expand(n, 4, b).contiguous()
            self.t += n
            return out

        def hgf_filter_chunk(self, chunk, carry, **kw):
            assert carry.shape == (4, chunk.shape[1]) and float(carry[0, 0]) == self.t - 1     # carry = last output row
            return self.hgf_filter(chunk)

    eng = RxInferenceEngine(StubCtx(), rx.hgf(), batch=3, iterations=2, keephistory=5, datastream=None)
    for n in (2, 3, 1, 4, 2):
        eng.push(torch.zeros(n, 3))
    h = eng.history
    assert eng.ticks == 12 and h["xt"].mean().shape == (5, 3)
    assert torch.equal(h["xt"].mean()[:, 0], torch.arange(7, 12, dtype=torch.float32))       # the LAST five ticks
    assert sum(p.m.shape[0] for p in eng._hist["xt"]) < 12                                    # older chunks were dropped
    with pytest.raises(KeyError):
        RxInferenceEngine(StubCtx(), rx.hgf(), batch=3, keephistory=2, historyvars=("nope",))
    with pytest.raises(RuntimeError):
        RxInferenceEngine(StubCtx(), rx.hgf(), batch=3).history                               # keephistory not requested
    done = RxInferenceEngine(StubCtx(), rx.hgf(), batch=3, keephistory=4, datastream=[torch.zeros(2, 3), torch.zeros(5, 3)])
    assert done.is_completed and not done.is_running and done.ticks == 7
    with pytest.raises(RuntimeError):
        done.start()
    with pytest.raises(NotImplementedError):
        RxInferenceEngine(StubCtx(), rx.linear_gaussian_ssm_smoothing(np.eye(2), np.eye(2), np.eye(2), np.eye(2),
                                                                       (np.zeros(2), np.eye(2))), batch=3)
    with pytest.raises(ValueError):
        rx.infer(model=rx.hgf(), data={"y": torch.zeros(2, 3)}, datastream=[torch.zeros(2, 3)])
    with pytest.raises(ValueError):
        rx.infer(model=rx.hgf(), autoupdates="zt_min_mean, zt_min_var = mean_var(q(zt))")      # batch missing


def test_streaming_engine_gamma_model_host_logic(rx):
    """Engine bookkeeping for the streaming Gamma-precision model (device call stubbed): carry = last output row,
    history of both variables, free-energy history averaged over the observations (streaming.jl:12)."""
    from rxinfer_jl_b200.streaming import RxInferenceEngine

    class StubCtx:
        device = 0
        t = 0

        def stream_vmp_gamma(self, chunk, iters, w, init, prev, want_free_energy):
            n, b = chunk.shape
            assert (prev is None) == (self.t == 0)
            if prev is not None:
                assert float(prev[0, 0]) == self.t - 1
            out = (self.t + torch.arange(n, dtype=torch.float32))[:, None, None].expand(n, 4, b).contiguous()
            fe = torch.ones(n, iters, b) * torch.arange(iters, dtype=torch.float32)[None, :, None] if want_free_energy else None
            self.t += n
            return out, fe

    eng = RxInferenceEngine(StubCtx(), rx.kalman_gamma_streaming(), batch=2, iterations=3, keephistory=4, free_energy=True,
                            datastream=[torch.zeros(3, 2), torch.zeros(2, 2)])
    assert eng.is_completed and eng.ticks == 5
    assert eng.history["x_t"].mean().shape == (4, 2) and eng.history["τ"].rate().shape == (4, 2)
    assert torch.equal(eng.free_energy_history, torch.arange(3, dtype=torch.float32)[:, None].expand(3, 2))
    assert float(eng.posteriors["τ"].shape()[0]) == 4.0


def test_padded_shards_and_assembly():
    """Unequal shards (batch % world != 0) are padded to a common slab width for the device gathers and the pad
    columns are dropped again on assembly (unequal counts would hang ncclAllGather)."""
    import torch
    from rxinfer_jl_b200.sharding import assemble_gathered, padded_shard, shard_bounds
    batch, world, T, d = 10, 4, 3, 2
    bp = padded_shard(batch, world)
    assert bp == 3 and sum(hi - lo for lo, hi in (shard_bounds(batch, world, r) for r in range(world))) == batch
    full = torch.arange(T * d * batch, dtype=torch.float32).reshape(T, d, batch)
    g = torch.full((world, T, d, bp), -1.0)
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        g[r, :, :, : hi - lo] = full[:, :, lo:hi]
    assert torch.equal(assemble_gathered(g, batch=batch), full)
    g2 = torch.stack([full[:, :, r * 5:(r + 1) * 5] for r in range(2)])
    assert torch.equal(assemble_gathered(g2), full)
