#!/usr/bin/env python
"""Secondary measurements (not the driver's bench contract): the other BASELINE.json configs and
the per-rule kernels, one JSON line each.  Device-resident data, CUDA events, >= 3 warm-ups.

  python bench_extra.py [--which per_chain,filter,hgf,rules,vmp,scaling_T]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import rxinfer_jl_b200 as rx  # noqa: E402
from bench import dense_model_f32, notebook_model_d2_f32, notebook_model_f32, peaks  # noqa: E402


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="per_chain,filter,hgf,rules,vmp,scaling_T,large,stream,round2")
    args = ap.parse_args()
    which = set(args.which.split(","))
    ctx = rx.Context(0)
    peak, _ = peaks()
    mod = notebook_model_f32()
    kw = dict(A=mod["A"], B=mod["B"], P=mod["P"], Q=mod["Q"], m0=mod["m0"], S0=mod["S0"])
    T, batch = 1000, 65536
    g = torch.Generator(device="cuda").manual_seed(7)

    if which & {"per_chain", "filter"}:
        y = torch.randn(T, 4, batch, device="cuda", generator=g) * 3.3
        mean = torch.empty(T, 4, batch, device="cuda"); cov = torch.empty(T, 4, 4, batch, device="cuda")
        if "per_chain" in which:
            ms = timed(lambda: ctx.lgssm(y, **kw, smooth=True, out_mean=mean, out_cov=cov, force_per_chain_path=True))
            print(json.dumps({"what": "lgssm smooth, per-chain covariance recursion (lgssm_chain_kernel)", "d": 4, "T": T,
                              "batch": batch, "ms": ms, "messages_per_s": 6 * T * batch / ms * 1e3,
                              "algorithmic_GBs": 96 * T * batch / ms / 1e6, "frac_of_hbm_peak": 96 * T * batch / ms / 1e6 / peak}))
        if "filter" in which:
            for tf in (False, True):
                ms = timed(lambda: ctx.lgssm(y, **kw, smooth=False, out_mean=mean, out_cov=cov, transition_first=tf))
                print(json.dumps({"what": "lgssm filter (forward half), gain-table path", "transition_first": tf, "d": 4, "T": T,
                                  "batch": batch, "ms": ms, "messages_per_s": 4 * T * batch / ms * 1e3,
                                  "algorithmic_GBs": 96 * T * batch / ms / 1e6}))
        del y, mean, cov

    if "scaling_T" in which:
        # the notebook's scaling table (ipynb:795-806) at d = 2, batched: T from 50 to 50 000
        m2 = notebook_model_d2_f32()
        for TT in (50, 1000, 10000, 50000):
            b = max(1024, min(65536, (1 << 26) // TT))
            y = torch.randn(TT, 2, b, device="cuda", generator=g) * 3.3
            ms = timed(lambda: ctx.lgssm(y, **m2, smooth=True), warm=3, reps=3)
            print(json.dumps({"what": "lgssm smooth d=2 (notebook scaling table shape)", "T": TT, "batch": b, "ms": ms,
                              "messages_per_s": 6 * TT * b / ms * 1e3, "ms_per_chain_equiv": ms / b}))
            del y

    if "stream" in which:
        # what HBM delivers to a dependency-free streaming kernel for a given read : write mix (the sweep kernel moves
        # 2.14 GB in and 5.25 GB out per launch, i.e. ~2 : 5)
        n = 1 << 28                                            # 1 GiB per row: far beyond L2
        for nr, nw in ((1, 1), (2, 5), (0, 4), (4, 1)):
            src = torch.randn(max(nr, 1), n, device="cuda")[:nr] if nr else torch.empty(0, n, device="cuda")
            dst = torch.empty(nw, n, device="cuda")
            ms = timed(lambda: ctx.selftest_stream(src, dst), warm=3, reps=5)
            print(json.dumps({"what": "stream_mix_kernel (HBM yardstick)", "rows_read": nr, "rows_written": nw, "ms": ms,
                              "GBs": (nr + nw) * n * 4 / ms / 1e6, "frac_of_copy_peak": (nr + nw) * n * 4 / ms / 1e6 / peak}))
            del src, dst

    if "round2" in which:
        # round-2 additions: shared vs per-chain missing-data pattern, embedded shapes, the generic one-CTA-per-chain kernel,
        # the IID Wishart VMP
        y = torch.randn(T, 4, batch, device="cuda", generator=g) * 3.3
        mean = torch.empty(T, 4, batch, device="cuda"); cov = torch.empty(T, 4, 4, batch, device="cuda")
        tm = (np.random.default_rng(1).random(T) > 0.2).astype(np.uint8)
        full = torch.as_tensor(np.repeat(tm[:, None], batch, axis=1), device="cuda")
        for name, mk in (("shared pattern, RXG_MASK_SHARED (gain-table path)", tm), ("same pattern as a per-chain mask (per-chain covariance recursion)", full)):
            ms = timed(lambda: ctx.lgssm(y, **kw, smooth=True, out_mean=mean, out_cov=cov, mask=mk))
            print(json.dumps({"what": "lgssm smooth with 20 % missing steps: " + name, "d": 4, "T": T, "batch": batch, "ms": ms,
                              "messages_per_s": 6 * T * batch / ms * 1e3, "frac_of_hbm_peak": 96 * T * batch / ms / 1e6 / peak}))
        del y, mean, cov, full
        rng = np.random.default_rng(5)
        for d, m, b in ((5, 3, 65536), (6, 6, 65536), (12, 7, 16384), (16, 16, 16384)):
            Aq, _ = np.linalg.qr(rng.standard_normal((d, d)))
            md = {k: v.astype(np.float32) for k, v in dict(A=0.95 * Aq, B=rng.standard_normal((m, d)) / np.sqrt(d), P=0.2 * np.eye(d),
                                                          Q=1.5 * np.eye(m), m0=np.zeros(d), S0=5.0 * np.eye(d)).items()}
            y = torch.randn(T, m, b, device="cuda", generator=g)
            mean = torch.empty(T, d, b, device="cuda"); cov = torch.empty(T, d, d, b, device="cuda")
            ms = timed(lambda: ctx.lgssm(y, **md, smooth=True, out_mean=mean, out_cov=cov), warm=2, reps=3)
            print(json.dumps({"what": "lgssm smooth, shared model, general shape" + (" (native)" if (d, m) in ((6, 6), (16, 16)) else " (embedded in the next native shape)"),
                              "d": d, "m": m, "T": T, "batch": b, "ms": ms, "messages_per_s": 6 * T * b / ms * 1e3,
                              "algorithmic_GBs": 4 * (m + d + d * d) * T * b / ms / 1e6}))
            del y, mean, cov
        for d, b in ((16, 2048), (64, 512)):
            md = dense_model_f32(d)
            y = torch.randn(T, d, b, device="cuda", generator=g) * 3.3
            mk = (torch.rand(T, b, device="cuda", generator=g) > 0.2).to(torch.uint8)
            mean = torch.empty(T, d, b, device="cuda"); cov = torch.empty(T, d, d, b, device="cuda")
            ms = timed(lambda: ctx.lgssm(y, **md, smooth=True, out_mean=mean, out_cov=cov, mask=mk), warm=1, reps=2)
            print(json.dumps({"what": "lgssm smooth, per-chain missing data, generic one-CTA-per-chain kernel (CUDA cores)", "d": d, "T": T,
                              "batch": b, "ms": ms, "messages_per_s": 6 * T * b / ms * 1e3, "us_per_chain_step": ms * 1e3 / (T * b) * min(b, 148 * (2 if d <= 32 else 1))}))
            del y, mk, mean, cov
        yw = torch.randn(1500, 2, 32768, device="cuda", generator=g)
        ms = timed(lambda: ctx.mv_iid_wishart_vmp(yw, iterations=10), warm=2, reps=3)
        print(json.dumps({"what": "IID Wishart-precision VMP (mv_iid_precision model), 10 iterations", "d": 2, "N": 1500, "batch": 32768, "ms": ms,
                          "datasets_per_s": 32768 / ms * 1e3, "GBs": yw.numel() * 4 / ms / 1e6}))
        del yw
        # latent AR (lar_tests.jl model): T = 500, 15 structured-VMP iterations = 15 filter + RTS passes per series
        for order, bl in ((1, 65536), (5, 16384)):
            yl = torch.randn(500, bl, device="cuda", generator=g)
            ms = timed(lambda: ctx.lar_vmp(yl, order, 5.0, iterations=15), warm=1, reps=3)
            print(json.dumps({"what": "latent AR structured VMP (lar_tests.jl model), 15 iterations", "order": order, "T": 500, "batch": bl,
                              "ms": ms, "series_per_s": bl / ms * 1e3, "chain_steps_per_s": 2 * 15 * 500 * bl / ms * 1e3}))
            del yl

    if "large" in which:
        # BASELINE configs[2] (d = 64, T = 1000, batch = 4096) and the smaller tensor-core sizes; shared model.
        # flops: textbook Kalman + RTS mean recursions only = 2 * (2 d^2 [F x + K y] + 2 d^2 [E x + G x]) per (chain, step)
        ctx.set_profiling(True)
        for d, b in ((64, 4096), (64, 18944), (32, 16384), (16, 65536)):
            md = dense_model_f32(d)
            y = torch.randn(T, d, b, device="cuda", generator=g) * 3.3
            mean = torch.empty(T, d, b, device="cuda")
            for no_umma in ("0", "1"):
                ctx.set_option("no_umma", int(no_umma))
                sw, gn = [], []
                def run():
                    ctx.lgssm(y, **md, smooth=True, out_mean=mean, cov_shared_out=True)
                    a, bb = ctx.profile_last_ms(); sw.append(a); gn.append(bb)
                ms = timed(run, warm=2, reps=3)
                print(json.dumps({"what": "lgssm smooth, large-state family (shared model, cov de-duplicated)", "d": d, "T": T, "batch": b,
                                  "sweep": "tcgen05 3xTF32 (umma_ky + lgssm_umma_sweep)" if no_umma == "0" else "FP32 pipe (lgssm_block_sweep)",
                                  "ms": ms, "sweep_ms": float(np.mean(sw[-3:])), "gain_tables_ms": float(np.mean(gn[-3:])),
                                  "messages_per_s": 6 * T * b / ms * 1e3,
                                  "sweep_TFLOPs": 8 * d * d * T * b / (float(np.mean(sw[-3:])) * 1e-3) / 1e12}))
            ctx.set_option("no_umma", 0)
            del y, mean
        ctx.set_profiling(False)

    if "hgf" in which:
        Th, bh, iters = 1000, 32768, 20
        yh = torch.randn(Th, bh, device="cuda", generator=g).cumsum(0) * 0.5
        out = torch.empty(Th, 4, bh, device="cuda")
        ms = timed(lambda: ctx.hgf_filter(yh, iters=iters, out=out), warm=3, reps=3)
        n_exp = (31 + 1 + iters * 32) * Th * bh
        print(json.dumps({"what": "HGF filter (BASELINE configs[3]): GCV node, GH-31, 20 VMP iterations", "T": Th, "batch": bh,
                          "iters": iters, "ms": ms, "vmp_iterations_per_s": iters * Th * bh / ms * 1e3,
                          "messages_per_s": 6 * iters * Th * bh / ms * 1e3, "exp_per_s": n_exp / ms * 1e3,
                          "io_GBs": 20 * Th * bh / ms / 1e6}))

    if "vmp" in which:
        yv = torch.randn(1000, 65536, device="cuda", generator=g).cumsum(0)
        ms = timed(lambda: ctx.lgssm_vmp_gamma(yv, iterations=10), warm=2, reps=3)
        print(json.dumps({"what": "Gamma-precision VMP around scalar smoother, 10 iterations", "T": 1000, "batch": 65536, "ms": ms,
                          "sweeps_per_s": 10 / ms * 1e3}))

    if "rules" in which:
      for n, d in ((1 << 22, 4), (1 << 16, 16), (1 << 14, 64)):       # d = 16 / 64: csrc/rxg_rules_large.cu (configs[2] message size)
        mu = torch.randn(d, n, device="cuda", generator=g)
        X = torch.randn(d, d, n, device="cuda", generator=g)
        S = torch.einsum("ikn,jkn->ijn", X, X).contiguous() + d * torch.eye(d, device="cuda")[:, :, None]
        S = S.contiguous()
        del X
        A = np.asarray(mod["A"]) if d == 4 else np.asarray(dense_model_f32(d)["A"])
        Pm = np.asarray(mod["P"]) if d == 4 else np.eye(d, dtype=np.float32)
        rows = []
        rows.append(("MvNormalMeanCovariance(:out)  (mu, S + Sigma)", timed(lambda: ctx.rule_add_cov(mu, S, Pm)), 2 * (d + d * d) * 4))
        rows.append(("*(:out)  (A mu, A S A')", timed(lambda: ctx.rule_mul_out(A, mu, S)), 2 * (d + d * d) * 4))
        rows.append(("*(:in)   cholinv + A' W A", timed(lambda: ctx.rule_mul_in(A, mu, S)), 2 * (d + d * d) * 4 + 4))
        mu2, S2 = (mu * 0.5).contiguous(), (S * 2.0).contiguous()          # distinct operands: 2 reads + 1 write per message
        rows.append(("prod (xi1 + xi2, W1 + W2)", timed(lambda: ctx.prod_gaussian(mu, S, mu2, S2)), 3 * (d + d * d) * 4))
        rows.append(("mean_cov <-> weightedmean_precision (cholinv)", timed(lambda: ctx.meancov_to_wmp(mu, S)), 2 * (d + d * d) * 4 + 4))
        for name, ms, bytes_per in rows:
            # note: torch.empty_like allocations are inside the timed call (caching allocator)
            print(json.dumps({"what": "rule kernel: " + name, "n": n, "d": d, "ms": ms, "messages_per_s": n / ms * 1e3,
                              "GBs": bytes_per * n / ms / 1e6, "frac_of_hbm_peak": bytes_per * n / ms / 1e6 / peak}))


if __name__ == "__main__":
    main()
