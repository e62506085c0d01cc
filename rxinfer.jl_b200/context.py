"""Thin object wrapper over the C ABI: one ``Context`` per GPU / host thread.

PyTorch is used only as plumbing (device memory, the current stream); every computation goes
through ``librxgauss.so``.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p

import numpy as np
import torch

from . import _lib as L


def _fp(t):
    if t is None:
        return L.as_fp(0)
    return L.as_fp(t.data_ptr())


def _model32(M):
    """Shared model matrices are small host arrays (row-major fp32)."""
    a = np.ascontiguousarray(np.asarray(M, dtype=np.float32))
    return a, a.ctypes.data_as(L.fp)


class Context:
    """``rxg_ctx`` bound to a CUDA device.  Launches go to torch's current stream on that device
    (so ``torch.cuda.Event`` timing and stream ordering with torch ops are meaningful)."""

    def __init__(self, device: int | None = None, use_torch_stream: bool = True):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.RxGaussError(L.RXG_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = c_void_p()
        rc = self.lib.rxg_create(ctypes.byref(h), self.device, 0)
        if rc != 0:
            raise L.RxGaussError(rc, "rxg_create failed")
        self.h = h
        self._stream = None
        if use_torch_stream:
            self.bind_stream()

    # ------------------------------------------------------------------ plumbing
    def bind_stream(self, stream: torch.cuda.Stream | None = None):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._stream = s
        self._check(self.lib.rxg_set_stream(self.h, c_void_p(s.cuda_stream)))

    def _check(self, rc):
        if rc != 0:
            raise L.RxGaussError(rc, self.lib.rxg_last_error(self.h).decode())

    def sync(self):
        self._check(self.lib.rxg_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.rxg_launch_count(self.h))

    def host_fill_threads(self) -> int:
        return int(self.lib.rxg_host_fill_threads())

    OPTIONS = {"gain_seq": 0, "large_seq": 1, "no_umma": 2, "sweep_variant": 3, "force_cpt": 4, "host_threads": 5,
               "host_cov_d2h": 6, "host_bcast_min_mb": 7, "host_slices": 8, "gather_mode": 9}

    def set_option(self, name: str, value: int):
        """``rxg_set_option``: per-context dispatch switches (cross-check kernels, host-pipeline tuning)."""
        self._check(self.lib.rxg_set_option(self.h, self.OPTIONS[name], int(value)))

    def get_option(self, name: str) -> int:
        v = ctypes.c_longlong()
        self._check(self.lib.rxg_get_option(self.h, self.OPTIONS[name], ctypes.byref(v)))
        return int(v.value)

    def set_profiling(self, on=True):
        self._check(self.lib.rxg_set_profiling(self.h, 1 if on else 0))

    def profile_last_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        self._check(self.lib.rxg_profile_last_ms(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.rxg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, *ts):
        for t in ts:
            if t is None:
                continue
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("expected contiguous float32 CUDA tensors")
            if t.device.index != self.device:
                raise ValueError(f"tensor lives on cuda:{t.device.index}, this context is bound to cuda:{self.device}")

    def _io(self, t, name, on_dev, dtype=torch.float32, shape=None):
        """Validate one I/O array of a fused sweep: dtype, contiguity, device (or host), shape."""
        if t is None:
            return
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous {dtype} tensor, got {t.dtype}, "
                             f"contiguous={t.is_contiguous()} (call .contiguous() / .to({dtype}) explicitly)")
        if on_dev:
            if not t.is_cuda or t.device.index != self.device:
                raise ValueError(f"{name}: expected a tensor on cuda:{self.device} (y is a device array), got {t.device}")
        elif t.is_cuda:
            raise ValueError(f"{name}: y is a host array, so every data array must be on the host; got {t.device}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=f"cuda:{self.device}")

    # ------------------------------------------------------------------ fused sweeps
    def lgssm(self, y, A, B, P, Q, m0, S0, *, u=None, smooth=True, mask=None, want_cov=True, want_evidence=False,
              want_status=False, per_chain_model=False, force_per_chain_path=False, cov_shared_out=False,
              transition_first=False, out_mean=None, out_cov=None, out_status=None, asynchronous=False):
        """y[T, m, batch] (CUDA fp32, or pinned/pageable CPU fp32 for the host-pointer path)
        -> dict(mean[T,d,batch], cov[T,d,d,batch] or [T,d,d], neg_log_evidence[batch], status[batch])."""
        on_dev = y.is_cuda
        if y.dim() != 3:
            raise ValueError(f"y: expected [T, m, batch], got shape {tuple(y.shape)}")
        T, m, batch = y.shape
        self._io(y, "y", on_dev)
        shared_mask = None
        if mask is not None and getattr(mask, "ndim", 2) == 1:
            # one missing-data pattern for every chain (RXG_MASK_SHARED): a host array [T]; the call stays on the gain-table path
            shared_mask = np.ascontiguousarray(np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask, dtype=np.uint8))
            if shared_mask.shape != (T,):
                raise ValueError(f"shared mask: expected shape ({T},), got {shared_mask.shape}")
            mask = None
        self._io(mask, "mask", on_dev, dtype=torch.uint8, shape=(T, batch))
        flags = L.PTR_DEVICE if on_dev else 0
        if shared_mask is not None:
            flags |= L.MASK_SHARED
        if per_chain_model:
            flags |= L.MODEL_PER_CHAIN
            d = A.shape[0]
            self._dev(A, B, P, Q, m0, S0, u)
            ptrs = [_fp(x) for x in (A, B, P, Q, m0, S0, u)]
            keep = None
        else:
            d = np.asarray(A).shape[-1]
            keep = [_model32(x) for x in (A, B, P, Q, m0, S0)]
            ptrs = [k[1] for k in keep]
            if u is not None:
                keep.append(_model32(u))
                ptrs.append(keep[-1][1])
            else:
                ptrs.append(L.as_fp(0))
        if force_per_chain_path:
            flags |= L.PATH_PER_CHAIN
        if cov_shared_out:
            flags |= L.COV_SHARED_OUT
        if transition_first:
            flags |= L.TRANSITION_FIRST
        if asynchronous:
            flags |= L.ASYNC
        mk = lambda *s, dt=torch.float32: (torch.empty(*s, dtype=dt, device=y.device) if on_dev
                                           else torch.empty(*s, dtype=dt).pin_memory())
        self._io(out_mean, "out_mean", on_dev, shape=(T, d, batch))
        self._io(out_cov, "out_cov", on_dev, shape=(T, d, d) if cov_shared_out else (T, d, d, batch))
        mean = out_mean if out_mean is not None else mk(T, d, batch)
        need_cov = want_cov or per_chain_model or force_per_chain_path or mask is not None
        cov = out_cov
        if cov is None and need_cov:
            cov = mk(T, d, d) if cov_shared_out else mk(T, d, d, batch)
        nle = mk(batch) if want_evidence else None
        self._io(out_status, "out_status", on_dev, dtype=torch.int32, shape=(batch,))
        status = out_status if out_status is not None else (mk(batch, dt=torch.int32) if want_status else None)
        fn = self.lib.rxg_lgssm_smooth_f32 if smooth else self.lib.rxg_lgssm_filter_f32
        mask_p = ctypes.cast(c_void_p(mask.data_ptr()), L.u8p) if mask is not None else ctypes.cast(c_void_p(None), L.u8p)
        if shared_mask is not None:
            mask_p = shared_mask.ctypes.data_as(L.u8p)
        st_p = ctypes.cast(c_void_p(status.data_ptr()), L.i32p) if status is not None else ctypes.cast(c_void_p(None), L.i32p)
        self._check(fn(self.h, d, m, T, batch, *ptrs, _fp(y), mask_p, _fp(mean), _fp(cov), _fp(nle), st_p, flags))
        return dict(mean=mean, cov=cov if (want_cov or need_cov) else None, neg_log_evidence=nle, status=status)

    def lgssm_filter_chunk(self, y, A, B, P, Q, prev_mean, carry_cov, *, u=None, want_evidence=False,
                           cov_shared_out=False, out_mean=None, out_cov=None):
        """One time-chunk of the streaming engine (``rxg_lgssm_filter_chunk_f32``).  ``prev_mean[d, batch]``
        (CUDA) and ``carry_cov[d, d]`` (host fp32 numpy, updated IN PLACE) are the autoupdate carry."""
        self._dev(y, prev_mean)
        T, m, batch = y.shape
        d = prev_mean.shape[0]
        if not (isinstance(carry_cov, np.ndarray) and carry_cov.dtype == np.float32 and carry_cov.shape == (d, d)
                and carry_cov.flags.c_contiguous):
            raise ValueError("carry_cov must be a C-contiguous float32 numpy array of shape (d, d)")
        keep = [_model32(x) for x in (A, B, P, Q)]
        ptrs = [k[1] for k in keep]
        if u is not None:
            keep.append(_model32(u))
            ptrs.append(keep[-1][1])
        else:
            ptrs.append(L.as_fp(0))
        mean = out_mean if out_mean is not None else self.empty(T, d, batch)
        cov = out_cov if out_cov is not None else (self.empty(T, d, d) if cov_shared_out else self.empty(T, d, d, batch))
        nle = self.empty(batch) if want_evidence else None
        flags = L.PTR_DEVICE | (L.COV_SHARED_OUT if cov_shared_out else 0)
        cc = carry_cov.ctypes.data_as(L.fp)
        self._check(self.lib.rxg_lgssm_filter_chunk_f32(self.h, d, m, T, batch, *ptrs, _fp(prev_mean), cc, _fp(y),
                                                        _fp(mean), _fp(cov), _fp(nle), flags))
        return dict(mean=mean, cov=cov, neg_log_evidence=nle)

    def hgf_filter_chunk(self, y, prev, iters=20, kappa=1.0, omega=0.0, z_variance=0.04, y_variance=0.01, out=None,
                         want_free_energy=False):
        """HGF datastream chunk; ``prev[4, batch]`` = ``out[-1]`` of the previous chunk."""
        return self.hgf_filter(y, iters, kappa, omega, z_variance, y_variance, init=None, out=out, prev=prev,
                               want_free_energy=want_free_energy)

    def hgf_filter(self, y, iters=20, kappa=1.0, omega=0.0, z_variance=0.04, y_variance=0.01,
                   init=(0.0, 5.0, 0.0, 5.0), out=None, prev=None, want_free_energy=False):
        """``rxg_hgf_filter_fe_f32``: y[T, batch] -> out[T, 4, batch] = (m_x, v_x, m_z, v_z); with
        ``want_free_energy`` also the Bethe free energy [T, iters, batch] (returned as a pair)."""
        self._dev(y, prev, out)
        T, batch = y.shape
        out = out if out is not None else self.empty(T, 4, batch)
        fe = self.empty(T, iters, batch) if want_free_energy else None
        ini = ctypes.cast((ctypes.c_float * 4)(*init), L.fp) if prev is None else L.as_fp(0)
        self._check(self.lib.rxg_hgf_filter_fe_f32(self.h, T, batch, iters, kappa, omega, z_variance, y_variance,
                                                   ini, _fp(prev), _fp(y), _fp(out), _fp(fe), L.PTR_DEVICE))
        return (out, fe) if want_free_energy else out

    def stream_vmp_gamma(self, y, iters=4, w=1.0, init=(0.0, 1e3, 1.0, 1.0), prev=None, want_free_energy=False):
        """Streaming mean-field VMP with a Gamma observation precision (``rxg_stream_vmp_gamma_f32``);
        y[T, batch] -> out[T, 4, batch] = (m_x, v_x, shape, rate), free energy [T, iters, batch] or None."""
        self._dev(y, prev)
        T, batch = y.shape
        out = self.empty(T, 4, batch)
        fe = self.empty(T, iters, batch) if want_free_energy else None
        ini = (ctypes.c_float * 4)(*init)
        self._check(self.lib.rxg_stream_vmp_gamma_f32(self.h, T, batch, iters, w, ctypes.cast(ini, L.fp), _fp(prev), _fp(y),
                                                      _fp(out), _fp(fe), L.PTR_DEVICE))
        return out, fe

    def lgssm_vmp_gamma(self, y, iterations=10, a=1.0, v_proc=1.0, prior=(0.0, 100.0), gamma_prior=(1.0, 1.0),
                        init_E_tau=1.0, want_free_energy=False):
        self._dev(y)
        T, batch = y.shape
        pm, pv = self.empty(T, batch), self.empty(T, batch)
        sh, rt = self.empty(batch), self.empty(batch)
        fe = self.empty(iterations, batch) if want_free_energy else None
        self._check(self.lib.rxg_lgssm_vmp_gamma_fe_f32(self.h, T, batch, iterations, a, v_proc, prior[0], prior[1],
                                                        gamma_prior[0], gamma_prior[1], init_E_tau, _fp(y), _fp(pm),
                                                        _fp(pv), _fp(sh), _fp(rt), _fp(fe), L.PTR_DEVICE))
        return dict(mean=pm, var=pv, shape=sh, rate=rt, free_energy=fe)

    # ------------------------------------------------------------------ per-rule kernels
    def _mat(self, M):
        """PointMass matrix operand: host array => shared; CUDA tensor [r,c,n] => per message."""
        if isinstance(M, torch.Tensor) and M.is_cuda:
            self._dev(M)
            return M, _fp(M), 0
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(M, dtype=np.float32)), device=f"cuda:{self.device}")
        return t, _fp(t), 1

    def rule_add_cov(self, mu, S, Sigma, which="out"):
        self._dev(mu, S)
        d, n = mu.shape
        keep, Sp, shared = self._mat(Sigma)
        mo, So = torch.empty_like(mu), torch.empty_like(S)
        fn = self.lib.rxg_rule_mvnormal_meancov_out_f32 if which == "out" else self.lib.rxg_rule_mvnormal_meancov_mean_f32
        self._check(fn(self.h, n, d, _fp(mu), _fp(S), Sp, shared, _fp(mo), _fp(So), L.PTR_DEVICE))
        return mo, So

    def rule_mean_from_data(self, y, Sigma):
        self._dev(y)
        d, n = y.shape
        keep, Sp, shared = self._mat(Sigma)
        mo, So = torch.empty_like(y), self.empty(d, d, n)
        self._check(self.lib.rxg_rule_mvnormal_meancov_mean_data_f32(self.h, n, d, _fp(y), Sp, shared, _fp(mo), _fp(So), L.PTR_DEVICE))
        return mo, So

    def rule_mul_out(self, A, mu, S):
        self._dev(mu, S)
        di, n = mu.shape
        keep, Ap, shared = self._mat(A)
        do = keep.shape[0]
        mo, So = self.empty(do, n), self.empty(do, do, n)
        self._check(self.lib.rxg_rule_mul_out_f32(self.h, n, do, di, Ap, shared, _fp(mu), _fp(S), _fp(mo), _fp(So), L.PTR_DEVICE))
        return mo, So

    def rule_mul_in(self, A, mu_out, S_out):
        self._dev(mu_out, S_out)
        do, n = mu_out.shape
        keep, Ap, shared = self._mat(A)
        di = keep.shape[1]
        xi, W = self.empty(di, n), self.empty(di, di, n)
        st = self.empty(n, dtype=torch.int32)
        self._check(self.lib.rxg_rule_mul_in_f32(self.h, n, do, di, Ap, shared, _fp(mu_out), _fp(S_out), _fp(xi), _fp(W),
                                                 ctypes.cast(c_void_p(st.data_ptr()), L.i32p), L.PTR_DEVICE))
        return xi, W, st

    def _pair(self, fn, a, Sa, b, Sb):
        self._dev(a, Sa, b, Sb)
        d, n = a.shape
        o, So = torch.empty_like(a), torch.empty_like(Sa)
        self._check(fn(self.h, n, d, _fp(a), _fp(Sa), _fp(b), _fp(Sb), _fp(o), _fp(So), L.PTR_DEVICE))
        return o, So

    def rule_add_out(self, mu1, S1, mu2, S2):
        return self._pair(self.lib.rxg_rule_add_out_f32, mu1, S1, mu2, S2)

    def rule_add_in(self, mu_out, S_out, mu_other, S_other):
        return self._pair(self.lib.rxg_rule_add_in_f32, mu_out, S_out, mu_other, S_other)

    def prod_gaussian(self, xi1, W1, xi2, W2):
        return self._pair(self.lib.rxg_prod_gaussian_f32, xi1, W1, xi2, W2)

    def _conv(self, fn, v, M):
        self._dev(v, M)
        d, n = v.shape
        vo, Mo = torch.empty_like(v), torch.empty_like(M)
        st = self.empty(n, dtype=torch.int32)
        self._check(fn(self.h, n, d, _fp(v), _fp(M), _fp(vo), _fp(Mo), ctypes.cast(c_void_p(st.data_ptr()), L.i32p), L.PTR_DEVICE))
        return vo, Mo, st

    def meancov_to_wmp(self, mu, S):
        return self._conv(self.lib.rxg_meancov_to_wmp_f32, mu, S)

    def wmp_to_meancov(self, xi, W):
        return self._conv(self.lib.rxg_wmp_to_meancov_f32, xi, W)

    def marginal_gaussian(self, msgs):
        d, n = msgs[0][0].shape
        k = len(msgs)
        for xi, W in msgs:
            self._dev(xi, W)
        xs = (L.fp * k)(*[_fp(x) for x, _ in msgs])
        ws = (L.fp * k)(*[_fp(w) for _, w in msgs])
        mu, S = self.empty(d, n), self.empty(d, d, n)
        st = self.empty(n, dtype=torch.int32)
        self._check(self.lib.rxg_marginal_gaussian_f32(self.h, n, d, k, xs, ws, _fp(mu), _fp(S),
                                                       ctypes.cast(c_void_p(st.data_ptr()), L.i32p), L.PTR_DEVICE))
        return mu, S, st

    def _six(self, fn, a, b, c, d_):
        self._dev(a, b, c, d_)
        n = a.numel()
        o1, o2 = torch.empty_like(a), torch.empty_like(a)
        self._check(fn(self.h, n, _fp(a), _fp(b), _fp(c), _fp(d_), _fp(o1), _fp(o2), L.PTR_DEVICE))
        return o1, o2

    def rule_normal_precision_tau(self, m_out, v_out, m_mu, v_mu):
        return self._six(self.lib.rxg_rule_normal_precision_tau_f32, m_out, v_out, m_mu, v_mu)

    def rule_normal_precision_out(self, m_mu, v_mu, shape, rate):
        return self._six(self.lib.rxg_rule_normal_precision_out_f32, m_mu, v_mu, shape, rate)

    def rule_normal_precision_tau_joint(self, m_joint, V_joint):
        """Structured tau rule: q(out, mu) jointly Gaussian, m_joint[2, n], V_joint[2, 2, n]."""
        self._dev(m_joint, V_joint)
        n = m_joint.shape[-1]
        sh, rt = self.empty(n), self.empty(n)
        self._check(self.lib.rxg_rule_normal_precision_tau_joint_f32(self.h, n, _fp(m_joint), _fp(V_joint), _fp(sh), _fp(rt), L.PTR_DEVICE))
        return sh, rt

    def rule_mvnormal_precision_lambda(self, m_out, V_out, m_mu, V_mu):
        self._dev(m_out, V_out, m_mu, V_mu)
        d, n = m_out.shape
        df, iS = self.empty(n), self.empty(d, d, n)
        self._check(self.lib.rxg_rule_mvnormal_precision_lambda_f32(self.h, n, d, _fp(m_out), _fp(V_out), _fp(m_mu), _fp(V_mu),
                                                                    _fp(df), _fp(iS), L.PTR_DEVICE))
        return df, iS

    def prod_wishart(self, df1, iS1, df2, iS2):
        self._dev(df1, iS1, df2, iS2)
        d, n = iS1.shape[0], iS1.shape[-1]
        df, iS = self.empty(n), self.empty(d, d, n)
        self._check(self.lib.rxg_prod_wishart_f32(self.h, n, d, _fp(df1), _fp(iS1), _fp(df2), _fp(iS2), _fp(df), _fp(iS), L.PTR_DEVICE))
        return df, iS

    def wishart_mean(self, df, iS):
        self._dev(df, iS)
        d, n = iS.shape[0], iS.shape[-1]
        out = self.empty(d, d, n)
        st = self.empty(n, dtype=torch.int32)
        self._check(self.lib.rxg_wishart_mean_f32(self.h, n, d, _fp(df), _fp(iS), _fp(out),
                                                  ctypes.cast(c_void_p(st.data_ptr()), L.i32p), L.PTR_DEVICE))
        return out, st

    def mv_iid_wishart_vmp(self, y, iterations=10, mu0=None, Lambda0=None, nu0=None, inv_scale0=None, init_E_P=None):
        """Fused mean-field VMP of the multivariate IID model with Wishart precision (``rxg_mv_iid_wishart_vmp_f32``);
        y[N, d, batch]; defaults = the reference test's priors (mv_iid_precision_tests.jl:10-30)."""
        self._dev(y)
        N, d, batch = y.shape
        mu0 = np.zeros(d) if mu0 is None else mu0
        Lambda0 = 100.0 * np.eye(d) if Lambda0 is None else Lambda0
        nu0 = d + 1.0 if nu0 is None else nu0
        inv_scale0 = np.eye(d) if inv_scale0 is None else inv_scale0
        init_E_P = d * 1e12 * np.eye(d) if init_E_P is None else init_E_P       # mean of vague(Wishart, d)
        keep = [_model32(x) for x in (mu0, Lambda0, inv_scale0, init_E_P)]
        mm, mc = self.empty(d, batch), self.empty(d, d, batch)
        df, iS = self.empty(batch), self.empty(d, d, batch)
        st = self.empty(batch, dtype=torch.int32)
        self._check(self.lib.rxg_mv_iid_wishart_vmp_f32(self.h, d, N, batch, iterations, keep[0][1], keep[1][1], float(nu0),
                                                        keep[2][1], keep[3][1], _fp(y), _fp(mm), _fp(mc), _fp(df), _fp(iS),
                                                        ctypes.cast(c_void_p(st.data_ptr()), L.i32p), L.PTR_DEVICE))
        return dict(m_mean=mm, m_cov=mc, df=df, inv_scale=iS, status=st)

    def ar_vmp(self, series, order, iterations=15, gamma_prior=(1.0, 1.0), theta_prior_precision=1.0, init_gamma=(1.0, 1.0),
               want_free_energy=True):
        """Fused VMP of the reference's autoregressive regression model (``rxg_ar_vmp_f32``); series[N, batch]."""
        self._dev(series)
        N, batch = series.shape
        tm, tc = self.empty(order, batch), self.empty(order, order, batch)
        gs, gr = self.empty(batch), self.empty(batch)
        fe = self.empty(iterations, batch, dtype=torch.float64) if want_free_energy else None      # fp64 output (see the header)
        fe_p = ctypes.cast(c_void_p(fe.data_ptr() if fe is not None else None), ctypes.POINTER(ctypes.c_double))
        self._check(self.lib.rxg_ar_vmp_f32(self.h, order, N, batch, iterations, gamma_prior[0], gamma_prior[1], theta_prior_precision,
                                            init_gamma[0], init_gamma[1], _fp(series), _fp(tm), _fp(tc), _fp(gs), _fp(gr), fe_p,
                                            L.PTR_DEVICE))
        return dict(theta_mean=tm, theta_cov=tc, gamma_shape=gs, gamma_rate=gr, free_energy=fe)

    def lar_vmp(self, y, order, tau, iterations=15, gamma_prior=(1.0, 1.0), theta_prior_precision=1.0, x0_prior_precision=1.0,
                init_gamma=(1.0, 1.0), init_theta_precision=1.0, want_states=True, want_free_energy=True):
        """Fused structured VMP of the reference's latent autoregressive model (``rxg_lar_vmp_f32``,
        /root/reference/test/models/autoregressive/lar_tests.jl); y[T, batch] on the device.  Returns the KeepLast
        state posteriors and the KeepEach parameter posteriors / free energy, as the reference's ``returnvars``."""
        self._dev(y)
        if y.dim() != 2:
            raise ValueError("lar_vmp: y must be [T, batch]")
        T, batch = y.shape
        iters = int(iterations)
        prm = (ctypes.c_float * 8)(float(tau), float(gamma_prior[0]), float(gamma_prior[1]), float(theta_prior_precision),
                                   float(x0_prior_precision), float(init_gamma[0]), float(init_gamma[1]), float(init_theta_precision))
        xm = self.empty(T, order, batch) if want_states else None
        xc = self.empty(T, order, order, batch) if want_states else None
        tm, tc = self.empty(iters, order, batch), self.empty(iters, order, order, batch)
        gs, gr = self.empty(iters, batch), self.empty(iters, batch)
        fe = self.empty(iters, batch, dtype=torch.float64) if want_free_energy else None
        st = self.empty(batch, dtype=torch.int32)
        fe_p = ctypes.cast(c_void_p(fe.data_ptr() if fe is not None else None), ctypes.POINTER(ctypes.c_double))
        self._check(self.lib.rxg_lar_vmp_f32(self.h, int(order), T, batch, iters, ctypes.cast(prm, L.fp), _fp(y), _fp(xm), _fp(xc),
                                             _fp(tm), _fp(tc), _fp(gs), _fp(gr), fe_p,
                                             ctypes.cast(c_void_p(st.data_ptr()), L.i32p), L.PTR_DEVICE))
        return dict(x_mean=xm, x_cov=xc, theta_mean=tm, theta_cov=tc, gamma_shape=gs, gamma_rate=gr, free_energy=fe, status=st)

    def prod_gamma(self, a1, b1, a2, b2):
        return self._six(self.lib.rxg_prod_gamma_f32, a1, b1, a2, b2)

    def prod_normal(self, m1, v1, m2, v2):
        return self._six(self.lib.rxg_prod_normal_f32, m1, v1, m2, v2)

    def rule_gcv_out(self, m_x, v_x, m_z, v_z, kappa, omega):
        self._dev(m_x, v_x, m_z, v_z)
        n = m_x.numel()
        mo, vo = torch.empty_like(m_x), torch.empty_like(m_x)
        self._check(self.lib.rxg_rule_gcv_out_f32(self.h, n, _fp(m_x), _fp(v_x), _fp(m_z), _fp(v_z), kappa, omega, _fp(mo), _fp(vo), L.PTR_DEVICE))
        return mo, vo

    def marginalrule_gcv_yx(self, m_y, v_y, m_x, v_x, m_z, v_z, kappa, omega):
        self._dev(m_y, v_y, m_x, v_x, m_z, v_z)
        n = m_y.numel()
        m, V = self.empty(2, n), self.empty(2, 2, n)
        self._check(self.lib.rxg_marginalrule_gcv_yx_f32(self.h, n, _fp(m_y), _fp(v_y), _fp(m_x), _fp(v_x), _fp(m_z), _fp(v_z),
                                                         kappa, omega, _fp(m), _fp(V), L.PTR_DEVICE))
        return m, V

    def rule_gcv_z_prod(self, m_yx, V_yx, m_zp, v_zp, kappa, omega):
        self._dev(m_yx, V_yx, m_zp, v_zp)
        n = m_zp.numel()
        mz, vz = torch.empty_like(m_zp), torch.empty_like(m_zp)
        self._check(self.lib.rxg_rule_gcv_z_prod_f32(self.h, n, _fp(m_yx), _fp(V_yx), _fp(m_zp), _fp(v_zp), kappa, omega,
                                                     _fp(mz), _fp(vz), L.PTR_DEVICE))
        return mz, vz

    def selftest_umma(self, A, B):
        """D[128, n] = A[128, k] @ B[n, k].T on the tcgen05 tensor pipe (3xTF32), for the sweeps' operand shapes."""
        self._dev(A, B)
        n, k = B.shape
        D = self.empty(128, n)
        self._check(self.lib.rxg_selftest_umma_shape_f32(self.h, n, k, _fp(A), _fp(B), _fp(D), L.PTR_DEVICE))
        return D

    def selftest_stream(self, src, dst, asynchronous=True):
        """dst[w, :] = sum_r src[r, :] -- streaming kernel with a (n_read : n_write) HBM traffic mix."""
        nr, n = src.shape
        nw = dst.shape[0]
        self._dev(src if nr else None, dst if nw else None)
        self._check(self.lib.rxg_selftest_stream_f32(self.h, n, nr, nw, _fp(src), _fp(dst),
                                                     L.PTR_DEVICE | (L.ASYNC if asynchronous else 0)))
        return dst

    # ------------------------------------------------------------------ multi-GPU
    def comm_init(self, nranks, rank, uid: bytes):
        buf = ctypes.create_string_buffer(uid, 128)
        self._check(self.lib.rxg_comm_init(self.h, nranks, rank, ctypes.cast(buf, c_void_p)))
        self.comm_nranks, self.comm_rank = int(nranks), int(rank)

    # ---- peer-mapped gather (NVLink P2P stores from the sweep itself; rxg_peer.cu)
    def peer_open(self, handle: bytes) -> int:
        buf = ctypes.create_string_buffer(handle, 64)
        p = c_void_p()
        self._check(self.lib.rxg_peer_open(self.h, ctypes.cast(buf, c_void_p), ctypes.byref(p)))
        return int(p.value)

    def peer_close(self, ptr: int):
        self._check(self.lib.rxg_peer_close(self.h, c_void_p(ptr)))

    def peer_group(self, nranks: int, rank: int, flag_ptrs):
        arr = (c_void_p * max(nranks, 1))(*[c_void_p(int(p)) for p in flag_ptrs]) if nranks > 1 else None
        self._check(self.lib.rxg_peer_group(self.h, nranks, rank, arr))
        self.peer_nranks, self.peer_rank = int(nranks), int(rank)

    def peer_barrier(self, asynchronous=False):
        self._check(self.lib.rxg_peer_barrier(self.h, L.ASYNC if asynchronous else 0))

    def peer_allgather(self, local, gathered_ptrs, asynchronous=False):
        """``local`` (contiguous CUDA fp32) -> slab ``rank`` of every rank's gathered buffer, then the barrier."""
        self._dev(local)
        arr = (L.fp * len(gathered_ptrs))(*[L.as_fp(p) for p in gathered_ptrs])
        self._check(self.lib.rxg_peer_allgather_f32(self.h, local.numel(), _fp(local), arr,
                                                    L.PTR_DEVICE | (L.ASYNC if asynchronous else 0)))

    def lgssm_smooth_gather(self, y, A, B, P, Q, m0, S0, gathered_mean_ptrs, gathered_cov_ptrs=None, *, u=None, mask=None,
                            replicate_cov=False, want_evidence=False, want_status=False, force_per_chain_path=False,
                            transition_first=False, asynchronous=False):
        """Fused smoothing sweep + all-gather (``rxg_lgssm_smooth_gather_f32``).  ``gathered_*_ptrs[g]`` = base address
        of rank g's gathered buffer as mapped in this process (see ``sharding.PeerGroup``)."""
        self._io(y, "y", True)
        T, m, batch = y.shape
        self._io(mask, "mask", True, dtype=torch.uint8, shape=(T, batch))
        d = np.asarray(A).shape[-1]
        keep = [_model32(x) for x in (A, B, P, Q, m0, S0)]
        ptrs = [k[1] for k in keep]
        if u is not None:
            keep.append(_model32(u)); ptrs.append(keep[-1][1])
        else:
            ptrs.append(L.as_fp(0))
        flags = L.PTR_DEVICE
        if replicate_cov:
            flags |= L.COV_REPLICATE
        if force_per_chain_path:
            flags |= L.PATH_PER_CHAIN
        if transition_first:
            flags |= L.TRANSITION_FIRST
        if asynchronous:
            flags |= L.ASYNC
        G = len(gathered_mean_ptrs)
        gm = (L.fp * G)(*[L.as_fp(p) for p in gathered_mean_ptrs])
        gc = (L.fp * G)(*[L.as_fp(p) for p in gathered_cov_ptrs]) if gathered_cov_ptrs is not None else None
        nle = self.empty(batch) if want_evidence else None
        status = self.empty(batch, dtype=torch.int32) if want_status else None
        mask_p = ctypes.cast(c_void_p(mask.data_ptr() if mask is not None else None), L.u8p)
        st_p = ctypes.cast(c_void_p(status.data_ptr() if status is not None else None), L.i32p)
        self._check(self.lib.rxg_lgssm_smooth_gather_f32(self.h, d, m, T, batch, *ptrs, _fp(y), mask_p, gm, gc, _fp(nle), st_p, flags))
        return dict(neg_log_evidence=nle, status=status)

    def allgather_posteriors(self, mean, cov, nranks, out_mean=None, out_cov=None, replicate_cov=False):
        """Rank-major gathered slabs ([G, T, d, b], [G, T, d, d, b]); pass out_* to reuse buffers.
        ``replicate_cov=True`` (shared model on every rank, no missing data: chain-independent
        covariances): only the means cross NVLink, the covariance slabs are filled locally;
        ``cov`` may then also be the de-duplicated [T, d, d] table of ``cov_shared_out=True``."""
        self._dev(mean, cov)
        T, d, bl = mean.shape
        gm = out_mean if out_mean is not None else self.empty(nranks, T, d, bl)
        gc = None
        flags = L.PTR_DEVICE
        if cov is not None:
            gc = out_cov if out_cov is not None else self.empty(nranks, T, d, d, bl)
            if replicate_cov:
                flags |= L.COV_REPLICATE
                if cov.dim() == 3:
                    flags |= L.COV_SHARED_OUT
            elif cov.dim() == 3:
                raise ValueError("a [T, d, d] covariance table can only be replicated (replicate_cov=True)")
        self._check(self.lib.rxg_allgather_posteriors(self.h, d, T, bl, _fp(mean), _fp(cov), _fp(gm), _fp(gc), flags))
        return gm, gc


def host_empty(*shape, dtype=torch.float32):
    """Pinned host tensor from ``rxg_host_alloc`` -- what a C / Julia host of the ABI would use: page-locked and, on a
    multi-socket machine, interleaved over the NUMA nodes (see rxg_api.cu).  Freed when the tensor is collected."""
    lib = L.load()
    n = int(np.prod(shape))
    item = torch.empty((), dtype=dtype).element_size()
    p = c_void_p()
    rc = lib.rxg_host_alloc(ctypes.byref(p), max(n * item, 1))
    if rc != 0:
        raise L.RxGaussError(rc, "rxg_host_alloc failed")
    buf = (ctypes.c_char * (n * item)).from_address(p.value)
    t = torch.frombuffer(buf, dtype=dtype, count=n).reshape(*shape)
    import weakref
    weakref.finalize(buf, lib.rxg_host_free, c_void_p(p.value))
    t._rxg_keep = buf
    return t


class DeviceBuffer:
    """Device memory from ``rxg_device_alloc`` (plain cudaMalloc: exportable as a CUDA IPC handle at offset 0),
    viewable as a torch tensor through ``__cuda_array_interface__``.  Freed with the object."""

    def __init__(self, ctx: "Context", nbytes: int, zero: bool = False):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = c_void_p()
        ctx._check(ctx.lib.rxg_device_alloc(ctx.h, self.nbytes, ctypes.byref(p)))
        self.ptr = int(p.value)
        if zero:
            ctx._check(ctx.lib.rxg_device_memset(ctx.h, c_void_p(self.ptr), 0, self.nbytes))

    def export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self.ctx._check(self.ctx.lib.rxg_peer_export(self.ctx.h, c_void_p(self.ptr), ctypes.cast(buf, c_void_p)))
        return buf.raw

    def tensor(self, *shape, dtype=torch.float32):
        n = int(np.prod(shape))
        itemsize = torch.empty((), dtype=dtype).element_size()
        assert n * itemsize <= self.nbytes
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        holder = type("_CAI", (), {})()
        holder.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                           "data": (self.ptr, False), "version": 2, "strides": None}
        holder._keep = self
        return torch.as_tensor(holder, device=f"cuda:{self.ctx.device}")

    def free(self):
        if getattr(self, "ptr", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.rxg_device_free(self.ctx.h, c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    lib = L.load()
    buf = ctypes.create_string_buffer(128)
    rc = lib.rxg_comm_unique_id(ctypes.cast(buf, c_void_p))
    if rc != 0:
        raise L.RxGaussError(rc, "rxg_comm_unique_id failed")
    return buf.raw
