/* rxgauss.h -- C ABI of librxgauss: B200 (sm_100a) kernels for the Gaussian message-passing hot
 * path of RxInfer.jl's infer().
 *
 * This header is the drop-in boundary.  Every entry point replaces one piece of the reference's
 * per-message / per-chain machinery; the reference-side interface each one stands in for is cited
 * as  [ref: file:line]  relative to /root/reference.  Rule bodies marked (upstream) live in the
 * un-vendored ReactiveMP ~6.0.0 / ExponentialFamily 2.1.0 / BayesBase 1.5.0 / FastCholesky 1.3.0
 * (Project.toml:43-73); the in-repo citation is the call site that binds or exercises them.
 *
 * Conventions
 *  - plain C, no exceptions, no callbacks.  Every function returns an rxg_status (0 = OK);
 *    rxg_last_error(ctx) gives a human-readable message for the last failure on that ctx.
 *  - all arrays are fp32, structure-of-arrays with the batch (message / chain) index INNERMOST:
 *        vectors  v[k][n]          -> v[k*n_total + i]
 *        matrices M[r][c][n]       -> M[(r*C + c)*n_total + i]
 *        series   y[t][k][batch]   -> y[(t*m + k)*batch + b]
 *    so that a warp reading one component of 32 consecutive chains issues one 128-byte request.
 *  - (mu, Sigma) = mean / covariance  (MvNormalMeanCovariance),
 *    (xi, W)     = weighted mean / precision (MvNormalWeightedMeanPrecision), W = inv(Sigma), xi = W mu.
 *  - pointers are device pointers when RXG_PTR_DEVICE is set in `flags`, host pointers otherwise
 *    (host buffers are staged through the context's stream; pinned memory from rxg_host_alloc
 *    makes the copies asynchronous).
 *  - the caller owns every buffer it passes; the library owns only what hangs off rxg_ctx
 *    (stream handle, workspace, gain tables, NCCL communicator).
 *  - a ctx is bound to one device and is not thread-safe: one ctx per host thread and per GPU.
 *  - calls return after the stream has been synchronised unless RXG_ASYNC is set.
 *  - there is NO CPU fallback: without a usable CUDA device rxg_create fails with
 *    RXG_ERR_NO_DEVICE and every compute entry point fails with RXG_ERR_BAD_ARG on a null ctx.
 */
#ifndef RXGAUSS_H
#define RXGAUSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RXG_VERSION 100 /* 0.1.0 */

typedef struct rxg_ctx rxg_ctx;

typedef enum rxg_status {
    RXG_OK = 0,
    RXG_ERR_BAD_ARG = 1,
    RXG_ERR_CUDA = 2,
    RXG_ERR_NCCL = 3,
    RXG_ERR_NOT_SPD = 4,     /* at least one chain/message hit a non-positive Cholesky pivot      */
    RXG_ERR_NAN = 5,
    RXG_ERR_UNSUPPORTED = 6, /* shape outside the compiled kernel families (see rxg_supports)   */
    RXG_ERR_NO_DEVICE = 7
} rxg_status;

enum rxg_flags {
    RXG_PTR_DEVICE      = 1u << 0, /* data pointers are device pointers                          */
    RXG_MODEL_PER_CHAIN = 1u << 1, /* A,B,P,Q,m0,S0 carry a trailing [batch] axis               */
    RXG_ASYNC           = 1u << 2, /* do not synchronise the stream before returning             */
    RXG_COV_SHARED_OUT  = 1u << 3, /* shared model only: write post_cov as [T][d][d] (one copy)  */
    RXG_PATH_PER_CHAIN  = 1u << 4, /* force the per-chain covariance recursion (no gain tables)  */
    RXG_TRANSITION_FIRST = 1u << 5, /* the prior sits one transition before the first datum      */
    RXG_COV_REPLICATE   = 1u << 6, /* all-gather: covariances are chain independent (shared model,
                                      no missing data) -- replicate them locally, gather only means */
    RXG_MASK_SHARED     = 1u << 7  /* ymask is ONE pattern for all chains: a HOST array ymask[T] (like the model
                                      matrices).  The covariances stay chain independent, so the call stays on the
                                      gain-table path (a per-chain mask forces the per-chain covariance recursion,
                                      ~2.7x slower at d = 4).  y at masked steps is ignored but must be finite.   */
};

/* Per-context options (rxg_set_option).  The RXG_* environment variables of the same name are read
 * ONCE, inside rxg_create, as the initial values; nothing on a compute path calls getenv.        */
typedef enum rxg_option {
    RXG_OPT_GAIN_SEQ = 0,          /* 1: sequential Riccati gain kernels (cross-check of the time-parallel scan)   */
    RXG_OPT_LARGE_SEQ = 1,         /* 1: sequential gain kernels of the large-state family (cross-check)            */
    RXG_OPT_NO_UMMA = 2,           /* 1: d >= 16 mean recursions on the FP32 pipe instead of tcgen05 (cross-check)  */
    RXG_OPT_SWEEP_VARIANT = 3,     /* shared-model sweep: 0 auto, 1 stash, 3 time-segmented (experimental, slower)  */
    RXG_OPT_FORCE_CPT = 4,         /* chains per thread of the shared-model sweep (0 = auto)                        */
    RXG_OPT_HOST_THREADS = 5,      /* host threads of the host-side covariance broadcast (0 = auto)                 */
    RXG_OPT_HOST_COV_D2H = 6,      /* 1: host-pointer calls copy the per-chain covariances over PCIe (no broadcast) */
    RXG_OPT_HOST_BCAST_MIN_MB = 7, /* below this covariance size the host broadcast is not used (default 64)        */
    RXG_OPT_HOST_SLICES = 8,       /* batch slices of the host-pointer pipeline (0 = auto)                          */
    RXG_OPT_GATHER_MODE = 9,       /* rxg_lgssm_smooth_gather_f32: 0/1 peer stores fused into the sweep, 2 push after it  */
    RXG_OPT_COUNT_ = 10
} rxg_option;

#define RXG_MAX_PEERS 8            /* ranks of one peer group (one NVSwitch domain)                                 */

/* ------------------------------------------------------------------ context / plumbing ------ */
int rxg_version(void);
/* Create a context on CUDA device `device`.  [ref: the reference has no device/ctx notion; this
 * replaces the per-infer() engine state built at src/inference/batch.jl:177-257]               */
int rxg_create(rxg_ctx** out, int device, unsigned flags);
int rxg_destroy(rxg_ctx* ctx);
const char* rxg_last_error(const rxg_ctx* ctx);
int rxg_set_option(rxg_ctx* ctx, int option, long long value);
int rxg_get_option(const rxg_ctx* ctx, int option, long long* value);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all launches/copies.  */
int rxg_set_stream(rxg_ctx* ctx, void* cuda_stream);
int rxg_sync(rxg_ctx* ctx);
/* Pinned host memory for asynchronous staging of host-pointer calls.                           */
int rxg_host_alloc(void** out, size_t bytes);
int rxg_host_free(void* p);
/* 1 if the fused sweeps accept (d, m): any 1 <= d, m <= 64.  Shapes without a dedicated kernel family are
 * embedded in the next larger one (shared model) or run on the generic one-CTA-per-chain kernel (per-chain
 * models, missing data); see csrc/rxg_lgssm_general.cu.                                          */
int rxg_supports(int d, int m);
/* Host threads the library will use to broadcast chain-independent covariances into a HOST output
 * buffer (host-pointer calls of a shared model fetch the [T][d][d] table once instead of copying the
 * per-chain duplicates over PCIe): RXG_HOST_THREADS, else min(affinity, cgroup quota, 16) divided by
 * LOCAL_WORLD_SIZE.  Below 6 the full device->host copy is used (also forced by RXG_HOST_COV_D2H=1). */
int rxg_host_fill_threads(void);
/* Number of kernels this ctx has launched so far (for bench.py's gpu_launches).               */
long long rxg_launch_count(const rxg_ctx* ctx);
/* Per-kernel timing of the most recent fused LGSSM sweep: when enabled, CUDA events are recorded
 * on the ctx stream around the gain-table kernels and around the dominant sweep kernel.
 * rxg_profile_last_ms synchronises on those events; *gain_ms is 0 on the per-chain path.        */
int rxg_set_profiling(rxg_ctx* ctx, int enabled);
int rxg_profile_last_ms(rxg_ctx* ctx, float* main_kernel_ms, float* gain_kernels_ms);

/* ------------------------------------------------------------------ per-rule kernels --------
 * Batched twins of the reference's @rule bodies: n independent messages per call, pure
 * functions.  `M_shared != 0` means the PointMass matrix operand is one d x d (row-major) matrix
 * shared by all n messages, else it is [r][c][n].
 * The reference reaches these through ReactiveMP.rule(...) dispatched from the edge pipelines
 * wired in activate_rmp_factornode! [ref: src/model/plugins/reactivemp_inference.jl:509-540];
 * callable directly as @call_rule [ref: test/inference/inference_tests.jl:547-585].
 * State sizes: any 1 <= d <= 64 (d_out, d_in <= 64 for the multiplication rules).  d in {1..6, 8} (and the rectangular
 * shapes 1x2, 1x4, 2x4) run register resident, one thread per message; every other size runs on the shared-memory /
 * left-GEMM kernels of csrc/rxg_rules_large.cu, where the multiplication rules need M_shared != 0.                  */

/* @rule MvNormalMeanCovariance(:out)(m_mu, q_Sigma) -> (mu, S + Sigma)  (upstream
 * rules/mv_normal_mean_covariance/out.jl)  [ref: alias src/model/graphppl.jl:372-376;
 * benchmarks/Linear...Benchmark.ipynb:102]                                                      */
int rxg_rule_mvnormal_meancov_out_f32(rxg_ctx*, int64_t n, int d, const float* mu_in,
                                      const float* S_in, const float* Sigma, int M_shared,
                                      float* mu_out, float* S_out, unsigned flags);
/* @rule MvNormalMeanCovariance(:mu)(m_out, q_Sigma) -> (mu_out, S_out + Sigma)  (upstream
 * .../mean.jl)  [ref: ipynb:102-103]                                                             */
int rxg_rule_mvnormal_meancov_mean_f32(rxg_ctx*, int64_t n, int d, const float* mu_in,
                                       const float* S_in, const float* Sigma, int M_shared,
                                       float* mu_out, float* S_out, unsigned flags);
/* same rule with q_out::PointMass (a datum pushed by new_observation!
 * [ref: src/inference/batch.jl:405-407]) -> (y, Sigma)                                          */
int rxg_rule_mvnormal_meancov_mean_data_f32(rxg_ctx*, int64_t n, int d, const float* y,
                                            const float* Sigma, int M_shared, float* mu_out,
                                            float* S_out, unsigned flags);
/* @rule typeof(*)(:out)(m_A::PointMass, m_in) -> (A mu, A S A')  (upstream
 * rules/multiplication/out.jl)  [ref: ipynb:102; src/model/graphppl.jl:58-83]; A is d_out x d_in */
int rxg_rule_mul_out_f32(rxg_ctx*, int64_t n, int d_out, int d_in, const float* A, int M_shared,
                         const float* mu_in, const float* S_in, float* mu_out, float* S_out,
                         unsigned flags);
/* @rule typeof(*)(:in)(m_out, m_A::PointMass) -> (xi, W) = (A' W_out mu_out, A' W_out A) with
 * W_out = cholinv(S_out)  (upstream rules/multiplication/in.jl).  status[n] (optional) receives
 * RXG_ERR_NOT_SPD per message.                                                                  */
int rxg_rule_mul_in_f32(rxg_ctx*, int64_t n, int d_out, int d_in, const float* A, int M_shared,
                        const float* mu_out, const float* S_out, float* xi_in, float* W_in,
                        int32_t* status, unsigned flags);
/* @rule typeof(+)(:out) -> (mu1 + mu2, S1 + S2); (:in1)/(:in2) -> (mu_out - mu_other, S_out +
 * S_other)  (upstream rules/addition)  [ref: test/models/statespace/ulgssm_tests.jl:12]         */
int rxg_rule_add_out_f32(rxg_ctx*, int64_t n, int d, const float* mu1, const float* S1,
                         const float* mu2, const float* S2, float* mu_out, float* S_out,
                         unsigned flags);
int rxg_rule_add_in_f32(rxg_ctx*, int64_t n, int d, const float* mu_out, const float* S_out,
                        const float* mu_other, const float* S_other, float* mu_in, float* S_in,
                        unsigned flags);
/* BayesBase.prod(::MvNormal, ::MvNormal) in (xi, W): (xi1 + xi2, W1 + W2)
 * [ref: fold at src/model/plugins/reactivemp_inference.jl:365-374]                              */
int rxg_prod_gaussian_f32(rxg_ctx*, int64_t n, int d, const float* xi1, const float* W1,
                          const float* xi2, const float* W2, float* xi, float* W, unsigned flags);
/* weightedmean_precision(::MvNormalMeanCovariance) / mean_cov(::MvNormalWeightedMeanPrecision):
 * one Cholesky SPD inverse each (FastCholesky.cholinv) [ref: re-export src/RxInfer.jl:6]        */
int rxg_meancov_to_wmp_f32(rxg_ctx*, int64_t n, int d, const float* mu, const float* S, float* xi,
                           float* W, int32_t* status, unsigned flags);
int rxg_wmp_to_meancov_f32(rxg_ctx*, int64_t n, int d, const float* xi, const float* W, float* mu,
                           float* S, int32_t* status, unsigned flags);
/* Marginal at a random variable: product of k inbound (xi, W) messages then mean_cov
 * [ref: src/model/plugins/reactivemp_inference.jl:370-455]; xi_list/W_list are arrays of k
 * HOST-resident pointers to the message buffers.                                                */
int rxg_marginal_gaussian_f32(rxg_ctx*, int64_t n, int d, int k, const float* const* xi_list,
                              const float* const* W_list, float* mu, float* S, int32_t* status,
                              unsigned flags);

/* Univariate / Gamma-precision VMP rules
 * @rule NormalMeanPrecision(:tau)(q_out, q_mu) -> GammaShapeRate(3/2, ((m_o-m_m)^2+v_o+v_m)/2)
 * [ref: test/models/aliases/aliases_gamma_tests.jl:13-18]                                        */
int rxg_rule_normal_precision_tau_f32(rxg_ctx*, int64_t n, const float* m_out, const float* v_out,
                                      const float* m_mu, const float* v_mu, float* shape,
                                      float* rate, unsigned flags);
/* @rule NormalMeanPrecision(:out)(m_mu, q_tau) -> N(m_mu, v_mu + rate/shape)                    */
int rxg_rule_normal_precision_out_f32(rxg_ctx*, int64_t n, const float* m_mu, const float* v_mu,
                                      const float* shape, const float* rate, float* m_out,
                                      float* v_out, unsigned flags);
/* structured variant (q_out_mu jointly Gaussian, m_joint[2][n], V_joint[2][2][n]): GammaShapeRate(3/2,
 * 1/2 [V11 + V22 - V12 - V21 + (m1 - m2)^2])  (upstream rules/normal_mean_precision/precision.jl)              */
int rxg_rule_normal_precision_tau_joint_f32(rxg_ctx*, int64_t n, const float* m_joint, const float* V_joint,
                                            float* shape, float* rate, unsigned flags);
/* Wishart precision -- the multivariate twin of the Gamma rules, in the WishartFast parametrisation (df, INVERSE
 * scale) so that products are additions [ref: test/models/iid/mv_iid_precision_tests.jl:10-41]:
 * @rule MvNormalMeanPrecision(:Lambda)(q_out, q_mu) -> Wishart(d + 2, inv(V_out + V_mu + (m_out - m_mu)(m_out - m_mu)'))  */
int rxg_rule_mvnormal_precision_lambda_f32(rxg_ctx*, int64_t n, int d, const float* m_out, const float* V_out,
                                           const float* m_mu, const float* V_mu, float* df, float* inv_scale,
                                           unsigned flags);
/* prod(Wishart, Wishart) = Wishart(df1 + df2 - d - 1, inv(inv(S1) + inv(S2)))                                   */
int rxg_prod_wishart_f32(rxg_ctx*, int64_t n, int d, const float* df1, const float* inv_scale1, const float* df2,
                         const float* inv_scale2, float* df, float* inv_scale, unsigned flags);
/* mean(Wishart(df, S)) = df * S = df * inv(inv_scale); status[n] optional                                      */
int rxg_wishart_mean_f32(rxg_ctx*, int64_t n, int d, const float* df, const float* inv_scale, float* mean,
                         int32_t* status, unsigned flags);
/* Fused mean-field VMP of the multivariate IID model with unknown mean and precision, `batch` independent data sets:
 *   m ~ MvNormal(mu0, inv(Lambda0)),  P ~ Wishart(nu0, inv(inv_scale0)),  y[i] ~ MvNormal(m, inv(P)),  q(m) q(P)
 * [ref: model, constraints and initialisation test/models/iid/mv_iid_precision_tests.jl:10-41].  y[N][d][batch];
 * init_E_P[d][d] = mean of the initial q(P) (host); outputs q(m) = (m_mean[d][batch], m_cov[d][d][batch]),
 * q(P) = Wishart(df[batch], inv(inv_scale[d][d][batch])).  d <= 6.                                                */
int rxg_mv_iid_wishart_vmp_f32(rxg_ctx*, int d, int N, int64_t batch, int iterations, const float* mu0,
                               const float* Lambda0, float nu0, const float* inv_scale0, const float* init_E_P,
                               const float* y, float* m_mean, float* m_cov, float* df, float* inv_scale,
                               int32_t* status, unsigned flags);
/* Fused mean-field VMP of the reference's autoregressive regression model, `batch` independent series:
 *   gamma ~ Gamma(a0, b0), theta ~ MvNormal(0, I / theta_prior_precision), y[i] ~ Normal(dot(x[i], theta), 1 / gamma),
 *   x[i] = (s[i-1], ..., s[i-order]) lags of the series itself, q(gamma) q(theta), q(gamma) initialised to
 *   Gamma(init_shape, init_rate) [ref: test/models/autoregressive/ar_tests.jl:7-36].  series[N][batch]; outputs
 *   theta_mean[order][batch], theta_cov[order][order][batch], gamma_shape / gamma_rate[batch], free_energy[iterations][batch]
 *   or NULL (Bethe free energy after every iteration, in fp64: the reference asserts it to decrease, :69-70, and the
 *   decreases are ~1e-5 on values of ~1.4e3).  order <= 8.                                                         */
int rxg_ar_vmp_f32(rxg_ctx*, int order, int N, int64_t batch, int iterations, float a0, float b0,
                   float theta_prior_precision, float init_shape, float init_rate, const float* series,
                   float* theta_mean, float* theta_cov, float* gamma_shape, float* gamma_rate,
                   double* free_energy, unsigned flags);
/* Fused structured VMP of the reference's LATENT autoregressive model, `batch` independent series, one launch:
 *   gamma ~ Gamma(a0, b0); theta ~ N(0, I / w0); x0 ~ N(0, I / p0); x[t] ~ AR(x[t-1], theta, gamma) with
 *   ARMeta(Multivariate | Univariate, order, ARsafe()); y[t] ~ Normal(dot(c, x[t]), 1 / tau), c = e1 (ReactiveMP.ar_unit);
 *   q(x, x0) q(gamma) q(theta); q(gamma), q(theta) initialised to Gamma(init_shape, init_rate), N(0, I / init_theta_precision)
 *   [ref: test/models/autoregressive/lar_tests.jl:52-122; free-energy pins :170 (AR(1): 518.9182342) and :201 (AR(5): 514.66086)].
 *   The Univariate AR(1) node is the order = 1 case.  params (HOST, 8 floats, all > 0) = {tau, a0, b0, w0, p0, init_shape,
 *   init_rate, init_theta_precision}.  y[T][batch]; outputs: x_mean[T][order][batch], x_cov[T][order][order][batch] of the
 *   LAST iteration (KeepLast; either may be NULL), theta_mean[iterations][order][batch],
 *   theta_cov[iterations][order][order][batch], gamma_shape / gamma_rate[iterations][batch] (KeepEach),
 *   free_energy[iterations][batch] in fp64 or NULL, status[batch] or NULL.  Every iteration is a covariance-form filter + RTS
 *   smoother over the companion-matrix state space in the exact limit of the AR node's deterministic coordinates (the
 *   reference regularises them with a precision of 1e12), fp32 recursions, fp64 statistics / parameter updates / free energy.
 *   order <= 6.  Device pointers.                                                                                   */
int rxg_lar_vmp_f32(rxg_ctx*, int order, int T, int64_t batch, int iterations, const float* params, const float* y,
                    float* x_mean, float* x_cov, float* theta_mean, float* theta_cov, float* gamma_shape,
                    float* gamma_rate, double* free_energy, int32_t* status, unsigned flags);
/* prod(GammaShapeRate, GammaShapeRate) = (a1 + a2 - 1, b1 + b2)                                 */
int rxg_prod_gamma_f32(rxg_ctx*, int64_t n, const float* a1, const float* b1, const float* a2,
                       const float* b2, float* a, float* b, unsigned flags);
/* prod of two univariate Normals in (mean, variance) I/O                                        */
int rxg_prod_normal_f32(rxg_ctx*, int64_t n, const float* m1, const float* v1, const float* m2,
                        const float* v2, float* m, float* v, unsigned flags);

/* GCV node rules [ref: test/models/statespace/hgf_tests.jl:10-40;
 * formulas restated at test/inference/inference_tests.jl:587-607]; kappa, omega PointMass.
 * @rule GCV(:y)(m_x, q_z, ...) -> N(m_x, v_x + 1/(A B))  (and symmetric :x)                     */
int rxg_rule_gcv_out_f32(rxg_ctx*, int64_t n, const float* m_x, const float* v_x,
                         const float* m_z, const float* v_z, float kappa, float omega,
                         float* m_out, float* v_out, unsigned flags);
/* @marginalrule GCV(:y_x) -> joint (m[2][n], V[2][2][n])                                        */
int rxg_marginalrule_gcv_yx_f32(rxg_ctx*, int64_t n, const float* m_y, const float* v_y,
                                const float* m_x, const float* v_x, const float* m_z,
                                const float* v_z, float kappa, float omega, float* m, float* V,
                                unsigned flags);
/* @rule GCV(:z)(q_y_x, ...) -> ExponentialLinearQuadratic(a,b,c,d), then prod(Normal prior, ELQ)
 * by GaussHermiteCubature(31) moment matching -> q(z) = N(m_z, v_z)                             */
int rxg_rule_gcv_z_prod_f32(rxg_ctx*, int64_t n, const float* m_yx, const float* V_yx,
                            const float* m_zprior, const float* v_zprior, float kappa, float omega,
                            float* m_z, float* v_z, unsigned flags);

/* ------------------------------------------------------------------ fused whole-chain sweeps --
 * Replace, for the batched case, the Rocket-driven schedule + per-message dispatch that one
 * infer(model = linear_gaussian_ssm_smoothing(...), data = (y = ...,)) executes
 * [ref: iteration loop src/inference/batch.jl:391-430; model benchmarks/...ipynb:95-105;
 *  test/models/statespace/mlgssm_test.jl:8-17].  One launch runs the forward sweep t = 1..T and
 * the backward sweep t = T..1 for `batch` independent chains: 6 rule messages + 2 products +
 * 1 marginal per (chain, step).
 *
 *   x[1] ~ N(m0, S0);  x[t] ~ N(A x[t-1] + u, P);  y[t] ~ N(B x[t], Q)     (A d x d, B m x d)
 *
 * `u` (d floats, or NULL for none) is a constant transition offset: the `+` rule with a PointMass
 * operand fused into the sweep [ref: `x[i] ~ x_prev + c`, test/models/statespace/
 * ulgssm_tests.jl:12].  With RXG_TRANSITION_FIRST the prior sits on the state BEFORE x[1]
 * (x_prior ~ N(m0, S0); x[1] ~ N(A x_prior + u, P)), as in test/models/statespace/
 * mlgssm_test.jl:8-17 and the notebook's one-step filtering model (ipynb:110-113).
 *
 * Inputs : y[T][m][batch]; ymask[T][batch] (uint8, 1 = observed) or NULL
 *          [ref: missing data semantics docs/src/manuals/inference/static.md:98-125];
 *          A,B,P,Q,m0,S0,u row-major, shared HOST arrays (or device [..][batch] arrays with
 *          RXG_MODEL_PER_CHAIN).
 * Outputs: post_mean[T][d][batch], post_cov[T][d][d][batch]  (== posteriors[:x], as
 *          MvNormalMeanCovariance) [ref: src/inference/batch.jl:475-481];
 *          neg_log_evidence[batch] or NULL (== Bethe free energy on this tree
 *          [ref: src/model/plugins/reactivemp_free_energy.jl:84-126]);
 *          status[batch] or NULL (per-chain RXG_OK / RXG_ERR_NOT_SPD / RXG_ERR_NAN).
 * post_mean / post_cov double as the forward->backward stash (no extra workspace).             */
int rxg_lgssm_smooth_f32(rxg_ctx*, int d, int m, int T, int64_t batch, const float* A,
                         const float* B, const float* P, const float* Q, const float* m0,
                         const float* S0, const float* u, const float* y, const uint8_t* ymask,
                         float* post_mean, float* post_cov, float* neg_log_evidence,
                         int32_t* status, unsigned flags);
/* Forward half only (filtering) -- what the streaming engine computes per datum with
 * @autoupdates x_min_t_mean, x_min_t_cov = mean_cov(q(x_t))
 * [ref: src/inference/streaming.jl:344-388; src/inference/autoupdates.jl:614-659; ipynb:199-216]. */
int rxg_lgssm_filter_f32(rxg_ctx*, int d, int m, int T, int64_t batch, const float* A,
                         const float* B, const float* P, const float* Q, const float* m0,
                         const float* S0, const float* u, const float* y, const uint8_t* ymask,
                         float* filt_mean, float* filt_cov, float* neg_log_evidence,
                         int32_t* status, unsigned flags);
/* VMP around the smoother with an unknown observation precision shared over time, one per chain
 * (d = m = 1): y[t] ~ N(x[t], 1/tau), tau ~ Gamma(a0, b0), q(x) q(tau)
 * [ref: rules of test/models/aliases/aliases_gamma_tests.jl; use case
 *  test/callbacks/benchmark_tests.jl:8-37].  y[T][batch]; outputs post_mean/var[T][batch],
 * shape/rate[batch].                                                                            */
int rxg_lgssm_vmp_gamma_f32(rxg_ctx*, int T, int64_t batch, int iterations, float a, float v_proc,
                            float m0, float v0, float a0, float b0, float init_E_tau,
                            const float* y, float* post_mean, float* post_var, float* shape,
                            float* rate, unsigned flags);
/* Same with the Bethe free energy after every iteration, free_energy[iterations][batch] (or NULL)
 * [ref: free_energy = true, src/inference/batch.jl:184-189; definition src/model/plugins/reactivemp_free_energy.jl:84-126]. */
int rxg_lgssm_vmp_gamma_fe_f32(rxg_ctx*, int T, int64_t batch, int iterations, float a, float v_proc,
                               float m0, float v0, float a0, float b0, float init_E_tau, const float* y,
                               float* post_mean, float* post_var, float* shape, float* rate,
                               float* free_energy, unsigned flags);
/* Hierarchical Gaussian Filter, streaming, `iters` VMP iterations per datum
 * [ref: test/models/statespace/hgf_tests.jl:10-69; loop src/inference/streaming.jl:349-407].
 * y[T][batch]; init = (m_z, v_z, m_x, v_x); out[T][4][batch] = (m_x, v_x, m_z, v_z).            */
int rxg_hgf_filter_f32(rxg_ctx*, int T, int64_t batch, int iters, float kappa, float omega,
                       float z_variance, float y_variance, const float init[4], const float* y,
                       float* out, unsigned flags);

/* Same filter with the streaming carry and the Bethe free energy as optional arguments: `init` (first chunk) or
 * `prev[4][batch]` (out[Tc-1] of the previous chunk), exactly one of them non-NULL; free_energy[T][iters][batch] or
 * NULL = the Bethe free energy of each datum's graph after every VMP iteration [ref: definition
 * src/model/plugins/reactivemp_free_energy.jl:84-126; the reference's regression pin for this model is the average
 * over the data after the last iteration, test/models/statespace/hgf_tests.jl:112-119 (1.009879989585)].        */
int rxg_hgf_filter_fe_f32(rxg_ctx*, int T, int64_t batch, int iters, float kappa, float omega,
                          float z_variance, float y_variance, const float init[4], const float* prev,
                          const float* y, float* out, float* free_energy, unsigned flags);

/* ------------------------------------------------------------------ streaming engine ----------
 * The reference's second entry point: infer(..., autoupdates = ..., keephistory = ...) builds an
 * RxInferenceEngine that re-triggers a ONE-step graph per datum and feeds q(x_t) back as the next
 * prior [ref: executor src/inference/streaming.jl:344-430; @autoupdates x_min_t_mean, x_min_t_cov =
 * mean_cov(q(x_t)) src/inference/autoupdates.jl:614-659; model ipynb:107-113, run ipynb:199-216].
 * Here the datastream is consumed in time-chunks: one call = one fused filtering sweep over Tc data
 * for all chains, with the autoupdate carry made explicit (no hidden state in the ctx):
 *   prev_mean[d][batch]  (device)  in : means of q(x_{t0-1}) -- for the first chunk the broadcast
 *                                       initialisation; afterwards filt_mean[Tc-1] of the last chunk
 *   carry_cov[d][d]      (HOST)    in : covariance of q(x_{t0-1}) (chain independent for a shared
 *                                       model);  out: covariance of q(x_{t0+Tc-1})
 * Model per datum: x_t ~ N(A x_{t-1} + u, P), y_t ~ N(B x_t, Q) (transition first).  Chunking is
 * exact: any split of the stream gives the same posteriors as one call (up to the fp32 rounding of
 * the carried covariance).  Shared model, no mask; device pointers; always synchronous.          */
int rxg_lgssm_filter_chunk_f32(rxg_ctx*, int d, int m, int Tc, int64_t batch, const float* A,
                               const float* B, const float* P, const float* Q, const float* u,
                               const float* prev_mean, float* carry_cov, const float* y,
                               float* filt_mean, float* filt_cov, float* neg_log_evidence,
                               unsigned flags);
/* HGF datastream in time-chunks: prev[4][batch] = out[Tc-1] of the previous chunk (rows m_x, v_x,
 * m_z, v_z: the @autoupdates of hgf_tests.jl:46-49).  First chunk: rxg_hgf_filter_f32 with init.  */
int rxg_hgf_filter_chunk_f32(rxg_ctx*, int Tc, int64_t batch, int iters, float kappa, float omega,
                             float z_variance, float y_variance, const float* prev, const float* y,
                             float* out, unsigned flags);

/* Streaming mean-field VMP with an unknown observation precision -- the reference's `test_model1`
 * [ref: test/inference/inference_tests.jl:752-775; @autoupdates :772-776; rules of
 * test/models/aliases/aliases_gamma_tests.jl]:  x_t_min ~ N(prior), tau ~ Gamma(prior),
 * x_t ~ N(x_t_min, 1/w), y ~ N(x_t, 1/tau), MeanField(), `iters` iterations per datum, priors autoupdated
 * from q(x_t), q(tau).  init = (m_x, v_x, shape, rate) of the @initialization, or prev[4][batch] = out[Tc-1]
 * of the previous time-chunk (then init may be NULL).  y[T][batch]; out[T][4][batch] = (m_x, v_x, shape, rate)
 * after the last iteration of each datum; free_energy[T][iters][batch] or NULL (Bethe free energy per
 * datum and iteration; the reference asserts its average over data to be non-increasing, :846).       */
int rxg_stream_vmp_gamma_f32(rxg_ctx*, int T, int64_t batch, int iters, float w, const float init[4],
                             const float* prev, const float* y, float* out, float* free_energy,
                             unsigned flags);

/* Diagnostic: D[128][64] = A[128][128] * B[64][128]' on the tcgen05 tensor pipe (kind::tf32, 3xTF32
 * split, TMEM accumulator), row-major device arrays.  Validates the hand-written UMMA descriptors
 * used by the large-state family; no reference counterpart.                                     */
int rxg_selftest_umma_f32(rxg_ctx*, const float* A, const float* B, float* D, unsigned flags);
/* Same for every operand shape the sweeps issue: D[128][n] = A[128][k] * B[n][k]' with
 * (n, k) in {(64,128), (128,64), (64,64), (64,32), (32,32), (32,16), (16,16)}.                    */
int rxg_selftest_umma_shape_f32(rxg_ctx*, int n, int k, const float* A, const float* B, float* D,
                                unsigned flags);

/* Diagnostic: stream n floats from each of n_read input rows (src[n_read][n]) and store their sum into each of
 * n_write output rows (dst[n_write][n]) -- a dependency-free kernel with a chosen HBM read : write mix, the
 * yardstick for the sweep kernels' achieved bandwidth (bench_extra.py --which stream).  No reference counterpart. */
int rxg_selftest_stream_f32(rxg_ctx*, int64_t n, int n_read, int n_write, const float* src, float* dst,
                            unsigned flags);

/* Diagnostic: write bandwidth (GB/s) of the host-side covariance broadcast into dst[rows][batch] (host memory).      */
double rxg_selftest_host_fill_gbs(float* dst, int64_t rows, int64_t batch, int nthreads, int reps);

/* ------------------------------------------------------------------ multi-GPU ----------------
 * Chains are independent: rank g owns chains [g*batch/G, (g+1)*batch/G); the only collective is
 * the all-gather of posterior marginals at the end (the reference has no distributed path).
 * rxg_comm_unique_id fills a 128-byte NCCL id on rank 0; the host broadcasts it out of band.    */
int rxg_comm_unique_id(void* id128);
int rxg_comm_init(rxg_ctx*, int nranks, int rank, const void* id128);
/* Gather each rank's (mean[T][d][b_local], cov[T][d][d][b_local]) slab into
 * gathered_*[G][...] (rank-major, each slab contiguous).  Device pointers only.
 * With RXG_COV_REPLICATE (shared model on every rank, no ymask: the covariances do not depend on
 * the chain) only the means cross NVLink; gathered_cov[G][T][d][d][b_local]
 * is filled locally from post_cov (its first chain column, or the [T][d][d] table when
 * RXG_COV_SHARED_OUT is also set) at HBM write speed, concurrently with the gather.  Results are
 * bit-identical to the full gather.                                                              */
int rxg_allgather_posteriors(rxg_ctx*, int d, int T, int64_t batch_local, const float* post_mean,
                             const float* post_cov, float* gathered_mean, float* gathered_cov,
                             unsigned flags);


/* ------------------------------------------------------------------ peer-mapped gather ---------
 * The all-gather WITHOUT a collective launch: every rank maps its peers' gathered buffers (CUDA IPC over
 * NVLink / NVSwitch) and the fused smoothing sweep stores each smoothed posterior straight into all G
 * buffers while the backward recursion is still running ("issue ... finished time-slabs
 * ... while earlier slabs are still being computed" -- here at the granularity of one time step).
 * Host protocol (one process per GPU; the host exchanges the 64-byte handles out of band, e.g.
 * torch.distributed / MPI / a Julia Distributed channel):
 *   1. rxg_device_alloc the gathered buffers [G][T][d][b_local] (+ [G][T][d][d][b_local]) and one flag
 *      buffer of RXG_MAX_PEERS ints (zero it with rxg_device_memset);
 *   2. rxg_peer_export each, all-gather the handles, rxg_peer_open the G-1 remote ones;
 *   3. rxg_peer_group(nranks, rank, flag pointers as mapped here);
 *   4. rxg_lgssm_smooth_gather_f32 (or any compute + rxg_peer_allgather_f32) per step.
 * Ranks inside ONE process (several contexts) skip step 2's export/open and pass plain device pointers.   */
int rxg_device_alloc(rxg_ctx*, size_t bytes, void** dev_ptr);
int rxg_device_free(rxg_ctx*, void* dev_ptr);
int rxg_device_memset(rxg_ctx*, void* dev_ptr, int value, size_t bytes);
/* synchronous copies on the ctx stream -- for hosts without a CUDA binding of their own (the Julia shim, C hosts)  */
int rxg_memcpy_h2d(rxg_ctx*, void* dst_dev, const void* src_host, size_t bytes);
int rxg_memcpy_d2h(rxg_ctx*, void* dst_host, const void* src_dev, size_t bytes);
int rxg_peer_export(rxg_ctx*, const void* dev_ptr, void* handle64);
int rxg_peer_open(rxg_ctx*, const void* handle64, void** dev_ptr);
int rxg_peer_close(rxg_ctx*, void* dev_ptr);
/* flag_ptrs[g] = rank g's flag buffer as mapped in this process (own buffer for g == rank); nranks <= RXG_MAX_PEERS */
int rxg_peer_group(rxg_ctx*, int nranks, int rank, void* const* flag_ptrs);
/* device-side barrier over the group on the ctx stream (st.release.sys / ld.acquire.sys on the flags)     */
int rxg_peer_barrier(rxg_ctx*, unsigned flags);
/* generic all-gather of any posterior array (HGF outputs, filtered means, ...): `local` (n_local floats, may
 * already be the own slab gathered[rank] + rank * n_local) is stored into slab `rank` of every gathered[g],
 * then the barrier.  On completion gathered[rank] holds all G slabs.                                       */
int rxg_peer_allgather_f32(rxg_ctx*, int64_t n_local, const float* local, float* const* gathered, unsigned flags);
/* rxg_lgssm_smooth_f32 + the all-gather of its posteriors in one call: gathered_mean[g] / gathered_cov[g] are the
 * bases of rank g's [G][T][d][b_local] / [G][T][d][d][b_local] buffers as mapped here.  Shared-model gain-table
 * path: the sweep kernel itself stores to the peers (no second pass); other paths push their finished slab.
 * RXG_COV_REPLICATE (shared model, no mask): only the means cross NVLink, the other ranks' covariance slabs are
 * replicated locally from the [T][d][d] table concurrently with the sweep (gathered_cov[g != rank] may be NULL);
 * the buffers end up bit-identical to the full gather.  neg_log_evidence / status are local ([b_local]).
 * A caller must not start the next gather into the same buffers before every rank has consumed the result.   */
int rxg_lgssm_smooth_gather_f32(rxg_ctx*, int d, int m, int T, int64_t batch_local, const float* A,
                                const float* B, const float* P, const float* Q, const float* m0,
                                const float* S0, const float* u, const float* y, const uint8_t* ymask,
                                float* const* gathered_mean, float* const* gathered_cov,
                                float* neg_log_evidence, int32_t* status, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* RXGAUSS_H */
