// tcgen05 (UMMA) kernels of the large-state family (d = 16 / 32 / 64): the mean recursions of the shared-model
// LGSSM sweep on the 5th-generation tensor cores, hand-written PTX (rxg_umma.cuh).
//
//   umma_selftest_kernel<N, K> : D[128 x N] = A[128 x K] * B[N x K]' with the 3xTF32 split -- validates the
//                                descriptors / TMEM plumbing of every shape the sweep uses against an fp64 product.
//   umma_ky_kernel<D>          : u_t = K_t y_t for all (t, chain): no dependency along t, so it is a plain batched
//                                GEMM, parallel over time and chain tiles (off the recursion's critical path).
//   lgssm_umma_sweep<D, SMOOTH>: the dependent chain  x_t = F_t x_{t-1} + u_t  (forward) and
//                                mu_s[t] = v_t + G_t mu_s[t+1],  v_t = E_t x_t  (backward).  One CTA = 128 chains =
//                                the M rows of the A operand X[128 x D] (K-major); the forward B operand is the
//                                stacked block [F_t ; E_{t-1}] (N = 2D), so one MMA group yields both x_t and the
//                                backward pass's v_{t-1} from the same A operand.
// Operands are split tf32 hi / lo and combined as hi*hi + hi*lo + lo*hi (3xTF32), accumulated in several TMEM
// accumulators that are summed in fp32 registers (the tensor pipe's adder truncates).
// Per-step gain records are pre-arranged in the canonical UMMA layout by large_gain_tables and arrive with one
// TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) per step, double buffered.
#include "rxg_internal.h"
#include "rxg_umma.cuh"

namespace rxg {

template <int D>
struct US {
    static constexpr int KS = D / 8;                       // MMAs (K = 8) per pass over the K axis
    static constexpr int NHH = (KS + 3) / 4;               // hi*hi accumulators: at most 4 MMAs each
    static constexpr int NACC = NHH + 1;                   // + one for the two cross terms
    static constexpr uint32_t X_BYTES = 128u * D * 4u;     // one part (hi or lo) of the A operand
    static constexpr uint32_t G_BYTES = (uint32_t)D * D * 4u;       // one part of a D x D gain block
    static constexpr uint32_t FE_BYTES = 2u * G_BYTES;              // one part of [F ; E] (2D x D)
    static constexpr uint32_t SBO = umma::sbo_bytes(D);
    static constexpr int DC = 16;                          // state components per thread
    static constexpr int CG = D / DC;                      // column groups: warp w owns TMEM lanes 32 (w % 4) .. +31 and columns DC (w / 4) .. +DC-1
    static constexpr int NT = 128 * CG;                    // threads per CTA
    static constexpr uint32_t tcols(int n) {               // power of two >= NACC * n, >= 32
        uint32_t c = 32;
        while (c < (uint32_t)(NACC * n)) c <<= 1;
        return c;
    }
};

// ------------------------------------------------------------------------------------------------ self-test
template <int N, int K>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout) {
    constexpr uint32_t A_BYTES = 128u * K * 4u, B_BYTES = (uint32_t)N * K * 4u;
    constexpr uint32_t TC = N < 32 ? 32 : N;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sAhi = sm;
    uint8_t* sAlo = sAhi + A_BYTES;
    uint8_t* sBhi = sAlo + A_BYTES;
    uint8_t* sBlo = sBhi + B_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sBlo + B_BYTES);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid / 32;

    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        float hi, lo;
        umma::split_tf32(A[idx], hi, lo);
        *reinterpret_cast<float*>(sAhi + umma::elem_off(r, k, K)) = hi;
        *reinterpret_cast<float*>(sAlo + umma::elem_off(r, k, K)) = lo;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        float hi, lo;
        umma::split_tf32(B[idx], hi, lo);
        *reinterpret_cast<float*>(sBhi + umma::elem_off(r, k, K)) = hi;
        *reinterpret_cast<float*>(sBlo + umma::elem_off(r, k, K)) = lo;
    }
    if (warp == 0) umma::tmem_alloc(tptr, TC);
    if (tid == 0) {
        umma::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_async_smem();          // generic-proxy smem writes -> visible to the async (tensor) proxy
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem = *tptr;
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t sbo = umma::sbo_bytes(K);
        const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(sAhi), a_lo = (uint32_t)__cvta_generic_to_shared(sAlo);
        const uint32_t b_hi = (uint32_t)__cvta_generic_to_shared(sBhi), b_lo = (uint32_t)__cvta_generic_to_shared(sBlo);
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a0 = pass == 2 ? a_lo : a_hi, b0 = pass == 1 ? b_lo : b_hi;
            for (int kk = 0; kk < K / 8; ++kk) {
                umma::mma_tf32(tmem, umma::smem_desc(a0 + kk * 2 * umma::LBO, umma::LBO, sbo),
                               umma::smem_desc(b0 + kk * 2 * umma::LBO, umma::LBO, sbo), idesc, acc);
                acc = 1;
            }
        }
        umma::commit(mbar);
    }
    umma::mbar_wait_bounded(mbar, 0);
    umma::fence_after();
    constexpr int NC = N >= 32 ? 32 : 16;
    float v[NC];
    const int row = tid;
#pragma unroll
    for (int c = 0; c < N / NC; ++c) {
        umma::tmem_ldn<NC>(tmem + ((uint32_t)(warp * 32) << 16) + c * NC, v);
#pragma unroll
        for (int j = 0; j < NC; ++j) Dout[row * N + c * NC + j] = v[j];
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TC);
}

// ------------------------------------------------------------------------------------------------ shared pieces
// write this thread's DC components [c0, c0 + DC) of row `row` (its chain) of the A operand: split hi / lo, 16-byte
// pieces of the core matrices
template <int D>
__device__ __forceinline__ void put_cols(uint8_t* sHi, uint8_t* sLo, int row, int c0, const float* x) {
#pragma unroll
    for (int k0 = 0; k0 < US<D>::DC; k0 += 4) {
        float4 hi, lo;
        umma::split_tf32(x[k0], hi.x, lo.x); umma::split_tf32(x[k0 + 1], hi.y, lo.y);
        umma::split_tf32(x[k0 + 2], hi.z, lo.z); umma::split_tf32(x[k0 + 3], hi.w, lo.w);
        const uint32_t off = umma::elem_off(row, c0 + k0, D);
        *reinterpret_cast<float4*>(sHi + off) = hi;
        *reinterpret_cast<float4*>(sLo + off) = lo;
    }
}
// one thread: D_acc[128 x N] = X * W'  as hi*hi (NHH accumulators, <= 4 MMAs each) + hi*lo + lo*hi (one accumulator)
template <int D>
__device__ __forceinline__ void issue_3xtf32(uint32_t tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             int N, uint64_t* mbar) {
    using S = US<D>;
    const uint32_t idesc = umma::idesc_tf32(128, N);
#pragma unroll
    for (int kk = 0; kk < S::KS; ++kk)
        umma::mma_tf32(tmem + (uint32_t)N * (kk / 4), umma::smem_desc(a_hi + kk * 2 * umma::LBO, umma::LBO, S::SBO),
                       umma::smem_desc(b_hi + kk * 2 * umma::LBO, umma::LBO, S::SBO), idesc, (kk % 4) != 0);
#pragma unroll
    for (int kk = 0; kk < S::KS; ++kk)
        umma::mma_tf32(tmem + (uint32_t)N * S::NHH, umma::smem_desc(a_hi + kk * 2 * umma::LBO, umma::LBO, S::SBO),
                       umma::smem_desc(b_lo + kk * 2 * umma::LBO, umma::LBO, S::SBO), idesc, kk != 0);
#pragma unroll
    for (int kk = 0; kk < S::KS; ++kk)
        umma::mma_tf32(tmem + (uint32_t)N * S::NHH, umma::smem_desc(a_lo + kk * 2 * umma::LBO, umma::LBO, S::SBO),
                       umma::smem_desc(b_hi + kk * 2 * umma::LBO, umma::LBO, S::SBO), idesc, 1u);
    umma::commit(mbar);
}
// this thread's row, columns [col, col + DC), summed over the NACC accumulators in fp32 registers (cross terms
// first: smallest).  All loads are issued before the single wait.
template <int D>
__device__ __forceinline__ void read_acc(uint32_t lane_base, int N, int col, float* out) {
    using S = US<D>;
    uint32_t r[S::NACC][16];
#pragma unroll
    for (int a = 0; a < S::NACC; ++a) umma::tmem_ld16_issue(lane_base + (uint32_t)N * a + col, r[a]);
    umma::tmem_ld_wait();
#pragma unroll
    for (int a = 0; a < S::NACC; ++a) umma::tmem_ld_fence16(r[a]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float s = __uint_as_float(r[S::NHH][k]);
#pragma unroll
        for (int a = 0; a < S::NHH; ++a) s += __uint_as_float(r[a][k]);
        out[k] = s;
    }
}

// ------------------------------------------------------------------------------------------------ u_t = K_t y_t
// grid = (chain tiles, time slices); CTA = 128 chains x CG column groups (thread = one chain, DC components).
// recK[t] = K_t (D x D) hi | lo in the canonical layout.  u is written in the [T][D][batch] layout of the posterior
// means (the sweep consumes u_t from mean[t] before it overwrites that row).
template <int D>
__global__ void __launch_bounds__(US<D>::NT)
umma_ky_kernel(const float* __restrict__ recK, const float* __restrict__ y, float* __restrict__ u, int T, int64_t batch) {
    using S = US<D>;
    constexpr int DC = S::DC;
    constexpr uint32_t REC_BYTES = 2 * S::G_BYTES;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sYhi = sm;
    uint8_t* sYlo = sYhi + S::X_BYTES;
    uint8_t* sK0 = sYlo + S::X_BYTES;
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sK0 + 2 * REC_BYTES);
    uint64_t* full = mma_bar + 1;                 // [2]
    uint32_t* tptr = reinterpret_cast<uint32_t*>(full + 2);
    const int tid = threadIdx.x, warp = tid / 32;
    const int row = (warp & 3) * 32 + (tid & 31), c0 = (warp >> 2) * DC;
    const int64_t b0 = (int64_t)blockIdx.x * 128;
    const bool active = (b0 + row) < batch;
    const int64_t bc = active ? b0 + row : b0;
    const int t_lo = (int)(((int64_t)T * blockIdx.y) / gridDim.y), t_hi = (int)(((int64_t)T * (blockIdx.y + 1)) / gridDim.y);
    if (t_lo >= t_hi) return;
    constexpr uint32_t TCOLS = S::tcols(D);
    if (warp == 0) umma::tmem_alloc(tptr, TCOLS);
    if (tid == 0) {
        umma::mbar_init(mma_bar, 1); umma::mbar_init(full, 1); umma::mbar_init(full + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float yv[DC], yn[DC];
    const float* yp = y + (int64_t)c0 * batch + bc;
    float* up = u + (int64_t)c0 * batch + bc;
#pragma unroll
    for (int k = 0; k < DC; ++k) yv[k] = __ldg(yp + ((int64_t)t_lo * D + k) * batch);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem = *tptr;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(sYhi), a_lo = (uint32_t)__cvta_generic_to_shared(sYlo);
    const uint32_t k_base = (uint32_t)__cvta_generic_to_shared(sK0);
    constexpr size_t REC = (size_t)2 * D * D;
    if (tid == 0) {
        for (int j = 0; j < 2; ++j)
            if (t_lo + j < t_hi) {
                umma::mbar_expect_tx(full + j, REC_BYTES);
                umma::bulk_g2s(sK0 + j * REC_BYTES, recK + (size_t)(t_lo + j) * REC, REC_BYTES, full + j);
            }
    }
    uint32_t use0 = 0, use1 = 0, mstep = 0;
    for (int t = t_lo; t < t_hi; ++t) {
        const int buf = (t - t_lo) & 1;
        put_cols<D>(sYhi, sYlo, row, c0, yv);
        umma::fence_async_smem();
        umma::fence_before();
        __syncthreads();
        umma::fence_after();
        if (tid == 0) {
            const uint32_t par = buf ? (use1++ & 1) : (use0++ & 1);
            umma::mbar_wait_bounded(full + buf, par);
            const uint32_t b_hi = k_base + buf * REC_BYTES;
            issue_3xtf32<D>(tmem, a_hi, a_lo, b_hi, b_hi + S::G_BYTES, D, mma_bar);
        }
        if (t + 1 < t_hi) {
#pragma unroll
            for (int k = 0; k < DC; ++k) yn[k] = __ldg(yp + ((int64_t)(t + 1) * D + k) * batch);
        }
        umma::mbar_wait_bounded(mma_bar, mstep & 1);
        ++mstep;
        umma::fence_after();
        if (tid == 0 && t + 2 < t_hi) {          // the MMAs that read this K buffer are complete
            umma::mbar_expect_tx(full + buf, REC_BYTES);
            umma::bulk_g2s(sK0 + buf * REC_BYTES, recK + (size_t)(t + 2) * REC, REC_BYTES, full + buf);
        }
        float out[DC];
        read_acc<D>(lane_base, D, c0, out);
        if (active) {
#pragma unroll
            for (int k = 0; k < DC; ++k) up[((int64_t)t * D + k) * batch] = out[k];
        }
#pragma unroll
        for (int k = 0; k < DC; ++k) yv[k] = yn[k];
        umma::fence_before();
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

// ------------------------------------------------------------------------------------------------ the recursion
// mean[t] holds u_t on entry.  Forward step t: D = X_{t-1} [F_t ; E_{t-1}]'  ->  x_t = D[:, :D] + u_t (next A
// operand; stored as the filtered mean when !SMOOTH), v_{t-1} = D[:, D:] (stored into mean[t-1]).  Backward step t:
// mu_s[t] = v_t + mu_s[t+1] G_t' (stored into mean[t]; next A operand).  mu_s[T-1] = x_{T-1}.
// CTA = 128 chains x CG column groups: a thread owns DC = 16 components of one chain, so the per-step epilogue
// (TMEM -> registers -> split -> shared memory, global loads / stores) is spread over 4 CG warps.
template <int D, bool SMOOTH>
__global__ void __launch_bounds__(US<D>::NT)
lgssm_umma_sweep(const float* __restrict__ recFE, const float* __restrict__ recG, const float* __restrict__ m0,
                 const float* __restrict__ m0c, float* mean, int T, int64_t batch) {
    using S = US<D>;
    constexpr int DC = S::DC;
    constexpr int NF = SMOOTH ? 2 * D : D;                       // forward N
    constexpr uint32_t FE_REC_BYTES = 2 * S::FE_BYTES, G_REC_BYTES = 2 * S::G_BYTES;
    constexpr size_t FE_REC = (size_t)4 * D * D, G_REC = (size_t)2 * D * D;   // floats per record
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sXhi = sm;
    uint8_t* sXlo = sXhi + S::X_BYTES;
    uint8_t* sB0 = sXlo + S::X_BYTES;              // two buffers of FE_REC_BYTES
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sB0 + 2 * FE_REC_BYTES);
    uint64_t* full = mma_bar + 1;                 // [2]
    uint32_t* tptr = reinterpret_cast<uint32_t*>(full + 2);
    const int tid = threadIdx.x, warp = tid / 32;
    const int row = (warp & 3) * 32 + (tid & 31), c0 = (warp >> 2) * DC;
    const int64_t b0 = (int64_t)blockIdx.x * 128;
    const bool active = (b0 + row) < batch;
    const int64_t bc = active ? b0 + row : b0;    // inactive rows shadow the tile's first chain (never stored)
    constexpr uint32_t TCOLS = S::tcols(NF);
    if (warp == 0) umma::tmem_alloc(tptr, TCOLS);
    if (tid == 0) {
        umma::mbar_init(mma_bar, 1); umma::mbar_init(full, 1); umma::mbar_init(full + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float* mp = mean + (int64_t)c0 * batch + bc;  // this thread's components of step 0
    const int64_t tstride = (int64_t)D * batch;
    float x[DC], cur[DC], nxt[DC];
#pragma unroll
    for (int k = 0; k < DC; ++k) {
        x[k] = m0c ? __ldg(m0c + (int64_t)(c0 + k) * batch + bc) : m0[c0 + k];
        cur[k] = mp[(int64_t)k * batch];                         // u_0
    }
    put_cols<D>(sXhi, sXlo, row, c0, x);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem = *tptr;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(sXhi), a_lo = (uint32_t)__cvta_generic_to_shared(sXlo);
    const uint32_t b_base = (uint32_t)__cvta_generic_to_shared(sB0);
    if (tid == 0) {
        for (int j = 0; j < 2; ++j)
            if (j < T) {
                umma::mbar_expect_tx(full + j, FE_REC_BYTES);
                umma::bulk_g2s(sB0 + j * FE_REC_BYTES, recFE + (size_t)j * FE_REC, FE_REC_BYTES, full + j);
            }
    }
    uint32_t use0 = 0, use1 = 0, mstep = 0;
    // ---------------------------------------------------------------- forward
    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        umma::fence_async_smem();                  // this thread's A-operand pieces -> visible to the tensor (async) proxy
        umma::fence_before();
        __syncthreads();
        umma::fence_after();
        if (tid == 0) {
            const uint32_t par = buf ? (use1++ & 1) : (use0++ & 1);
            umma::mbar_wait_bounded(full + buf, par);
            const uint32_t b_hi = b_base + buf * FE_REC_BYTES;
            issue_3xtf32<D>(tmem, a_hi, a_lo, b_hi, b_hi + S::FE_BYTES, NF, mma_bar);
        }
        if (t + 1 < T) {
#pragma unroll
            for (int k = 0; k < DC; ++k) nxt[k] = mp[(t + 1) * tstride + (int64_t)k * batch];     // u_{t+1}
        }
        umma::mbar_wait_bounded(mma_bar, mstep & 1);
        ++mstep;
        umma::fence_after();
        if (tid == 0 && t + 2 < T) {               // this B buffer is free again: the MMAs of step t are complete
            umma::mbar_expect_tx(full + buf, FE_REC_BYTES);
            umma::bulk_g2s(sB0 + buf * FE_REC_BYTES, recFE + (size_t)(t + 2) * FE_REC, FE_REC_BYTES, full + buf);
        }
        read_acc<D>(lane_base, NF, c0, x);
#pragma unroll
        for (int k = 0; k < DC; ++k) x[k] += cur[k];
        put_cols<D>(sXhi, sXlo, row, c0, x);       // next step's A operand (and the backward pass's first one)
        if (!SMOOTH) {
            if (active) {
#pragma unroll
                for (int k = 0; k < DC; ++k) mp[t * tstride + (int64_t)k * batch] = x[k];
            }
        } else if (t >= 1) {
            float v[DC];
            read_acc<D>(lane_base, NF, D + c0, v); // v_{t-1} = E_{t-1} x_{t-1}
            if (active) {
#pragma unroll
                for (int k = 0; k < DC; ++k) mp[(t - 1) * tstride + (int64_t)k * batch] = v[k];
            }
        }
#pragma unroll
        for (int k = 0; k < DC; ++k) cur[k] = nxt[k];
        umma::fence_before();
    }
    if (SMOOTH) {
        // ------------------------------------------------------------ backward
        if (active) {
#pragma unroll
            for (int k = 0; k < DC; ++k) mp[(T - 1) * tstride + (int64_t)k * batch] = x[k];        // mu_s[T-1] = x_{T-1}
        }
        if (tid == 0) {
            for (int j = 0; j < 2; ++j)
                if (T - 2 - j >= 0) {
                    umma::mbar_expect_tx(full + j, G_REC_BYTES);
                    umma::bulk_g2s(sB0 + j * FE_REC_BYTES, recG + (size_t)(T - 2 - j) * G_REC, G_REC_BYTES, full + j);
                }
        }
        if (T >= 2) {
#pragma unroll
            for (int k = 0; k < DC; ++k) cur[k] = mp[(T - 2) * tstride + (int64_t)k * batch];      // v_{T-2}
        }
        for (int r = 0; T - 2 - r >= 0; ++r) {
            const int t = T - 2 - r, buf = r & 1;
            umma::fence_async_smem();
            umma::fence_before();
            __syncthreads();
            umma::fence_after();
            if (tid == 0) {
                const uint32_t par = buf ? (use1++ & 1) : (use0++ & 1);
                umma::mbar_wait_bounded(full + buf, par);
                const uint32_t b_hi = b_base + buf * FE_REC_BYTES;
                issue_3xtf32<D>(tmem, a_hi, a_lo, b_hi, b_hi + S::G_BYTES, D, mma_bar);
            }
            if (t - 1 >= 0) {
#pragma unroll
                for (int k = 0; k < DC; ++k) nxt[k] = mp[(t - 1) * tstride + (int64_t)k * batch];  // v_{t-1}
            }
            umma::mbar_wait_bounded(mma_bar, mstep & 1);
            ++mstep;
            umma::fence_after();
            if (tid == 0 && t - 2 >= 0) {
                umma::mbar_expect_tx(full + buf, G_REC_BYTES);
                umma::bulk_g2s(sB0 + buf * FE_REC_BYTES, recG + (size_t)(t - 2) * G_REC, G_REC_BYTES, full + buf);
            }
            read_acc<D>(lane_base, D, c0, x);
#pragma unroll
            for (int k = 0; k < DC; ++k) x[k] += cur[k];
            if (t > 0) put_cols<D>(sXhi, sXlo, row, c0, x);
            if (active) {
#pragma unroll
                for (int k = 0; k < DC; ++k) mp[t * tstride + (int64_t)k * batch] = x[k];
            }
#pragma unroll
            for (int k = 0; k < DC; ++k) cur[k] = nxt[k];
            umma::fence_before();
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

template <int D>
static int launch_umma_sweep_d(rxg_ctx* ctx, bool smooth, const float* recFE, const float* recG, const float* recK,
                               const float* m0, const float* m0c, const float* y, float* mean, int T, int64_t batch) {
    using S = US<D>;
    const size_t smem_ky = 2 * S::X_BYTES + 4 * S::G_BYTES + 64;
    const size_t smem_sw = 2 * S::X_BYTES + 4 * S::FE_BYTES + 64;
    {   // per-device attributes: set on every call
        RXG_CUDA(ctx, cudaFuncSetAttribute(umma_ky_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ky));
        RXG_CUDA(ctx, cudaFuncSetAttribute(lgssm_umma_sweep<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sw));
        RXG_CUDA(ctx, cudaFuncSetAttribute(lgssm_umma_sweep<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sw));
    }
    const unsigned tiles = (unsigned)((batch + 127) / 128);
    // time slices of the K y pre-pass: about two CTAs per SM in flight overall
    int tsplit = (int)((2 * (unsigned)ctx->sm_count + tiles - 1) / tiles);
    if (tsplit < 1) tsplit = 1;
    if (tsplit > T) tsplit = T;
    umma_ky_kernel<D><<<dim3(tiles, (unsigned)tsplit), S::NT, smem_ky, ctx->stream>>>(recK, y, mean, T, batch);
    if (smooth) lgssm_umma_sweep<D, true><<<tiles, S::NT, smem_sw, ctx->stream>>>(recFE, recG, m0, m0c, mean, T, batch);
    else        lgssm_umma_sweep<D, false><<<tiles, S::NT, smem_sw, ctx->stream>>>(recFE, recG, m0, m0c, mean, T, batch);
    ctx->launches += 2;
    return check_cuda(ctx, cudaGetLastError(), "lgssm_umma_sweep");
}

int launch_umma_sweep(rxg_ctx* ctx, int d, bool smooth, const float* recFE, const float* recG, const float* recK,
                      const float* m0, const float* m0c, const float* y, float* mean, int T, int64_t batch) {
    switch (d) {
        case 16: return launch_umma_sweep_d<16>(ctx, smooth, recFE, recG, recK, m0, m0c, y, mean, T, batch);
        case 32: return launch_umma_sweep_d<32>(ctx, smooth, recFE, recG, recK, m0, m0c, y, mean, T, batch);
        case 64: return launch_umma_sweep_d<64>(ctx, smooth, recFE, recG, recK, m0, m0c, y, mean, T, batch);
        default: return fail(ctx, RXG_ERR_UNSUPPORTED, "umma sweep: d=%d", d);
    }
}

template <int N, int K>
static int run_selftest(rxg_ctx* ctx, const float* A, const float* B, float* D) {
    const size_t smem = 2 * (size_t)128 * K * 4 + 2 * (size_t)N * K * 4 + 64;
    RXG_CUDA(ctx, cudaFuncSetAttribute(umma_selftest_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<N, K><<<1, 128, smem, ctx->stream>>>(A, B, D);
    ctx->launches += 1;
    int rc = check_cuda(ctx, cudaGetLastError(), "umma_selftest_kernel");
    if (rc != RXG_OK) return rc;
    RXG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RXG_OK;
}

}  // namespace rxg

using namespace rxg;

extern "C" int rxg_selftest_umma_shape_f32(rxg_ctx* ctx, int n, int k, const float* A, const float* B, float* D,
                                           unsigned flags) {
    if (!ctx) return RXG_ERR_BAD_ARG;
    if (!(flags & RXG_PTR_DEVICE) || !A || !B || !D) return fail(ctx, RXG_ERR_BAD_ARG, "selftest_umma: device pointers required");
    RXG_CUDA(ctx, cudaSetDevice(ctx->device));
    switch (n * 1000 + k) {
        case 64 * 1000 + 128: return run_selftest<64, 128>(ctx, A, B, D);
        case 128 * 1000 + 64: return run_selftest<128, 64>(ctx, A, B, D);
        case 64 * 1000 + 64: return run_selftest<64, 64>(ctx, A, B, D);
        case 64 * 1000 + 32: return run_selftest<64, 32>(ctx, A, B, D);
        case 32 * 1000 + 32: return run_selftest<32, 32>(ctx, A, B, D);
        case 32 * 1000 + 16: return run_selftest<32, 16>(ctx, A, B, D);
        case 16 * 1000 + 16: return run_selftest<16, 16>(ctx, A, B, D);
        default: return fail(ctx, RXG_ERR_UNSUPPORTED, "selftest_umma: shape (N=%d, K=%d) is not one the sweeps use", n, k);
    }
}

extern "C" int rxg_selftest_umma_f32(rxg_ctx* ctx, const float* A, const float* B, float* D, unsigned flags) {
    return rxg_selftest_umma_shape_f32(ctx, 64, 128, A, B, D, flags);
}
