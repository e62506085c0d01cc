// tcgen05 (UMMA) kernels of the large-state family.
//   umma_selftest_kernel : D[128 x 64] = A[128 x 128] * B[64 x 128]' with the 3xTF32 split -- validates the
//                          hand-written descriptors / TMEM plumbing against an fp64 product (tests/).
#include "rxg_internal.h"
#include "rxg_umma.cuh"

namespace rxg {

constexpr int UM_M = 128, UM_N = 64, UM_K = 128;
constexpr uint32_t UM_A_BYTES = UM_M * UM_K * 4, UM_B_BYTES = UM_N * UM_K * 4;

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sAhi = sm;
    uint8_t* sAlo = sAhi + UM_A_BYTES;
    uint8_t* sBhi = sAlo + UM_A_BYTES;
    uint8_t* sBlo = sBhi + UM_B_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sBlo + UM_B_BYTES);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

    for (int idx = tid; idx < UM_M * UM_K; idx += 128) {
        const int r = idx / UM_K, k = idx % UM_K;
        float hi, lo;
        umma::split_tf32(A[idx], hi, lo);
        *reinterpret_cast<float*>(sAhi + umma::elem_off(r, k, UM_K)) = hi;
        *reinterpret_cast<float*>(sAlo + umma::elem_off(r, k, UM_K)) = lo;
    }
    for (int idx = tid; idx < UM_N * UM_K; idx += 128) {
        const int r = idx / UM_K, k = idx % UM_K;
        float hi, lo;
        umma::split_tf32(B[idx], hi, lo);
        *reinterpret_cast<float*>(sBhi + umma::elem_off(r, k, UM_K)) = hi;
        *reinterpret_cast<float*>(sBlo + umma::elem_off(r, k, UM_K)) = lo;
    }
    if (warp == 0) umma::tmem_alloc(tptr, 64);
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
        const uint32_t idesc = umma::idesc_tf32(UM_M, UM_N);
        const uint32_t sbo = umma::sbo_bytes(UM_K);
        const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(sAhi), a_lo = (uint32_t)__cvta_generic_to_shared(sAlo);
        const uint32_t b_hi = (uint32_t)__cvta_generic_to_shared(sBhi), b_lo = (uint32_t)__cvta_generic_to_shared(sBlo);
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a0 = pass == 2 ? a_lo : a_hi, b0 = pass == 1 ? b_lo : b_hi;
            for (int kk = 0; kk < UM_K / 8; ++kk) {
                umma::mma_tf32(tmem, umma::smem_desc(a0 + kk * 2 * umma::LBO, umma::LBO, sbo),
                               umma::smem_desc(b0 + kk * 2 * umma::LBO, umma::LBO, sbo), idesc, acc);
                acc = 1;
            }
        }
        umma::commit(mbar);
    }
    umma::mbar_wait(mbar, 0);
    umma::fence_after();
    float v[32];
    const int row = warp * 32 + lane;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) Dout[row * UM_N + half * 32 + j] = v[j];
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}


// ------------------------------------------------------------------------------------------------
// lgssm_umma_sweep: the d = 64 mean sweep on the tensor pipe.  One CTA = 128 chains = the M rows of
// the UMMA A operand Z[128 x 128] = [x ; y_t] (forward) or [mu_f,t ; x_next] (backward), K-major;
// the B operand is the per-step gain block W_t[64 x 128] = [F_t | K_t] or [E_t | G_t], pre-split
// into tf32 hi / lo parts and pre-arranged in the canonical layout by large_gain_tables, so one
// cp.async stream brings it in.  D[128 x 64] (fp32, TMEM) is the new state: every thread owns one
// chain, reads its 64 new components with tcgen05.ld, stores them to HBM (coalesced over chains) and
// writes them back, hi/lo split, as next step's A operand.  Z = Zhi + Zlo, W = Whi + Wlo, and the
// product is Zhi Whi + Zhi Wlo + Zlo Whi (3xTF32) so the recursion keeps fp32-level accuracy.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cpa16u(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g));
}

template <bool SMOOTH>
__global__ void __launch_bounds__(128, 1)
lgssm_umma_sweep(const float* __restrict__ fwdU, const float* __restrict__ bwdU, const float* __restrict__ m0,
                 const float* __restrict__ m0c, const float* __restrict__ y, float* __restrict__ mean, int T,
                 int64_t batch) {
    constexpr int D = 64;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sZhi = sm;
    uint8_t* sZlo = sZhi + UM_A_BYTES;
    uint8_t* sWhi = sZlo + UM_A_BYTES;            // Whi and Wlo are contiguous (one 64 KB record per step)
    uint8_t* sWlo = sWhi + UM_B_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sWlo + UM_B_BYTES);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid / 32;
    const int64_t b0 = (int64_t)blockIdx.x * UM_M;
    const bool active = (b0 + tid) < batch;
    const int64_t bc = active ? b0 + tid : b0;    // inactive rows shadow the first chain (never stored)
    constexpr size_t REC = (size_t)2 * UM_N * UM_K;   // floats per per-step record (hi then lo)

    auto load_W = [&](const float* rec) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(rec);
        for (int p = tid; p < (int)(2 * UM_B_BYTES / 16); p += 128) cpa16u(sWhi + 16 * p, src + 16 * p);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // write 4 consecutive k of row `tid` (one 16-byte chunk of a core matrix), hi and lo parts
    auto put4 = [&](int k0, float a, float b, float c, float d) {
        float4 hi, lo;
        umma::split_tf32(a, hi.x, lo.x); umma::split_tf32(b, hi.y, lo.y);
        umma::split_tf32(c, hi.z, lo.z); umma::split_tf32(d, hi.w, lo.w);
        const uint32_t off = umma::elem_off(tid, k0, UM_K);
        *reinterpret_cast<float4*>(sZhi + off) = hi;
        *reinterpret_cast<float4*>(sZlo + off) = lo;
    };

    // Five fp32 accumulators of 64 columns each: the tensor pipe's adder truncates, so a single
    // accumulator over 48 MMAs drifts by ~3e-6 per step (measured: 2e-5 after the recursion).  The
    // hi*hi product is accumulated in four K-chunks (4 MMAs each) and the two cross terms in a
    // fifth; the five partial sums are added in fp32 registers (round-to-nearest).
    constexpr uint32_t NACC = 5, TCOLS = 512;
    if (warp == 0) umma::tmem_alloc(tptr, TCOLS);
    if (tid == 0) {
        umma::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_W(fwdU);
    float x[D], nxt[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { x[k] = m0c ? __ldg(m0c + (int64_t)k * batch + bc) : m0[k]; nxt[k] = __ldg(y + (int64_t)k * batch + bc); }
#pragma unroll
    for (int k = 0; k < D; k += 4) { put4(k, x[k], x[k + 1], x[k + 2], x[k + 3]); put4(D + k, nxt[k], nxt[k + 1], nxt[k + 2], nxt[k + 3]); }
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem = *tptr;
    const uint32_t idesc = umma::idesc_tf32(UM_M, UM_N), sbo = umma::sbo_bytes(UM_K);
    const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(sZhi), a_lo = (uint32_t)__cvta_generic_to_shared(sZlo);
    const uint32_t b_hi = (uint32_t)__cvta_generic_to_shared(sWhi), b_lo = (uint32_t)__cvta_generic_to_shared(sWlo);
    uint32_t phase = 0;

    auto step_mma = [&]() {     // operands are in place (generic writes + cp.async): publish, issue, commit
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        umma::fence_async_smem();
        umma::fence_before();
        __syncthreads();
        umma::fence_after();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < UM_K / 8; ++kk)          // hi * hi, accumulator kk / 4
                umma::mma_tf32(tmem + 64 * (kk / 4), umma::smem_desc(a_hi + kk * 2 * umma::LBO, umma::LBO, sbo),
                               umma::smem_desc(b_hi + kk * 2 * umma::LBO, umma::LBO, sbo), idesc, (kk % 4) != 0);
#pragma unroll
            for (int kk = 0; kk < UM_K / 8; ++kk)          // hi * lo
                umma::mma_tf32(tmem + 64 * 4, umma::smem_desc(a_hi + kk * 2 * umma::LBO, umma::LBO, sbo),
                               umma::smem_desc(b_lo + kk * 2 * umma::LBO, umma::LBO, sbo), idesc, kk != 0);
#pragma unroll
            for (int kk = 0; kk < UM_K / 8; ++kk)          // lo * hi
                umma::mma_tf32(tmem + 64 * 4, umma::smem_desc(a_lo + kk * 2 * umma::LBO, umma::LBO, sbo),
                               umma::smem_desc(b_hi + kk * 2 * umma::LBO, umma::LBO, sbo), idesc, 1u);
            umma::commit(mbar);
        }
    };
    auto read_state = [&]() {   // sum of the five accumulators, row of this thread's chain -> x[0..63]
        umma::mbar_wait(mbar, phase);
        phase ^= 1;
        umma::fence_after();
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        float v[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            umma::tmem_ld32(lane_base + 64 * 4 + 32 * half, x + 32 * half);          // cross terms first (smallest)
#pragma unroll
            for (uint32_t a = 0; a < NACC - 1; ++a) {
                umma::tmem_ld32(lane_base + 64 * a + 32 * half, v);
#pragma unroll
                for (int k = 0; k < 32; ++k) x[32 * half + k] += v[k];
            }
        }
        umma::fence_before();
    };

    // ---------------------------------------------------------------- forward: Z = [x ; y_t], W = [F_t | K_t]
    for (int t = 0; t < T; ++t) {
        step_mma();
        if (t + 1 < T) {
#pragma unroll
            for (int k = 0; k < D; ++k) nxt[k] = __ldg(y + ((int64_t)(t + 1) * D + k) * batch + bc);
        }
        read_state();
        if (t + 1 < T) load_W(fwdU + (size_t)(t + 1) * REC);
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (active) mean[((int64_t)t * D + k) * batch + bc] = x[k];
        if (t + 1 < T) {
#pragma unroll
            for (int k = 0; k < D; k += 4) { put4(k, x[k], x[k + 1], x[k + 2], x[k + 3]); put4(D + k, nxt[k], nxt[k + 1], nxt[k + 2], nxt[k + 3]); }
        }
    }
    if (SMOOTH) {
        // ------------------------------------------------------------ backward: Z = [mu_f,t ; x_next], W = [E_t | G_t]
        // x currently holds mu_f[T-1]; record T-1 is E = I, G = 0
        load_W(bwdU + (size_t)(T - 1) * REC);
#pragma unroll
        for (int k = 0; k < D; k += 4) { put4(k, x[k], x[k + 1], x[k + 2], x[k + 3]); put4(D + k, 0.f, 0.f, 0.f, 0.f); }
        for (int t = T - 1; t >= 0; --t) {
            step_mma();
            if (t - 1 >= 0) {
#pragma unroll
                for (int k = 0; k < D; ++k) nxt[k] = mean[((int64_t)(t - 1) * D + k) * batch + bc];
            }
            read_state();
            if (t - 1 >= 0) load_W(bwdU + (size_t)(t - 1) * REC);
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (active) mean[((int64_t)t * D + k) * batch + bc] = x[k];
            if (t - 1 >= 0) {
#pragma unroll
                for (int k = 0; k < D; k += 4) { put4(k, nxt[k], nxt[k + 1], nxt[k + 2], nxt[k + 3]); put4(D + k, x[k], x[k + 1], x[k + 2], x[k + 3]); }
            }
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

int launch_umma_sweep(rxg_ctx* ctx, bool smooth, const float* fwdU, const float* bwdU, const float* m0, const float* m0c,
                      const float* y, float* mean, int T, int64_t batch) {
    const size_t smem = 2 * UM_A_BYTES + 2 * UM_B_BYTES + 64;
    static bool done = false;
    if (!done) {
        RXG_CUDA(ctx, cudaFuncSetAttribute(lgssm_umma_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RXG_CUDA(ctx, cudaFuncSetAttribute(lgssm_umma_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    const unsigned blocks = (unsigned)((batch + UM_M - 1) / UM_M);
    if (smooth) lgssm_umma_sweep<true><<<blocks, 128, smem, ctx->stream>>>(fwdU, bwdU, m0, m0c, y, mean, T, batch);
    else        lgssm_umma_sweep<false><<<blocks, 128, smem, ctx->stream>>>(fwdU, bwdU, m0, m0c, y, mean, T, batch);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "lgssm_umma_sweep");
}

}  // namespace rxg

using namespace rxg;

extern "C" int rxg_selftest_umma_f32(rxg_ctx* ctx, const float* A, const float* B, float* D, unsigned flags) {
    if (!ctx) return RXG_ERR_BAD_ARG;
    if (!(flags & RXG_PTR_DEVICE) || !A || !B || !D) return fail(ctx, RXG_ERR_BAD_ARG, "selftest_umma: device pointers required");
    const size_t smem = 2 * UM_A_BYTES + 2 * UM_B_BYTES + 64;
    RXG_CUDA(ctx, cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, ctx->stream>>>(A, B, D);
    ctx->launches += 1;
    int rc = check_cuda(ctx, cudaGetLastError(), "umma_selftest_kernel");
    if (rc != RXG_OK) return rc;
    RXG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RXG_OK;
}
