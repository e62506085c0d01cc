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
