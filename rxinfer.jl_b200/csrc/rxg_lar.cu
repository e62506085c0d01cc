// rxg_lar_vmp_f32: the latent autoregressive model of /root/reference/test/models/autoregressive/lar_tests.jl as one
// launch over a batch of series (kernel body: rxg_lar.cuh).  SURVEY.md section 8(f)-3.
#include "rxg_internal.h"
#include "rxg_lar.cuh"

namespace rxg {

template <int P>
__global__ void __launch_bounds__(64)
lar_vmp_kernel(const float* __restrict__ y, int T, int64_t batch, int iters, lar::Params prm, float* __restrict__ ws,
               float* __restrict__ x_mean, float* __restrict__ x_cov, float* __restrict__ th_mean,
               float* __restrict__ th_cov, float* __restrict__ g_shape, float* __restrict__ g_rate,
               double* __restrict__ fe, int32_t* __restrict__ status) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const bool bad = lar::chain<P>(b, batch, y, T, iters, prm, ws, x_mean, x_cov, th_mean, th_cov, g_shape, g_rate, fe);
    if (status) status[b] = bad ? RXG_ERR_NOT_SPD : RXG_OK;
}

}  // namespace rxg

extern "C" int rxg_lar_vmp_f32(rxg_ctx* ctx, int order, int T, int64_t batch, int iterations, const float* params,
                               const float* y, float* x_mean, float* x_cov, float* theta_mean, float* theta_cov,
                               float* gamma_shape, float* gamma_rate, double* free_energy, int32_t* status,
                               unsigned flags) {
    if (!ctx) return RXG_ERR_BAD_ARG;
    if (!(flags & RXG_PTR_DEVICE)) return rxg::fail(ctx, RXG_ERR_UNSUPPORTED, "lar_vmp takes device pointers");
    if (order < 1 || order > 6) return rxg::fail(ctx, RXG_ERR_UNSUPPORTED, "lar_vmp: order=%d unsupported (1-6)", order);
    if (T < 1 || batch < 1 || iterations < 1 || !params || !y || !theta_mean || !theta_cov || !gamma_shape || !gamma_rate)
        return rxg::fail(ctx, RXG_ERR_BAD_ARG, "lar_vmp: bad argument");
    for (int i = 0; i < 8; ++i)
        if (!(params[i] > 0.f)) return rxg::fail(ctx, RXG_ERR_BAD_ARG, "lar_vmp: params[%d] must be positive", i);
    rxg::lar::Params prm;
    prm.tau = params[0]; prm.a0 = params[1]; prm.b0 = params[2]; prm.w0 = params[3]; prm.p0 = params[4];
    prm.init_shape = params[5]; prm.init_rate = params[6]; prm.init_theta_prec = params[7];
    RXG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ns = (size_t)order + (size_t)order * (order + 1) / 2;
    float* ws = (float*)rxg::workspace(ctx, (size_t)T * ns * (size_t)batch * sizeof(float));
    if (!ws) return RXG_ERR_CUDA;
    const unsigned grid = (unsigned)((batch + 63) / 64);
    switch (order) {
#define RXG_LAR(PP) case PP: rxg::lar_vmp_kernel<PP><<<grid, 64, 0, ctx->stream>>>(y, T, batch, iterations, prm, ws, x_mean, x_cov, theta_mean, theta_cov, gamma_shape, gamma_rate, free_energy, status); break;
        RXG_LAR(1) RXG_LAR(2) RXG_LAR(3) RXG_LAR(4) RXG_LAR(5) RXG_LAR(6)
#undef RXG_LAR
    }
    ctx->launches += 1;
    {
        int rc = rxg::check_cuda(ctx, cudaGetLastError(), "lar_vmp_kernel");
        if (rc != RXG_OK) return rc;
    }
    if (!(flags & RXG_ASYNC)) RXG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RXG_OK;
}
