// Test harness (NOT part of librxgauss.so): compiles the body of the latent-AR kernel (csrc/rxg_lar.cuh, __host__ __device__)
// for the host so that the exact code the GPU runs can be checked against the fp64 oracle without a GPU
// (tests/test_lar.py).  The product path has no CPU route: rxg_lar_vmp_f32 launches the CUDA kernel or fails.
#include <cuda_runtime.h>
#include "../../rxinfer.jl_b200/csrc/rxg_lar.cuh"

extern "C" int lar_host_run(int order, int T, long long batch, int iters, const float* params, const float* y, float* ws,
                            float* x_mean, float* x_cov, float* th_mean, float* th_cov, float* g_shape, float* g_rate,
                            double* fe, int* status) {
    rxg::lar::Params prm;
    prm.tau = params[0]; prm.a0 = params[1]; prm.b0 = params[2]; prm.w0 = params[3]; prm.p0 = params[4];
    prm.init_shape = params[5]; prm.init_rate = params[6]; prm.init_theta_prec = params[7];
    for (long long b = 0; b < batch; ++b) {
        bool bad = false;
        switch (order) {
#define CASE(PP) case PP: bad = rxg::lar::chain<PP>(b, batch, y, T, iters, prm, ws, x_mean, x_cov, th_mean, th_cov, g_shape, g_rate, fe); break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
            default: return -1;
        }
        if (status) status[b] = bad ? 1 : 0;
    }
    return 0;
}
