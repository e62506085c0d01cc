// Host side of host-pointer calls with a shared model: the per-chain posterior covariances the contract asks for
// are chain independent, so only the [T][d][d] table crosses PCIe and the per-chain copies are written into the
// caller's buffer by host threads (4 d^2 of the 4 (d + d^2) bytes per (chain, step) never touch the link).
// This file is plain C++ (g++, no CUDA): non-temporal AVX-512 / AVX2 stores selected at run time, worker threads
// pinned to the NUMA node that holds the destination pages (a cross-socket fill runs at UPI speed, not DRAM speed).
// [ref: the reference materialises posteriors[:x] as T heap objects per chain, src/inference/batch.jl:475-481.]
#include <immintrin.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <thread>
#include <vector>

namespace rxg {

namespace {

__attribute__((target("avx512f"))) void fill_row_512(float* dst, int64_t n, float v) {
    int64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) dst[i++] = v;
    const __m512 vv = _mm512_set1_ps(v);
    for (; i + 64 <= n; i += 64) {        // four cache lines per iteration
        _mm512_stream_ps(dst + i, vv);
        _mm512_stream_ps(dst + i + 16, vv);
        _mm512_stream_ps(dst + i + 32, vv);
        _mm512_stream_ps(dst + i + 48, vv);
    }
    for (; i + 16 <= n; i += 16) _mm512_stream_ps(dst + i, vv);
    for (; i < n; ++i) dst[i] = v;
}

__attribute__((target("avx2"))) void fill_row_256(float* dst, int64_t n, float v) {
    int64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) dst[i++] = v;
    const __m256 vv = _mm256_set1_ps(v);
    for (; i + 8 <= n; i += 8) _mm256_stream_ps(dst + i, vv);
    for (; i < n; ++i) dst[i] = v;
}

void fill_row_scalar(float* dst, int64_t n, float v) {
    for (int64_t i = 0; i < n; ++i) dst[i] = v;
}

typedef void (*fill_fn)(float*, int64_t, float);

fill_fn pick_fill() {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f")) return fill_row_512;
    if (__builtin_cpu_supports("avx2")) return fill_row_256;
    return fill_row_scalar;
}

// NUMA node that backs the page of `p` (move_pages with a null target list only queries), or -1
int node_of(const void* p) {
    void* page = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)4095);
    int status = -1;
    const long rc = syscall(SYS_move_pages, 0, 1UL, &page, nullptr, &status, 0);
    return (rc == 0 && status >= 0) ? status : -1;
}

// CPUs of a NUMA node that this process may run on
bool node_cpus(int node, cpu_set_t* out) {
    char path[128], buf[4096];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return false;
    CPU_ZERO(out);
    int n = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo = 0, hi = 0;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) { /* range */ }
        else if (sscanf(tok, "%d", &lo) == 1) hi = lo;
        else continue;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, out); ++n; }
    }
    return n > 0;
}

}  // namespace

// cov[row][b] = tab[row] for rows [0, rows): broadcast on the HOST side of the PCIe link
void host_broadcast_cov(float* cov, const float* tab, int64_t rows, int64_t batch, int nthreads) {
    static const fill_fn fill = pick_fill();
    if (nthreads < 1) nthreads = 1;
    // pin the workers to the node of the destination when the whole buffer sits on one node
    cpu_set_t pin;
    bool do_pin = false;
    {
        const int n0 = node_of(cov), n1 = node_of(cov + (rows * batch) / 2), n2 = node_of(cov + rows * batch - 1);
        if (n0 >= 0 && n0 == n1 && n1 == n2) do_pin = node_cpus(n0, &pin);
    }
    auto work = [=](int tid) {
        if (do_pin) sched_setaffinity(0, sizeof(pin), &pin);     // worker threads only: the caller's thread is never re-pinned
        const int64_t lo = rows * tid / nthreads, hi = rows * (tid + 1) / nthreads;
        for (int64_t r = lo; r < hi; ++r) fill(cov + r * batch, batch, tab[r]);
        _mm_sfence();
    };
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
}

}  // namespace rxg

// stand-alone bandwidth probe of the fill (bench_extra / tuning): bytes, threads -> GB/s
extern "C" double rxg_selftest_host_fill_gbs(float* dst, int64_t rows, int64_t batch, int nthreads, int reps) {
    std::vector<float> tab((size_t)rows, 1.5f);
    rxg::host_broadcast_cov(dst, tab.data(), rows, batch, nthreads);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int i = 0; i < reps; ++i) rxg::host_broadcast_cov(dst, tab.data(), rows, batch, nthreads);
    clock_gettime(CLOCK_MONOTONIC, &b);
    const double s = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
    return (double)rows * batch * 4 * reps / s / 1e9;
}
