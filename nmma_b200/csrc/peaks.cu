// Roofline denominators measured on the device the engine runs on (include/nmma_b200.h, "knobs / introspection"):
//   nmma_b200_tf32_peak   dense tcgen05 kind::tf32 rate (the pipe fused_tc_logl_kernel's contraction runs on)
//   nmma_b200_dfma_peak   fp64 FMA rate (the pipe the GP front end's kernel values run on)
// Both time a dependent-free instruction stream on every SM with CUDA events on the handle's own stream.
#include <algorithm>

#include "handle.h"
#include "tc_kernel.cuh"

namespace nmma {

// One CTA per SM, one elected lane issues `iters` x 2 MMAs D[128 x 128] += A[128 x 8] . B[128 x 8]^T (A in TMEM, B in
// shared memory, K-major, no swizzle) into two alternating accumulators: 2 * 128 * 128 * 8 FLOP per MMA.
constexpr int kPeakN = 128;
__global__ void __launch_bounds__(128) tf32_peak_kernel(int iters) {
    __shared__ __align__(128) float sB[kPeakN * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < kPeakN * 8; i += 128) sB[i] = 1.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __float_as_uint(1.0f);
    tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 2 * kPeakN, a);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        const uint64_t db = tc_smem_desc(smem_u32(sB), kPeakN * 16, 128);
        const uint32_t lo = (uint32_t)db, hi = (uint32_t)(db >> 32);
        constexpr uint32_t idesc = tc_idesc(kPeakN);
        if (elect_one()) {
            for (int r = 0; r < iters; ++r) {
                mma_tf32_ts(tmem, tmem + 2 * kPeakN, lo, hi, idesc, r > 0 ? 1u : 0u);
                mma_tf32_ts(tmem + kPeakN, tmem + 2 * kPeakN, lo, hi, idesc, r > 0 ? 1u : 0u);
            }
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double* sink) {
    constexpr int CH = 8;
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = seed + (double)(threadIdx.x + i);
    const double m = 0.999 + seed * 1e-12, c = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace nmma

using namespace nmma;

namespace {
template <typename Launch>
int time_best(nmma_b200_t* h, Launch launch, double work, double* out) {
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, h->own_stream));
        launch();
        CU(cudaGetLastError());
        CU(cudaEventRecord(e1, h->own_stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        h->launches += 1;
        if (rep > 0) best = std::max(best, work / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out = best;
    return NMMA_B200_OK;
}
}  // namespace

extern "C" {

int nmma_b200_tf32_peak(nmma_b200_t* h, int iters, double* flops_per_s) {
    if (!h || !flops_per_s || iters < 1) return NMMA_B200_ERR_ARG;
    CU(cudaSetDevice(h->device));
    const unsigned grid = (unsigned)h->sm_count;
    const double work = 2.0 * (double)iters * 2.0 * 128.0 * kPeakN * 8.0 * grid;
    return time_best(h, [&] { tf32_peak_kernel<<<grid, 128, 0, h->own_stream>>>(iters); }, work, flops_per_s);
}

int nmma_b200_dfma_peak(nmma_b200_t* h, int iters, double* flops_per_s) {
    if (!h || !flops_per_s || iters < 1) return NMMA_B200_ERR_ARG;
    CU(cudaSetDevice(h->device));
    double* sink = nullptr;
    CU(cudaMalloc((void**)&sink, sizeof(double)));
    const unsigned grid = (unsigned)h->sm_count * 8;
    const double work = 2.0 * 8.0 * (double)iters * 256.0 * grid;
    const int rc = time_best(h, [&] { dfma_peak_kernel<<<grid, 256, 0, h->own_stream>>>(iters, 1.0, sink); }, work, flops_per_s);
    cudaFree(sink);
    return rc;
}

}  // extern "C"
