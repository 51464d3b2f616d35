// tcgen05 probe: validates the TMEM / shared-memory operand layouts, the instruction and matrix descriptors and the
// rounding behaviour that nmma_b200/csrc/tc_kernel.cuh relies on, and times kind::tf32 MMAs of the shapes it issues.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
//
// Test 1: D[128 x N] = A[128 x 8] . B0[N x 8]^T + A . B1[N x 8]^T, A in TMEM (written with tcgen05.st, one row per
//         thread), B in shared memory (K-major, no swizzle, core matrix = 8 rows x 16 B), accumulators in TMEM.
// Test 2: which rounding the tensor core applies to fp32 operands of kind::tf32 (truncate vs round-to-nearest) and to
//         the fp32 accumulator (sum of 2048 products vs an fp64 reference).
// Test 3: cycles per MMA for N = 16 / 32 / 64 (A from TMEM).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- tcgen05 wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T, kind::tf32
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// K-major, no-swizzle matrix descriptor: core matrix = 8 rows x 16 B, rows 16 B apart; `lbo` = byte distance between
// the two 16-byte K halves, `sbo` = byte distance between 8-row groups (cute::UMMA::SmemDescriptor, version 1).
__host__ __device__ inline uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128 (cute::UMMA::InstrDescriptor)
__host__ __device__ inline uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// B tile in shared memory for an N x 8 operand: float index of element (n, k)
__host__ __device__ inline int b_index(int N, int n, int k) { return (k >> 2) * (N * 4) + n * 4 + (k & 3); }

constexpr int MAXN = 64;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B0,
                                                    const float* __restrict__ B1, int N, int reps, float* __restrict__ D,
                                                    long long* __restrict__ cycles) {
    __shared__ __align__(128) float sB[2][MAXN * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int n = i / 8, k = i % 8;
        sB[0][b_index(N, n, k)] = B0[i];
        sB[1][b_index(N, n, k)] = B1[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy (MMA) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t colD = 0, colA = 64;
    // A row of this thread -> TMEM columns colA..colA+7 of lane tid
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __float_as_uint(A[tid * 8 + k]);
    tmem_st8(tmem + lane_base + colA, a);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(N);
        const uint64_t d0 = make_desc(smem_u32(sB[0]), N * 16, 128), d1 = make_desc(smem_u32(sB[1]), N * 16, 128);
        t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
                mma_tf32_ts(tmem + colD, tmem + colA, d0, idesc, r > 0 ? 1u : 0u);
                mma_tf32_ts(tmem + colD, tmem + colA, d1, idesc, 1u);
            }
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    if (tid == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + colD + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}


// Timing variants: `nacc` independent accumulators used round-robin (dependent-accumulate latency vs issue rate),
// A from TMEM (ts = 1) or from shared memory (ts = 0).
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The issuing warp enters a warp-uniform branch and elects one lane (the pattern ptxas maps onto a plain UTCHMMA
// stream; a lane-divergent `if (tid == 0)` makes it wrap every MMA in an ELECT/BRA.U.ANY loop).
template <int N, int NACC, int TS>
__global__ void __launch_bounds__(128) timing_kernel(int reps, int nwarps_issue, long long* __restrict__ cycles) {
    __shared__ __align__(128) float sB[MAXN * 8];
    __shared__ __align__(128) float sA[128 * 8];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < MAXN * 8; i += 128) sB[i] = 1.0f;
    for (int i = tid; i < 128 * 8; i += 128) sA[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __float_as_uint(1.0f);
    tmem_st8(tmem + lane_base + 448, a);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp < nwarps_issue) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(N);
        const uint64_t db = make_desc(smem_u32(sB), N * 16, 128);
        const uint64_t da = make_desc(smem_u32(sA), 128 * 16, 128);
        const uint32_t dbase = tmem + warp * 96;
        const long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int q = 0; q < NACC; ++q) {
                    if (TS) mma_tf32_ts(dbase + q * N, tmem + 448, db, idesc, r > 0 ? 1u : 0u);
                    else mma_tf32_ss(dbase + q * N, da, db, idesc, r > 0 ? 1u : 0u);
                }
            }
            tc_commit(&bar[warp]);
        }
        __syncwarp();
        mbar_wait(&bar[warp], 0);
        if ((tid & 31) == 0) cycles[warp] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// kind::f16 (K = 16) timing and the per-chunk MMA pattern of fused_tc_logl_kernel (round 2, fp16 split operands):
// per 64-hidden chunk 4 x (N = 32 MMA + N = 16 MMA into the same accumulator), commit, commit, one N = 64 MMA, commit.
// Reports the cycles the elected lane needs to ISSUE the pattern and the cycles until the last commit has arrived.
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ inline uint32_t make_idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
template <int MODE>   // 0: N=16 stream, 1: N=32 stream, 2: N=64 stream, 3: chunk pattern with commits, 4: chunk pattern, one commit
__global__ void __launch_bounds__(128) f16_kernel(int reps, int nwarps_issue, long long* __restrict__ cycles) {
    __shared__ __align__(128) float sB[MAXN * 8 * 4];
    __shared__ __align__(8) uint64_t bar[8];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < MAXN * 8 * 4; i += 128) sB[i] = 0.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0u;
    for (int c0 = 0; c0 < 64; c0 += 8) tmem_st8(tmem + lane_base + 384 + c0, a);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp < nwarps_issue) {
        tc_fence_after();
        const uint32_t dbase = tmem + warp * 128;
        const uint32_t abase = tmem + 384;
        const long long t0 = clock64();
        long long t1 = t0;
        if (elect_one()) {
            if (MODE < 3) {
                constexpr int N = MODE == 0 ? 16 : (MODE == 1 ? 32 : 64);
                const uint64_t db = make_desc(smem_u32(sB), N * 16, 128);
                for (int r = 0; r < reps; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) mma_f16_ts(dbase + (q & 1) * 64, abase + 8 * q, db, make_idesc_f16(N), r > 0 ? 1u : 0u);
            } else {
                const uint64_t db2 = make_desc(smem_u32(sB), 32 * 16, 128), db1 = make_desc(smem_u32(sB) + 4096, 64 * 16, 128);
                for (int r = 0; r < reps; ++r) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        mma_f16_ts(dbase + 64, abase + 16 * (q >> 1) * 2 + 8 * (q & 1), db2 + q * 64, make_idesc_f16(32), q > 0 ? 1u : 0u);
                        mma_f16_ts(dbase + 64, abase + 16 * (q >> 1) * 2 + 8 * (q & 1) + 16, db2 + q * 64, make_idesc_f16(16), 1u);
                    }
                    if (MODE == 3) { tc_commit(&bar[4 + warp]); tc_commit(&bar[6 + warp]); }
                    mma_f16_ts(dbase, abase + 56, db1, make_idesc_f16(64), 0u);
                    if (MODE == 3) tc_commit(&bar[2 + warp]);
                }
            }
            t1 = clock64();
            tc_commit(&bar[warp]);
        }
        __syncwarp();
        mbar_wait(&bar[warp], 0);
        if ((tid & 31) == 0) { cycles[2 * warp] = clock64() - t0; }
        if (t1 != t0) cycles[2 * warp + 1] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int MODE>
void run_f16(const char* name, int per_rep, long long* dC) {
    for (int nw : {1, 2}) {
        const int reps = 1024;
        f16_kernel<MODE><<<1, 128>>>(reps, nw, dC);
        CK(cudaDeviceSynchronize());
        long long c[4];
        CK(cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost));
        printf("[f16] %s, issuing warps=%d: %.1f cycles per %s until complete, %.1f to issue (per warp)\n", name, nw,
               (double)c[0] / reps / (per_rep ? 1 : 4), per_rep ? "chunk pattern" : "MMA", (double)c[1] / reps / (per_rep ? 1 : 4));
    }
}

template <int N, int NACC, int TS>
void run_timing(long long* dC) {
    for (int nw : {1, 2}) {
        const int reps = 4096 / NACC;
        timing_kernel<N, NACC, TS><<<1, 128>>>(reps, nw, dC);
        CK(cudaDeviceSynchronize());
        long long c[2];
        CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
        printf("[timing2] %s N=%d accumulators=%d issuing warps=%d: %.2f cycles/MMA per warp (%.2f aggregate)\n",
               TS ? "A=TMEM" : "A=SMEM", N, NACC, nw, (double)c[0] / (reps * NACC), (double)c[0] / (reps * NACC) / nw);
    }
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}
static float tf32_rn(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    srand(1);
    float *dA, *dB0, *dB1, *dD;
    long long* dC;
    CK(cudaMalloc(&dA, 128 * 8 * 4));
    CK(cudaMalloc(&dB0, MAXN * 8 * 4));
    CK(cudaMalloc(&dB1, MAXN * 8 * 4));
    CK(cudaMalloc(&dD, 128 * MAXN * 4));
    CK(cudaMalloc(&dC, 64));
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (int N : {16, 32, 64}) {
        std::vector<float> A(128 * 8), B0(N * 8), B1(N * 8), D(128 * N);
        for (auto& v : A) v = rnd();
        for (auto& v : B0) v = rnd();
        for (auto& v : B1) v = rnd();
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB0, B0.data(), B0.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice));
        probe_kernel<<<1, 128>>>(dA, dB0, dB1, N, 1, dD, dC);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double e_exact = 0, e_trunc = 0, e_rn = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0, st = 0, sr = 0;
                for (int k = 0; k < 8; ++k) {
                    s += (double)A[m * 8 + k] * B0[n * 8 + k] + (double)A[m * 8 + k] * B1[n * 8 + k];
                    st += (double)tf32_trunc(A[m * 8 + k]) * tf32_trunc(B0[n * 8 + k]) +
                          (double)tf32_trunc(A[m * 8 + k]) * tf32_trunc(B1[n * 8 + k]);
                    sr += (double)tf32_rn(A[m * 8 + k]) * tf32_rn(B0[n * 8 + k]) +
                          (double)tf32_rn(A[m * 8 + k]) * tf32_rn(B1[n * 8 + k]);
                }
                e_exact = fmax(e_exact, fabs(D[m * N + n] - s));
                e_trunc = fmax(e_trunc, fabs(D[m * N + n] - st));
                e_rn = fmax(e_rn, fabs(D[m * N + n] - sr));
            }
        printf("[layout] N=%d  max|D - exact| = %.3e   max|D - tf32-truncated operands| = %.3e   max|D - tf32-RN operands| = %.3e\n",
               N, e_exact, e_trunc, e_rn);
    }
    // accumulator rounding: sum of 2 * reps products per element with operands exactly representable in tf32
    {
        const int N = 16, reps = 1024;
        std::vector<float> A(128 * 8), B0(N * 8), B1(N * 8), D(128 * N);
        for (auto& v : A) v = tf32_trunc(fabsf(rnd()) + 0.25f);
        for (auto& v : B0) v = tf32_trunc(fabsf(rnd()) + 0.25f);
        for (auto& v : B1) v = tf32_trunc(fabsf(rnd()) + 0.25f);
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB0, B0.data(), B0.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice));
        probe_kernel<<<1, 128>>>(dA, dB0, dB1, N, reps, dD, dC);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double mean_rel = 0, max_rel = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < 8; ++k) s += (double)A[m * 8 + k] * B0[n * 8 + k] + (double)A[m * 8 + k] * B1[n * 8 + k];
                s *= reps;
                const double rel = (D[m * N + n] - s) / s;
                mean_rel += rel;
                max_rel = fmax(max_rel, fabs(rel));
            }
        mean_rel /= 128.0 * N;
        printf("[accum] %d accumulating MMAs (K=8) of positive terms: mean rel err %.3e (negative mean => truncating adds), max |rel| %.3e; fp32 eps 6e-8\n",
               2 * reps, mean_rel, max_rel);
    }
    for (int N : {16, 32, 64}) {
        const int reps = 2000;
        probe_kernel<<<1, 128>>>(dA, dB0, dB1, N, reps, dD, dC);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
        printf("[timing] N=%d: %d MMAs (128xNx8 tf32, A in TMEM) in %lld cycles = %.2f cycles/MMA\n", N, 2 * reps, c,
               (double)c / (2 * reps));
    }

    run_f16<0>("kind::f16 N=16 K=16 A=TMEM", 0, dC);
    run_f16<1>("kind::f16 N=32 K=16 A=TMEM", 0, dC);
    run_f16<2>("kind::f16 N=64 K=16 A=TMEM", 0, dC);
    run_f16<3>("chunk pattern (8 L2 MMAs, 2 commits, 1 L1 MMA, commit)", 1, dC);
    run_f16<4>("chunk pattern without the commits", 1, dC);
    run_timing<16, 1, 1>(dC);
    run_timing<16, 2, 1>(dC);
    run_timing<16, 4, 1>(dC);
    run_timing<32, 1, 1>(dC);
    run_timing<32, 2, 1>(dC);
    run_timing<64, 1, 1>(dC);
    run_timing<16, 1, 0>(dC);
    run_timing<16, 4, 0>(dC);
    run_timing<32, 2, 0>(dC);
    return 0;
}
