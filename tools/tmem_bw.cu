// TMEM port throughput: back-to-back tcgen05.ld / tcgen05.st (32x32b.x32 = 4 KB per warp instruction) from 4 or 8 warps
// of one CTA (warp w reaches lanes 32 (w % 4) .. +31).  Prints bytes per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD32(addr, v) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]), \
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(addr) : "memory")
#define ST32(addr, v) asm volatile( \
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
    :: "r"(addr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]), \
      "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory")

// mode 0: loads, 1: stores, 2: one load + two stores per rep (the activation-warp mix)
__global__ void __launch_bounds__(256) bw_kernel(int mode, int reps, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 * (warp >> 2));
    uint32_t v[32], acc = 0;
    for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
    ST32(tb, v); ST32(tb + 32, v); ST32(tb + 64, v); ST32(tb + 96, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const uint32_t col = 32 * (r & 3);
        if (mode == 0 || mode == 2) {
            LD32(tb + col, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] ^ v[31];
        }
        if (mode == 1 || mode == 2) {
            ST32(tb + 128 + col, v);
            if (mode == 2) ST32(tb + 128 + ((col + 32) & 127), v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "r"(512u) : "memory");
}

int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
    const char* names[3] = {"ld.x32", "st.x32", "ld + 2 st"};
    const int bytes_per_rep[3] = {4096, 4096, 3 * 4096};
    for (int warps : {4, 8})
        for (int mode = 0; mode < 3; ++mode) {
            const int reps = 4000;
            bw_kernel<<<1, warps * 32>>>(mode, reps, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-10s %d warps: %.1f cycles per rep per warp, %.1f B/clk/SM\n", names[mode], warps, (double)c / reps,
                   (double)warps * reps * bytes_per_rep[mode] / c);
        }
    return 0;
}
