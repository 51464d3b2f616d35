// Issue-rate micro-benchmark for the activation warps' ReLU + fp16 hi/lo split (tc_kernel.cuh: relu_split_f16x2):
// F2FP.RELU..RZ / F2FP.RELU (cvt.{rz,rn}.relu.f16x2.f32), FHFMA (fma.rn.f32.f16), the 1:2:1 mix the kernel issues, and FADD /
// LOP3 / FFMA for comparison; 1..8 warps per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/alu_rate tools/alu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t cvt_rz(float lo, float hi) { uint32_t r; asm volatile("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t cvt_rn(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ void rem2(uint32_t hp, float v0, float v1, float& l0, float& l1) {
    asm volatile("{\n\t.reg .b16 a, b, m;\n\tmov.b32 {a, b}, %2;\n\tmov.b16 m, 0xBC00;\n\t"
                 "fma.rn.f32.f16 %0, a, m, %3;\n\tfma.rn.f32.f16 %1, b, m, %4;\n\t}" : "=f"(l0), "=f"(l1) : "r"(hp), "f"(v0), "f"(v1));
}

template <int OP>
__global__ void rate_kernel(int iters, float seed, float* sink, long long* cycles) {
    constexpr int CH = 8;
    float v[2 * CH];
    uint32_t h[CH], l[CH];
    for (int i = 0; i < 2 * CH; ++i) v[i] = seed * (threadIdx.x + i) - 3.f;
    for (int i = 0; i < CH; ++i) { h[i] = threadIdx.x + i; l[i] = 0; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) h[i] = cvt_rz(v[2 * i], v[2 * i + 1]);                                   // 1 F2FP.RZ
            if (OP == 1) h[i] = cvt_rn(v[2 * i], v[2 * i + 1]);                                   // 1 F2FP
            if (OP == 2) rem2(h[i], v[2 * i], v[2 * i + 1], v[2 * i], v[2 * i + 1]);               // 2 FHFMA
            if (OP == 3) {                                                                        // the kernel's mix: 4 instr
                float l0, l1;
                h[i] = cvt_rz(v[2 * i], v[2 * i + 1]);
                rem2(h[i], v[2 * i], v[2 * i + 1], l0, l1);
                l[i] = cvt_rn(l0, l1);
                v[2 * i] += __uint_as_float(l[i] & 0x3f800000u) ; // keep a dependence without much extra work (LOP + FADD)
            }
            if (OP == 4) { v[2 * i] = v[2 * i] + seed; v[2 * i + 1] = v[2 * i + 1] + seed; }       // 2 FADD
            if (OP == 5) { h[i] = (h[i] & 0xFFFFE000u) ^ (uint32_t)it; }                           // 1 LOP3
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CH; ++i) s += v[2 * i] + v[2 * i + 1] + __uint_as_float(h[i]) + __uint_as_float(l[i]);
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, int warps_per_sm, int sms) {
    float* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000, CH = 8;
    rate_kernel<OP><<<sms, warps_per_sm * 32>>>(100, 1.0f, sink, cyc);
    rate_kernel<OP><<<sms, warps_per_sm * 32>>>(iters, 1.0f, sink, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double winst = (double)iters * CH * per_iter * warps_per_sm;  // per SM
    printf("%-22s warps/SM %2d: %.3f warp-instr/clk/SM = %.2f clk per warp-instr per SMSP\n", name, warps_per_sm, winst / c, c / (winst / 4.0));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {4, 8, 16, 32}) {
        run<0>("F2FP.RELU.RZ", 1, w, sms); run<1>("F2FP.RELU.RN", 1, w, sms); run<2>("FHFMA", 2, w, sms);
        run<3>("split mix (4+2 instr)", 6, w, sms); run<4>("FADD", 2, w, sms); run<5>("LOP3", 1, w, sms);
    }
    return 0;
}
