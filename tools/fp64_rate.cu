// Issue-rate micro-benchmark for the back-end instruction mix: dependent-free DFMA / DADD / DSETP / FFMA / FMNMX / LOP3
// streams, 1..8 warps per SM sub-partition.  Prints warp-instructions per clock per SM and the implied chip rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_rate tools/fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void rate_kernel(int iters, double seed, double* sink, long long* cycles) {
    constexpr int CH = 8;
    double a[CH];
    float b[CH];
    unsigned u[CH];
    double b2[CH], c2[CH];
    for (int i = 0; i < CH; ++i) { a[i] = seed + threadIdx.x + i; b[i] = (float)a[i]; u[i] = threadIdx.x * 7 + i;
                                   b2[i] = 1.0 + 1e-12 * (threadIdx.x + i) * seed; c2[i] = 1e-9 * (threadIdx.x + 3 * i) * seed; }
    const double m = 1.0 + seed * 1e-12, c = seed * 1e-9;
    const float mf = (float)m, cf = (float)c;
    int cnt = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) a[i] = fma(a[i], m, c);
            if (OP == 1) a[i] = a[i] + c;
            if (OP == 2) cnt += (a[i] > (double)(it + i)) ? 1 : 0, a[i] += 0.0;  // DSETP (+ int add); keep a live
            if (OP == 3) b[i] = fmaf(b[i], mf, cf);
            if (OP == 4) b[i] = fmaxf(b[i], cf + i);
            if (OP == 5) u[i] = (u[i] & 0xFFFFE000u) ^ (unsigned)it;
            if (OP == 6) a[i] = a[i] * m;
            if (OP == 7) a[i] = fma(a[i], b2[i], c2[i]);                                   // three distinct register operands
            if (OP == 8) { a[i] = fma(a[i], b2[i], c2[i]); u[i] = (u[i] & 0xFFFFE000u) ^ (unsigned)it; }   // + one LOP3 each
            if (OP == 9) { a[i] = fma(a[i], m, c); u[i] = (u[i] & 0xFFFFE000u) ^ (unsigned)it; }
            if (OP == 10) { a[i] = fma(a[i], m, c); if (i == 0) a[0] += (double)(int)(u[0] + it); }          // one I2F.F64 per 8 DFMA
            if (OP == 11) a[0] = fma(a[0], m, c);                                          // one dependent chain: latency
            if (OP == 12) a[i & 1] = fma(a[i & 1], m, c);                                  // two chains
            if (OP == 13) a[i & 3] = fma(a[i & 3], m, c);                                  // four chains
        }
    }
    const long long t1 = clock64();
    double s = cnt;
    for (int i = 0; i < CH; ++i) s += a[i] + b[i] + u[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps_per_sm, int sms) {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000, CH = 8;
    rate_kernel<OP><<<sms, warps_per_sm * 32>>>(100, 1.0, sink, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<OP><<<sms, warps_per_sm * 32>>>(iters, 1.0, sink, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double winst = (double)iters * CH * warps_per_sm;  // per SM
    printf("%-11s warps/SM %2d: %.3f warp-instr/clk/SM (%.1f clk per warp-instr per SMSP), %.2f T thread-op/s chip\n", name,
           warps_per_sm, winst / c, c / (winst / 4.0), winst * 32.0 * sms / (ms * 1e-3) / 1e12);
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {4, 8, 16, 32}) {
        run<0>("DFMA", w, sms); run<1>("DADD", w, sms); run<6>("DMUL", w, sms); run<2>("DSETP", w, sms);
        run<3>("FFMA", w, sms); run<4>("FMNMX", w, sms); run<5>("LOP3", w, sms);
    }
    for (int w : {4, 8, 16}) {
        run<7>("DFMA3r", w, sms); run<8>("DFMA3r+LOP", w, sms); run<9>("DFMA+LOP", w, sms); run<10>("DFMA+I2F/8", w, sms);
        run<11>("DFMA chain1", w, sms); run<12>("DFMA chain2", w, sms); run<13>("DFMA chain4", w, sms);
    }
    return 0;
}
