// FP32 issue-rate probe for sm_100a (cycles from clock64 inside the kernel, so clock changes do not matter):
// FFMA with three distinct registers, FFMA with a repeated source, FADD+FFMA, packed fma.rn.f32x2.
// Prints warp-instructions per clock per SM (4 = one per scheduler per clock) and FMA lanes/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(float *out, long long *cyc, int iters, float a0, float b0) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = (float)(threadIdx.x + i);
    float a[4] = {a0, a0 + 1.f, a0 + 2.f, a0 + 3.f};
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = b0 + (float)i;
    unsigned long long acc2[16], x2[4], a2[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc2[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(x2[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
#pragma unroll
    for (int i = 0; i < 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(a2[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // dot tile: acc[c*8+i] += x[i] * a[c]   (3 distinct registers)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c * 8 + i] = fmaf(x[i], a[c], acc[c * 8 + i]);
        } else if (MODE == 1) {  // acc += x * x   (2 distinct registers)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(x[i & 7], x[i & 7], acc[i]);
        } else if (MODE == 2) {  // d = x - a; acc += d * d
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float d = x[i] - a[c];
                    acc[c * 8 + i] = fmaf(d, d, acc[c * 8 + i]);
                }
        } else if (MODE == 3) {  // packed: 16 x f32x2 accumulators, operands stay packed
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(x2[i & 3]), "l"(a2[(i >> 2) & 1]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *out;
    long long *cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 1024));
    CK(cudaMalloc(&cyc, 8));
    const int iters = 20000;
    const char *names[4] = {"FFMA 3 regs (dot tile)", "FFMA x*x+acc (2 regs)", "FADD + FFMA d*d", "fma.rn.f32x2 (FFMA2)"};
    const double fma_per_iter[4] = {32, 32, 16, 32};
    const double instr_per_iter[4] = {32, 32, 32, 16};
    for (int threads : {128, 256, 512, 768, 1024}) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<sms, threads>>>(out, cyc, iters, 1.0f, 2.0f);
                if (mode == 1) probe<1><<<sms, threads>>>(out, cyc, iters, 1.0f, 2.0f);
                if (mode == 2) probe<2><<<sms, threads>>>(out, cyc, iters, 1.0f, 2.0f);
                if (mode == 3) probe<3><<<sms, threads>>>(out, cyc, iters, 1.0f, 2.0f);
                CK(cudaDeviceSynchronize());
            }
            long long h = 0;
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            const double warps = threads / 32.0;
            printf("threads %4d  %-26s cycles %9lld  FP warp-instr/clk/SM %.2f   FMA lanes/clk/SM %.1f\n", threads,
                   names[mode], h, instr_per_iter[mode] * iters * warps / (double)h,
                   fma_per_iter[mode] * iters * warps * 32 / (double)h);
        }
    }
    return 0;
}
