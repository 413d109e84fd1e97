// Dependent-issue latency of the FP64 pipe on this GPU: one warp, a chain of N dependent DFMA / DADD / DMUL, clock64
// around it.  The chain kernel of the clustering replay (cluster_replay.cu) is ONE dependent sequence per centroid --
// sub, mul, 4 FMA, add per row -- so this number times 7 is the floor of a chain step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_latency tools/dp_latency.cu && ./dp_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, long long *cycles, double a, double b, int n) {
    double x = a;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
        if (OP == 0) x = __fma_rn(x, b, a);
        if (OP == 1) x = __dadd_rn(x, b);
        if (OP == 2) x = __dmul_rn(x, b);
        if (OP == 3) x = __fma_rn(__fma_rn(-(x * b), a, x), b, x * b);   // one Markstein correction round: mul, fma, fma
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
        *out = x;
        *cycles = t1 - t0;
    }
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, 8);
    const int n = 1 << 16;
    const char *names[4] = {"DFMA", "DADD", "DMUL", "mul+fma+fma"};
    for (int op = 0; op < 4; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) chain<0><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n);
            if (op == 1) chain<1><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n);
            if (op == 2) chain<2><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n);
            if (op == 3) chain<3><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("{\"op\": \"%s\", \"dependent_cycles_per_step\": %.2f}\n", names[op], (double)h / n);
    }
    return 0;
}
