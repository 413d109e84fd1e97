// mma.sync issue-rate probe for sm_100a: TF32 m16n8k8, BF16 m16n8k16, FP16 m16n8k16 (FP32 accumulate).
// Prints cycles per MMA per warp (dependent chain = latency; 4 independent chains = throughput) and
// the MAC rate per SM at a given number of warps.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE, int CHAINS>
__global__ void __launch_bounds__(1024, 1) probe(float *out, long long *cyc, int iters) {
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = (float)(threadIdx.x + i + j);
    unsigned a0 = threadIdx.x * 2654435761u | 0x3f000000u, a1 = a0 ^ 0x1234u, a2 = a0 ^ 0x4321u, a3 = a0 ^ 0x1111u;
    unsigned b0 = a0 ^ 0x2222u, b1 = a0 ^ 0x3333u;
    a0 &= 0x3fffffffu; a1 &= 0x3fffffffu; a2 &= 0x3fffffffu; a3 &= 0x3fffffffu; b0 &= 0x3fffffffu; b1 &= 0x3fffffffu;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE, int CHAINS>
int run(const char *name, int sms, float *out, long long *cyc, double macs) {
    const int iters = 4000;
    for (int threads : {32, 128, 256, 768}) {
        for (int rep = 0; rep < 2; ++rep) {
            probe<MODE, CHAINS><<<sms, threads>>>(out, cyc, iters);
            CK(cudaDeviceSynchronize());
        }
        long long h = 0;
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double warps = threads / 32.0, n = (double)iters * CHAINS;
        printf("%-22s chains %d warps %2.0f  cycles/MMA/warp %6.2f  MMA/clk/SM %.3f  MAC/clk/SM %.0f\n", name, CHAINS, warps,
               (double)h / n, n * warps / (double)h, n * warps * macs / (double)h);
    }
    return 0;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *out;
    long long *cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 1024));
    CK(cudaMalloc(&cyc, 8));
    run<0, 1>("tf32 m16n8k8", sms, out, cyc, 16 * 8 * 8);
    run<0, 4>("tf32 m16n8k8", sms, out, cyc, 16 * 8 * 8);
    run<1, 1>("bf16 m16n8k16", sms, out, cyc, 16 * 8 * 16);
    run<1, 4>("bf16 m16n8k16", sms, out, cyc, 16 * 8 * 16);
    run<2, 4>("f16 m16n8k16", sms, out, cyc, 16 * 8 * 16);
    return 0;
}
