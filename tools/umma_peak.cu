// tcgen05 (UMMA) TF32 peak on this pool's B200: the roofline denominator of the search prefilter tile.
// Every SM runs one CTA that issues back-to-back kind::tf32 MMAs (M = 128, N = 256 or 128, K = 8) on operands that stay
// resident in shared memory (SWIZZLE_128B K-major planes, contents irrelevant), accumulating into TMEM; no loads, no
// epilogue -- the tensor pipe's issue rate and nothing else.  Prints TFLOP/s for cta_group::1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_peak tools/umma_peak.cu && /tmp/umma_peak
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
            return 1;                                                               \
        }                                                                           \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;               // LBO (unused with 128-byte swizzle, K-major)
    d |= (uint64_t)(1024 >> 4) << 32;     // SBO: 8-row atoms 1024 bytes apart
    d |= (uint64_t)1 << 46;               // version 1 (Blackwell)
    d |= (uint64_t)2 << 61;               // 128-byte swizzle
    return d;
}

template <int N>
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int iters, unsigned long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];   // A: 128 x 32 tf32 (16 KB), B: N x 32 tf32
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + N) * 32; i += 128) reinterpret_cast<float *>(smem)[i] = 1.0f / (float)(1 + (i & 15));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 128 * 128;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // 4 K-steps of 8 tf32 (32 bytes) inside the 128-byte swizzled row
                const uint64_t da = make_desc_sw128(a0 + j * 32), db = make_desc_sw128(b0 + j * 32);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_c),
                    "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(it | j)), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0u)
                : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "r"(N));
}

template <int N>
static int run(int sms, double clock_ghz) {
    const int iters = 20000;
    const int smem = (128 + N) * 128 + 1024;
    unsigned long long *dcyc;
    CK(cudaMalloc(&dcyc, 8));
    CK(cudaFuncSetAttribute(umma_peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_peak_kernel<N><<<sms, 128, smem>>>(100, dcyc);   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        umma_peak_kernel<N><<<sms, 128, smem>>>(iters, dcyc);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    unsigned long long cyc = 0;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const double flop = 2.0 * 128 * N * 8 * 4.0 * iters * sms;
    printf("{\"kind\": \"tf32\", \"cta_group\": 1, \"M\": 128, \"N\": %d, \"K\": 8, \"sms\": %d, \"ms\": %.3f, \"tflops\": %.1f, "
           "\"cycles_per_mma\": %.1f}\n",
           N, sms, best, flop / (best * 1e-3) / 1e12, (double)cyc / (4.0 * iters));
    (void)clock_ghz;
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    if (run<256>(sms, prop.clockRate * 1e-6)) return 1;
    if (run<128>(sms, prop.clockRate * 1e-6)) return 1;
    return 0;
}
