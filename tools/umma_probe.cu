// tcgen05 (UMMA) bring-up probe for the round-2 move of the search prefilter tile from mma.sync to the 5th-generation
// tensor cores: C(128 x 128, f32 in TMEM) = A(128 x K) * B(128 x K)^T with kind::tf32, operands written to shared memory
// by plain stores in the two K-major canonical layouts (no swizzle / 128-byte swizzle), one CTA, no TMA.
//   * terms = 1:  A and B are truncated to TF32 first -> products are exact, the result must match a double
//                 reference to FP32-accumulation accuracy.  A wrong descriptor shows as garbage, not as small error.
//   * terms = 3:  3xTF32 (hi*lo + lo*hi + hi*hi into one accumulator) on full FP32 inputs -> prints the worst
//                 |error| / sum|a||b|, the number the certified bound of search_pf.cuh needs for this data path.
// Compile-checked with nvcc 12.9 for sm_100a; NOT yet run (written after the round's GPU budget was spent).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe
// Descriptor fields follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / InstrDescriptor), the PTX strings
// cute/arch/mma_sm100_umma.hpp, tmem_allocator_sm100.hpp, copy_sm100.hpp and cutlass/arch/barrier.h.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
            return 1;                                                               \
        }                                                                           \
    } while (0)

constexpr int M = 128, N = 128, KB = 32;      // tile; KB tf32 = 128 bytes per row per K block
constexpr int PLANE = 128 * KB * 4;           // bytes of one operand plane of one K block (16 KB)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, column k < 32) inside an operand plane
template <int SW128>
__device__ __forceinline__ int plane_offset(int r, int k) {
    if (SW128)   // rows of 128 bytes, 8-row atoms of 1024 bytes, 16-byte chunks XOR-swizzled with (row % 8)
        return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4;
    // no swizzle: core matrix = 8 rows x 16 bytes contiguous; row groups 128 bytes apart (SBO), K chunks 2048 (LBO)
    return (k >> 2) * 2048 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4;
}

template <int SW128>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint64_t lbo = SW128 ? 1 : (2048 >> 4);
    const uint64_t sbo = SW128 ? (1024 >> 4) : (128 >> 4);
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fff);
    d |= lbo << 16;
    d |= sbo << 32;
    d |= (uint64_t)1 << 46;                      // version = 1 (Blackwell)
    d |= (uint64_t)(SW128 ? 2 : 0) << 61;        // layout type: 0 none, 2 = 128-byte swizzle
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

// terms == 5: BF16x3 -- hi = bf16(v), lo = bf16(v - hi), kind::f16 with K = 16 per instruction, 128-byte swizzled rows of 64
// BF16 values (the planes of search_umma.cuh with search_umma_bf16 = 1)
__global__ void __launch_bounds__(128, 1) umma_probe_bf16_kernel(const float *__restrict__ A, const float *__restrict__ B, int K,
                                                                 float *__restrict__ C) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_hi = smem, *a_lo = smem + PLANE, *b_hi = smem + 2 * PLANE, *b_lo = smem + 3 * PLANE;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int KB16 = 64;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = tmem_base_s;
    // D = F32, A = B = BF16 (format 1), both K-major
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int nkb = K / KB16;
    uint32_t parity = 0, first = 1;
    for (int kb = 0; kb < nkb; ++kb) {
        for (int idx = tid; idx < 128 * KB16; idx += 128) {
            const int r = idx / KB16, k = idx % KB16;
            const float va = A[(size_t)r * K + kb * KB16 + k], vb = B[(size_t)r * K + kb * KB16 + k];
            const __nv_bfloat16 ah = __float2bfloat16_rn(va), bh = __float2bfloat16_rn(vb);
            const __nv_bfloat16 al = __float2bfloat16_rn(va - __bfloat162float(ah)), bl = __float2bfloat16_rn(vb - __bfloat162float(bh));
            const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2;
            *reinterpret_cast<__nv_bfloat16 *>(a_hi + off) = ah;
            *reinterpret_cast<__nv_bfloat16 *>(a_lo + off) = al;
            *reinterpret_cast<__nv_bfloat16 *>(b_hi + off) = bh;
            *reinterpret_cast<__nv_bfloat16 *>(b_lo + off) = bl;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int j = 0; j < KB16 / 16; ++j) {   // one instruction covers K = 16 bf16 = 32 bytes
                const uint32_t step = j * 32;
                const uint64_t dah = make_desc<1>(smem_u32(a_hi) + step), dal = make_desc<1>(smem_u32(a_lo) + step);
                const uint64_t dbh = make_desc<1>(smem_u32(b_hi) + step), dbl = make_desc<1>(smem_u32(b_lo) + step);
                umma_bf16(tmem_c, dah, dbl, idesc, first ? 0u : 1u);
                umma_bf16(tmem_c, dal, dbh, idesc, 1u);
                umma_bf16(tmem_c, dah, dbh, idesc, 1u);
                first = 0;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        {
            uint32_t done = 0;
            while (!done)
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}\n"
                    : "=r"(done)
                    : "r"(smem_u32(&mbar)), "r"(parity)
                    : "memory");
            parity ^= 1;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t addr = tmem_c + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "r"(128));
}

template <int SW128>
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float *__restrict__ A, const float *__restrict__ B, int K,
                                                            int terms, float *__restrict__ C) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_hi = smem, *a_lo = smem + PLANE, *b_hi = smem + 2 * PLANE, *b_lo = smem + 3 * PLANE;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = tmem_base_s;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int nkb = K / KB;
    uint32_t parity = 0, first = 1;
    for (int kb = 0; kb < nkb; ++kb) {
        for (int idx = tid; idx < 128 * KB; idx += 128) {
            const int r = idx / KB, k = idx % KB;
            const float va = A[(size_t)r * K + kb * KB + k], vb = B[(size_t)r * K + kb * KB + k];
            const float ah = __uint_as_float(__float_as_uint(va) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(vb) & 0xffffe000u);
            const int off = plane_offset<SW128>(r, k);
            // terms == 4: the hi planes hold the RAW fp32 value -- if the tensor core ignores the 13 low mantissa bits of a
            // tf32 operand (truncation), the result is bit-identical to terms == 3 and the hi planes need no preparation
            *reinterpret_cast<float *>(a_hi + off) = terms == 4 ? va : ah;
            *reinterpret_cast<float *>(a_lo + off) = va - ah;
            *reinterpret_cast<float *>(b_hi + off) = terms == 4 ? vb : bh;
            *reinterpret_cast<float *>(b_lo + off) = vb - bh;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int j = 0; j < KB / 8; ++j) {                          // one instruction covers K = 8 tf32 = 32 bytes
                const uint32_t step = SW128 ? j * 32 : j * 4096;
                const uint64_t dah = make_desc<SW128>(smem_u32(a_hi) + step), dal = make_desc<SW128>(smem_u32(a_lo) + step);
                const uint64_t dbh = make_desc<SW128>(smem_u32(b_hi) + step), dbl = make_desc<SW128>(smem_u32(b_lo) + step);
                if (terms >= 3) {
                    umma_tf32(tmem_c, dah, dbl, idesc, first ? 0u : 1u);
                    umma_tf32(tmem_c, dal, dbh, idesc, 1u);
                    umma_tf32(tmem_c, dah, dbh, idesc, 1u);
                } else {
                    umma_tf32(tmem_c, dah, dbh, idesc, first ? 0u : 1u);
                }
                first = 0;
            }
            // arrives on the mbarrier when every MMA issued so far has read its operands and written TMEM
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        {   // everyone waits before the planes are overwritten (or the accumulator is read)
            uint32_t done = 0;
            while (!done)
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}\n"
                    : "=r"(done)
                    : "r"(smem_u32(&mbar)), "r"(parity)
                    : "memory");
            parity ^= 1;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // accumulator row m lives in TMEM lane m, column n; warp w may touch lanes 32 w .. 32 w + 31
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t addr = tmem_c + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "r"(128));
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    const int K = 384;
    std::vector<float> A((size_t)M * K), B((size_t)N * K), C((size_t)M * N);
    srand(7);
    for (auto &v : A) v = (float)rand() / RAND_MAX;
    for (auto &v : B) v = (float)rand() / RAND_MAX;
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dC, C.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 4 * PLANE + 1024;
    CK(cudaFuncSetAttribute(umma_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(umma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> C3;
    for (int sw = 0; sw < 2; ++sw)
        for (int terms = 1; terms <= 4; terms += (terms == 3 ? 1 : 2)) {
            CK(cudaMemset(dC, 0xff, C.size() * 4));
            if (sw) umma_probe_kernel<1><<<1, 128, smem>>>(dA, dB, K, terms, dC);
            else umma_probe_kernel<0><<<1, 128, smem>>>(dA, dB, K, terms, dC);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
            double worst = 0.0, worst_rel = 0.0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0.0, mag = 0.0;
                    for (int k = 0; k < K; ++k) {
                        const double a = terms == 1 ? tf32_trunc(A[(size_t)m * K + k]) : A[(size_t)m * K + k];
                        const double b = terms == 1 ? tf32_trunc(B[(size_t)n * K + k]) : B[(size_t)n * K + k];
                        ref += a * b;
                        mag += fabs(a * b);
                    }
                    const double err = fabs((double)C[(size_t)m * N + n] - ref);
                    (void)0;
                    if (err > worst) worst = err;
                    if (err / mag > worst_rel) worst_rel = err / mag;
                }
            if (terms == 3) C3 = C;
            if (terms == 4) {
                size_t diff = 0;
                for (size_t i = 0; i < C.size(); ++i) diff += memcmp(&C[i], &C3[i], 4) != 0;
                printf("{\"raw_fp32_hi_planes_vs_truncated\": \"%zu of %zu accumulators differ\"}\n", diff, C.size());
            }
            printf("{\"layout\": \"%s\", \"terms\": %d, \"K\": %d, \"max_abs_err\": %.3e, \"max_err_over_sum_abs\": %.3e, \"c00\": %.6f}\n",
                   sw ? "K-major SWIZZLE_128B" : "K-major SWIZZLE_NONE", terms, K, worst, worst_rel, C[0]);
        }
    {   // BF16x3 on kind::f16
        CK(cudaFuncSetAttribute(umma_probe_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaMemset(dC, 0xff, C.size() * 4));
        umma_probe_bf16_kernel<<<1, 128, smem>>>(dA, dB, K, dC);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0.0, worst_rel = 0.0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0.0, mag = 0.0;
                for (int k = 0; k < K; ++k) {
                    const double a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
                    ref += a * b;
                    mag += fabs(a * b);
                }
                const double err = fabs((double)C[(size_t)m * N + n] - ref);
                if (err > worst) worst = err;
                if (err / mag > worst_rel) worst_rel = err / mag;
            }
        printf("{\"layout\": \"K-major SWIZZLE_128B\", \"terms\": \"bf16x3 (kind::f16, K=16)\", \"K\": %d, \"max_abs_err\": %.3e, "
               "\"max_err_over_sum_abs\": %.3e, \"bound_over_sum_abs\": %.3e, \"c00\": %.6f}\n",
               K, worst, worst_rel, 3.0 * 3.814697265625e-6 + (17.0 * (3.0 * K / 16.0) + 16.0) * 1.1920928955078125e-7, C[0]);
    }
    return 0;
}
