// search_umma.cuh -- the certified prefilter tile of search_pf.cuh on Blackwell's own tensor pipe (included by search.cu).
//
// Same scheme, same certificate, same candidate lists and the same finishing kernel as search_pf.cuh (read its header
// first); only the contraction cos~ = <q^, x^> moves from mma.sync + cp.async to
//   * tcgen05.mma kind::tf32 (SASS UTCHMMA...), cta_group::1, M = 128 queries x N = 256 items x K = 8 per instruction,
//     3xTF32 as three instructions per K-step into ONE accumulator: hi*lo + lo*hi + hi*hi;
//   * operands staged by TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) from four planes prepared once per call -- q^ / x^
//     split into hi = the TF32 truncation (exactly representable, so the tensor core's own operand conversion has
//     nothing to drop) and lo = fl32(x^) - hi -- as K-major 128-byte-swizzled boxes of 32 features: 96 KB per stage
//     (2 x 16 KB query boxes, 2 x 32 KB item boxes), UM_STAGES stages on full / empty mbarriers;
//   * accumulators in tensor memory: 2 x 256 columns, so the tile t + 1 is multiplied while the epilogue warps read tile
//     t back with tcgen05.ld (SASS LDTM): warp-specialised -- one TMA thread, one MMA thread, four epilogue warps that
//     each own 32 TMEM lanes = 32 queries, ONE THREAD PER QUERY: the blended score, the slab's k best (a sorted list
//     per query in shared memory) and the emission test are thread-local, no shuffles, no block barriers.
// Error model: the bound E of search_pf.cuh assumes exact products and a truncating aligned accumulation of at most
// 9 * 2^-23 per instruction; tools/umma_probe.cu measured this data path at K = 384 (144 instructions into one
// accumulator): worst |error| 9.7e-6 of sum|a||b| against the bound's 1.56e-4 (profiles/r02_umma_probe.log) -- the
// same 0.06 the mma.sync path shows, so E is kept unchanged.
#pragma once

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

namespace {

constexpr int UM_TQ = 128, UM_TN = 256;
constexpr int UM_THREADS = 384;                                    // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4..7 / 8..11 epilogue
constexpr int UM_EPI_GROUPS = 2;                                   // epilogue group g drains the tiles t = g (mod 2)
constexpr int UM_RING_BYTES = 192 * 1024;                          // operand ring: UM_RING_BYTES / stage bytes stages
constexpr int UM_MAX_STAGES = 4;
// features per stage KC = 32 (128-byte rows, SWIZZLE_128B, 96 KB stages, 2 in the ring) or 16 (64-byte rows, SWIZZLE_64B,
// 48 KB stages, 4 in the ring: three loads in flight while one stage is multiplied -- a 96 KB stage alone cannot keep
// 64 B/clk of operand traffic in flight across the L2 latency, profiles/r02_umma_v1_summary.csv: tensor pipe 21 %)
template <int KC>
struct UmCfg {
    static constexpr int A_BYTES = UM_TQ * KC * 4;                 // one query plane box
    static constexpr int B_BYTES = UM_TN * KC * 4;                 // one item plane box
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // hi + lo of both operands
    static constexpr int STAGES = UM_RING_BYTES / STAGE_BYTES;
    static constexpr int ROW_BYTES = KC * 4;
    static constexpr unsigned long long LAYOUT = KC == 32 ? 2ull : 4ull;   // SWIZZLE_128B : SWIZZLE_64B
};

struct UmBarriers {
    unsigned long long full[UM_MAX_STAGES], empty[UM_MAX_STAGES], tfull[2], tempty[2];
    unsigned tmem_base;
};

__device__ __forceinline__ unsigned um_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void um_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void um_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(um_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void um_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(um_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void um_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "UM_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra UM_DONE;\n\t"
        "bra UM_WAIT_LOOP;\n\t"
        "UM_DONE:\n\t}" ::"r"(um_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void um_tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            um_smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(um_smem_u32(bar))
        : "memory");
}
// K-major operand plane, rows of ROW_BYTES (one swizzle span), 8-row atoms 8 ROW_BYTES apart -- the layout TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B / _64B (canonical layouts of cute/atom/mma_traits_sm100.hpp: Swizzle<3,4,3> / <2,4,3> o
// ((8,n),2):((8 | 4,SBO),1) in 16-byte units; descriptor fields as cute/arch/mma_sm100_desc.hpp; tools/umma_probe.cu)
template <int KC>
__device__ __forceinline__ unsigned long long um_desc(unsigned saddr) {
    unsigned long long d = (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= 1ull << 16;                                                   // leading byte offset (unused for these layouts)
    d |= (unsigned long long)((8 * UmCfg<KC>::ROW_BYTES) >> 4) << 32;  // stride byte offset: one 8-row atom
    d |= 1ull << 46;                                                   // descriptor version (Blackwell)
    d |= UmCfg<KC>::LAYOUT << 61;
    return d;
}
__device__ __forceinline__ void um_mma_tf32(unsigned tmem_c, unsigned long long da, unsigned long long db, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// kind::f16 with BF16 operands: K = 16 per instruction (32 bytes per row, the same descriptor step as TF32's K = 8)
__device__ __forceinline__ void um_mma_bf16(unsigned tmem_c, unsigned long long da, unsigned long long db, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// the same box delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA in mask
__device__ __forceinline__ void um_tma_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar,
                                             unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], "
        "[%4], %5;" ::"r"(um_smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(um_smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void um_commit_mc(unsigned long long *bar, unsigned short mask) {   // ... in every CTA of mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     um_smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ unsigned um_cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void um_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void um_commit(unsigned long long *bar) {   // arrives when every MMA issued so far is done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem_u32(bar)) : "memory");
}

// One warp per row: hi = TF32 truncation of fl32(row / |row|), lo = fl32(row / |row|) - hi (exact), zero padded to fp.
__global__ void __launch_bounds__(256) um_split_rows_kernel(const double *__restrict__ rows, const double *__restrict__ norms2,
                                                            long long n, int f, int fp, float *__restrict__ hi,
                                                            float *__restrict__ lo, int *__restrict__ flags,
                                                            double *__restrict__ norm_out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double n2 = norms2[r];
    const bool ok = (n2 == 0.0) || (n2 >= 1e-290 && n2 <= 1e290);   // false for NaN / inf as well
    if (!ok && lane == 0) atomicOr(flags, PF_FLAG_FALLBACK);
    const double inv = (ok && n2 > 0.0) ? 1.0 / sqrt(n2) : 0.0;
    if (norm_out && lane == 0) norm_out[r] = ok ? sqrt(n2) : 0.0;
    const double *src = rows + r * (long long)f;
    for (int j = lane; j < fp; j += 32) {
        const float v = (ok && j < f) ? (float)(src[j] * inv) : 0.0f;
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[r * (long long)fp + j] = h;
        lo[r * (long long)fp + j] = v - h;
    }
}

// BF16x3 planes: hi = bf16(v) (round to nearest), lo = bf16(v - hi): |v - hi - lo| <= 2^-18 |v|.  Half the operand bytes
// and twice the tensor rate of the TF32 planes for the same certificate (search_pf.cuh: the accumulation term dominates
// E_cos, not the representation term).
__global__ void __launch_bounds__(256) um_split_rows_bf16_kernel(const double *__restrict__ rows, const double *__restrict__ norms2,
                                                                 long long n, int f, int fp, __nv_bfloat16 *__restrict__ hi,
                                                                 __nv_bfloat16 *__restrict__ lo, int *__restrict__ flags,
                                                                 double *__restrict__ norm_out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double n2 = norms2[r];
    const bool ok = (n2 == 0.0) || (n2 >= 1e-290 && n2 <= 1e290);   // false for NaN / inf as well
    if (!ok && lane == 0) atomicOr(flags, PF_FLAG_FALLBACK);
    const double inv = (ok && n2 > 0.0) ? 1.0 / sqrt(n2) : 0.0;
    if (norm_out && lane == 0) norm_out[r] = ok ? sqrt(n2) : 0.0;
    const double *src = rows + r * (long long)f;
    for (int j = 2 * lane; j < fp; j += 64) {   // two features per lane: 4-byte stores
        float v[2];
        __nv_bfloat16 h[2], l[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            v[u] = (ok && j + u < f) ? (float)(src[j + u] * inv) : 0.0f;
            h[u] = __float2bfloat16_rn(v[u]);
            l[u] = __float2bfloat16_rn(v[u] - __bfloat162float(h[u]));
        }
        *reinterpret_cast<__nv_bfloat162 *>(hi + r * (long long)fp + j) = __nv_bfloat162(h[0], h[1]);
        *reinterpret_cast<__nv_bfloat162 *>(lo + r * (long long)fp + j) = __nv_bfloat162(l[0], l[1]);
    }
}

struct UmMaps {
    CUtensorMap qhi, qlo, xhi, xlo;
};

// shared memory: [UM_STAGES][A_hi | A_lo | B_hi | B_lo] (1024-byte aligned), then per accumulator buffer the item-side
// epilogue inputs (2 x 256 doubles each), then the per-query sorted lists (float, rounded down: still lower bounds)
// CL = CTAs per cluster (1, 2 or 4; consecutive query tiles of one slab): they walk the same item tiles in step, so
// every item box is fetched from L2 ONCE per cluster and multicast into all CL shared memories -- rank c fetches piece c
// of the 2 planes x 256 rows (CL = 2: one plane each; CL = 4: half a plane each).  A stage may only be refilled when
// every CTA of the cluster has multiplied it: the MMA threads commit their "stage free" signal to all CL empty barriers.
// BF: BF16x3 planes (kind::f16, K = 16 per instruction): a stage row of 4 UM_KC bytes then holds 2 UM_KC features.
template <int MODE, int UM_KC, int CL, bool BF>
__global__ void __launch_bounds__(UM_THREADS, 1) search_umma_kernel(const __grid_constant__ UmMaps maps, PfArgs A) {
    constexpr int FPC = BF ? 2 * UM_KC : UM_KC;   // features per pipeline stage
    constexpr int UM_A_BYTES = UmCfg<UM_KC>::A_BYTES, UM_B_BYTES = UmCfg<UM_KC>::B_BYTES;
    constexpr int UM_STAGE_BYTES = UmCfg<UM_KC>::STAGE_BYTES, UM_STAGES = UmCfg<UM_KC>::STAGES;
    extern __shared__ __align__(1024) unsigned char um_smem[];
    __shared__ UmBarriers bars;
    // the swizzled boxes need 1024-byte alignment; the launch reserves the slack (um_smem_bytes)
    unsigned char *stages = um_smem + ((1024u - (um_smem_u32(um_smem) & 1023u)) & 1023u);
    double *ep_x0 = reinterpret_cast<double *>(stages + (size_t)UM_STAGES * UM_STAGE_BYTES);   // [2][UM_TN]: lambda / |x|^2
    double *ep_x1 = ep_x0 + 2 * UM_TN;                                                           // [2][UM_TN]: |x| (PF_L2 / PF_NEAR)
    float *ep_f0 = reinterpret_cast<float *>(ep_x1 + (MODE != PF_COSINE ? 2 * UM_TN : 0));       // FP32 copies of both
    float *ep_f1 = ep_f0 + 2 * UM_TN;
    float *lists = ep_f1 + (MODE != PF_COSINE ? 2 * UM_TN : 0);                                  // [UM_TQ][k]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = A.k, fp = A.fp;
    const long long q0 = (long long)blockIdx.x * UM_TQ;
    const long long ntiles_total = (A.n + UM_TN - 1) / UM_TN;
    const long long t_begin = (long long)blockIdx.y * A.tiles_per_slab;
    long long t_end = t_begin + A.tiles_per_slab;
    if (t_end > ntiles_total) t_end = ntiles_total;
    const int ntile = (int)(t_end > t_begin ? t_end - t_begin : 0);
    const int nchunks = fp / FPC;
    if (ntile == 0) return;

    const unsigned crank = CL > 1 ? um_cluster_ctarank() : 0u;
    constexpr unsigned short kAllCtas = (unsigned short)((1u << CL) - 1u);
    if (tid == 0) {
        for (int s = 0; s < UM_STAGES; ++s) {
            um_mbar_init(&bars.full[s], 1);
            um_mbar_init(&bars.empty[s], CL);
        }
        for (int b = 0; b < 2; ++b) {
            um_mbar_init(&bars.tfull[b], 1);
            um_mbar_init(&bars.tempty[b], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(&bars.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) um_cluster_sync();   // every CTA's barriers exist before a peer multicasts to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = bars.tmem_base;

    if (warp == 0) {
        // ---- TMA producer
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < ntile; ++t) {
                const int row_x = (int)((t_begin + t) * UM_TN);
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % UM_STAGES;
                    if (it >= UM_STAGES) um_mbar_wait(&bars.empty[s], ((it / UM_STAGES) - 1) & 1);
                    unsigned char *st = stages + (size_t)s * UM_STAGE_BYTES;
                    um_mbar_expect_tx(&bars.full[s], UM_STAGE_BYTES);
                    um_tma_2d(st, &maps.qhi, c * FPC, (int)q0, &bars.full[s]);
                    um_tma_2d(st + UM_A_BYTES, &maps.qlo, c * FPC, (int)q0, &bars.full[s]);
                    if constexpr (CL == 1) {
                        um_tma_2d(st + 2 * UM_A_BYTES, &maps.xhi, c * FPC, row_x, &bars.full[s]);
                        um_tma_2d(st + 2 * UM_A_BYTES + UM_B_BYTES, &maps.xlo, c * FPC, row_x, &bars.full[s]);
                    } else if constexpr (CL == 2) {   // rank 0: the hi plane, rank 1: the lo plane
                        um_tma_2d_mc(st + 2 * UM_A_BYTES + crank * UM_B_BYTES, crank == 0 ? &maps.xhi : &maps.xlo, c * FPC,
                                     row_x, &bars.full[s], kAllCtas);
                    } else {                          // rank 0 / 1: halves of the hi plane, rank 2 / 3: of the lo plane
                        um_tma_2d_mc(st + 2 * UM_A_BYTES + (crank >> 1) * UM_B_BYTES + (crank & 1) * (UM_B_BYTES / 2),
                                     (crank >> 1) == 0 ? &maps.xhi : &maps.xlo, c * FPC, row_x + (int)(crank & 1) * (UM_TN / 2),
                                     &bars.full[s], kAllCtas);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: D[128 x 256] (+)= A[128 x 8] B[256 x 8]^T, three instructions per K-step
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            // (formats: TF32 = 2 with kind::tf32, BF16 = 1 with kind::f16)
            const unsigned fmt = BF ? 1u : 2u;
            const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(UM_TN >> 3) << 17) | ((unsigned)(UM_TQ >> 4) << 24);
            int it = 0;
            for (int t = 0; t < ntile; ++t) {
                const int b = t & 1;
                if (t >= 2) um_mbar_wait(&bars.tempty[b], ((t >> 1) - 1) & 1);   // the epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tmem_c = tmem_base + (unsigned)(b * UM_TN);
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % UM_STAGES;
                    um_mbar_wait(&bars.full[s], (it / UM_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned sa = um_smem_u32(stages + (size_t)s * UM_STAGE_BYTES);
#pragma unroll
                    for (int j = 0; j < UM_KC / 8; ++j) {
                        const unsigned long long dah = um_desc<UM_KC>(sa + 32 * j), dal = um_desc<UM_KC>(sa + UM_A_BYTES + 32 * j);
                        const unsigned long long dbh = um_desc<UM_KC>(sa + 2 * UM_A_BYTES + 32 * j);
                        const unsigned long long dbl = um_desc<UM_KC>(sa + 2 * UM_A_BYTES + UM_B_BYTES + 32 * j);
                        if constexpr (BF) {
                            um_mma_bf16(tmem_c, dah, dbl, idesc, (c | j) ? 1u : 0u);
                            um_mma_bf16(tmem_c, dal, dbh, idesc, 1u);
                            um_mma_bf16(tmem_c, dah, dbh, idesc, 1u);
                        } else {
                            um_mma_tf32(tmem_c, dah, dbl, idesc, (c | j) ? 1u : 0u);
                            um_mma_tf32(tmem_c, dal, dbh, idesc, 1u);
                            um_mma_tf32(tmem_c, dah, dbh, idesc, 1u);
                        }
                    }
                    // the stage is free once these instructions have read it -- in every CTA that feeds it
                    if constexpr (CL == 1) um_commit(&bars.empty[s]);
                    else um_commit_mc(&bars.empty[s], kAllCtas);
                }
                um_commit(&bars.tfull[b]);       // the accumulator is complete
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: thread e = query q0 + e = TMEM lane e.  TWO groups of four warps: group g owns accumulator buffer g,
        // i.e. the tiles t = g (mod 2), with its own sorted list and bounds per query (like two interleaved slabs; the
        // groups meet in gthr like slabs do).  ncu of the one-group version (profiles/r02_umma_bf16_*): tensor pipe 45 %
        // active, the four epilogue warps -- one per scheduler, nothing to hide a tcgen05.ld or a dependent FP32 chain
        // behind -- took two tiles' worth of MMA time per tile.
        const int eg = (warp - 4) >> 2, ew = (warp - 4) & 3;   // group, TMEM lane quarter (= warp % 4)
        const int e = ew * 32 + lane;
        const int ngroups = A.epi_groups;   // 1 when the second group's lists do not fit shared memory (k > 20)
        if (eg < ngroups) {
        const long long gq = q0 + e;
        const bool okq = gq < A.nq;
        double lq = 1.0, qn = 0.0, qband = A.band;
        long long self = -1;
        if constexpr (MODE == PF_COSINE) {
            lq = okq ? A.lambda_q[gq] : 1.0;
            if (okq && blockIdx.y == 0 && lq == 0.0) atomicOr(A.status, STATUS_ZERO_LAMBDA);   // core.rs:773-776
        } else {
            const double xmax2 = __longlong_as_double((long long)*A.xn2max_bits);
            lq = okq ? A.qn2[gq] : 0.0;
            qn = okq ? A.qnrm[gq] : 0.0;
            qband = A.band_rel * qn * sqrt(xmax2) + A.band_abs * (lq + xmax2);
            self = (okq && A.self_idx) ? A.self_idx[gq] : -1ll;
        }
        float *lst = lists + ((size_t)eg * UM_TQ + e) * k;
        int len = 0;
        double kth = -INFINITY;
        bool saw_nan = false;
        float near1 = -INFINITY, near2 = -INFINITY;   // PF_NEAR: largest and second largest FP32 score, column of the largest
        int near_i = -1;
        const double beta = 1.0 - A.alpha;
        // The hot loop is FP32: sf approximates the FP64 score s~ of search_pf.cuh to within fdelta (operands rounded to
        // FP32, five FP32 operations on magnitudes bounded by smag), and only an element with sf >= thr_f = (bound - band
        // - fdelta) rounded down -- or a NaN -- reaches the FP64 code below, which evaluates s~ exactly as the mma.sync
        // kernel does and applies the SAME tests: sf < thr_f implies s~ < bound - band, so the lists, the published bounds
        // and the emitted candidates are unchanged.  (The FP64 epilogue cost 4x the tile's MMAs: profiles/r02_umma_v1_*.)
        const float lqf = (float)lq, qnf2 = (float)(2.0 * qn), alf = (float)A.alpha, bef = (float)beta;
        // |sf - s~| <= fdelta, u = 2^-24 < 6e-8.  Cosine: the lambda term is clamped at |lq - lx| >= 1, so only |lx| <=
        // |lq| + 1.1 matter: u (2 |alpha| + |beta| (3 (|lq| + 1.1) + 4)).  L2: u (3 (q2 + x2) + 8 |q| |x|) <= 7 u (q2 + max x2).
        // A non-finite fdelta (absurd lambda / norm) makes thr_f -inf: every element takes the exact path.
        double fdelta;
        if constexpr (MODE == PF_COSINE) fdelta = 1.0e-7 * (2.0 * fabs(A.alpha) + fabs(beta) * (3.0 * fabs(lq) + 8.0)) + 1e-30;
        else fdelta = 1.0e-6 * (lq + __longlong_as_double((long long)*A.xn2max_bits)) + 1e-300;
        if (!(fdelta < 1e25)) fdelta = INFINITY;   // beyond FP32's range (or NaN): no pre-test
        for (int t = eg; t < ntile; t += ngroups) {
            const int b = t & 1;   // (= eg with two groups)
            const long long i0 = (t_begin + t) * UM_TN;
            // item-side inputs of this tile (two items per thread) go to the group's buffer: every thread of the group
            // must be done reading the previous tile's values first
            if (t >= ngroups) {
                if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                else asm volatile("bar.sync 2, 128;" ::: "memory");
            }
            for (int c = e; c < UM_TN; c += 128) {
                const long long gi = i0 + c;
                if constexpr (MODE == PF_COSINE) {
                    const double v0 = gi < A.n ? A.lambdas[gi] : 0.0;
                    ep_x0[b * UM_TN + c] = v0;
                    ep_f0[b * UM_TN + c] = (float)v0;
                } else {
                    const double v0 = gi < A.n ? A.xn2[gi] : 0.0, v1 = gi < A.n ? A.xnrm[gi] : 0.0;
                    ep_x0[b * UM_TN + c] = v0;
                    ep_x1[b * UM_TN + c] = v1;
                    ep_f0[b * UM_TN + c] = (MODE == PF_NEAR && gi >= A.n) ? INFINITY : (float)v0;   // padding never wins
                    ep_f1[b * UM_TN + c] = (float)v1;
                }
            }
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            double gb = -INFINITY;                                             // bound published by any slab
            if constexpr (MODE != PF_NEAR) gb = okq ? pf_dec(__ldcg(&A.gthr[gq])) : -INFINITY;
            um_mbar_wait(&bars.tfull[b], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long nvalid_ll = A.n - i0;   // columns below this index are real items
            const int nvalid = nvalid_ll < UM_TN ? (int)nvalid_ll : UM_TN;
            const long long self_rel_ll = self - i0;
            const int self_rel = (self_rel_ll >= 0 && self_rel_ll < UM_TN) ? (int)self_rel_ll : -1;
            bool changed = false;
            double bound = fmax(kth, gb);
            float thr_f = okq ? __double2float_rd(bound - qband - fdelta) : INFINITY;   // -inf while the list is filling
            const float *f0 = ep_f0 + b * UM_TN, *f1 = ep_f1 + b * UM_TN;
#pragma unroll 1
            for (int c0 = 0; c0 < UM_TN; c0 += 32) {
                unsigned v[32];
                const unsigned taddr = tmem_base + ((unsigned)(ew * 32) << 16) + (unsigned)(b * UM_TN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (!okq || c0 >= nvalid) continue;
                if constexpr (MODE == PF_NEAR) {
                    // the two largest FP32 scores -|q - x|^2 and the column of the largest, branch-free
#pragma unroll
                    for (int j4 = 0; j4 < 32; j4 += 4) {
                        const float4 a0 = *reinterpret_cast<const float4 *>(f0 + c0 + j4);
                        const float4 a1 = *reinterpret_cast<const float4 *>(f1 + c0 + j4);
                        const float x0[4] = {a0.x, a0.y, a0.z, a0.w}, x1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float sf = fmaf(qnf2 * x1[u], __uint_as_float(v[j4 + u]), -(lqf + x0[u]));
                            near2 = fmaxf(near2, fminf(near1, sf));
                            near_i = sf > near1 ? (int)(i0 + c0 + j4 + u) : near_i;
                            near1 = fmaxf(near1, sf);
                        }
                    }
                    continue;
                }
                // branch-free over the 32 columns: a bit per column whose FP32 score reaches the threshold
                unsigned hot = 0u;
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(f0 + c0 + j4);
                    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if constexpr (MODE == PF_L2) a1 = *reinterpret_cast<const float4 *>(f1 + c0 + j4);
                    const float x0[4] = {a0.x, a0.y, a0.z, a0.w}, x1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float cosf = __uint_as_float(v[j4 + u]);
                        float sf;
                        if constexpr (MODE == PF_COSINE) sf = fmaf(alf, cosf, bef * (1.0f - fminf(fabsf(lqf - x0[u]), 1.0f)));
                        else sf = fmaf(qnf2 * x1[u], cosf, -(lqf + x0[u]));
                        hot |= (sf < thr_f) ? 0u : (1u << (j4 + u));   // (a NaN sets its bit)
                    }
                }
                while (hot) {
                    const int j = __ffs(hot) - 1;
                    hot &= hot - 1;
                    unsigned vj = 0u;
#pragma unroll
                    for (int t = 0; t < 32; ++t) vj = (j == t) ? v[t] : vj;   // v stays in registers
                    {   // the threshold may have risen since the bit was set
                        const float cosf = __uint_as_float(vj);
                        float sf;
                        if constexpr (MODE == PF_COSINE) sf = fmaf(alf, cosf, bef * (1.0f - fminf(fabsf(lqf - f0[c0 + j]), 1.0f)));
                        else sf = fmaf(qnf2 * f1[c0 + j], cosf, -(lqf + f0[c0 + j]));
                        if (sf < thr_f) continue;
                    }
                    // ---- rare: the exact FP64 evaluation and tests of search_pf.cuh
                    const int c = c0 + j;
                    const double cosv = (double)__uint_as_float(vj);
                    double s;
                    bool valid = c < nvalid;
                    if constexpr (MODE == PF_COSINE) {
                        const double lam = 1.0 - fmin(fabs(lq - ep_x0[b * UM_TN + c]), 1.0);   // core.rs:136-137
                        s = A.alpha * cosv + beta * lam;                                       // core.rs:165 (approximate cos)
                    } else {
                        s = -(lq + ep_x0[b * UM_TN + c] - 2.0 * qn * ep_x1[b * UM_TN + c] * cosv);   // -|q - x|^2
                        valid = valid && c != self_rel;
                    }
                    if (valid && s != s) saw_nan = true;
                    if (valid && s >= bound - qband) {
                        if (s > bound) {
                            // keep the slab's k best approximate scores: sorted descending, stored rounded DOWN
                            const float sd = __double2float_rd(s);
                            int pos = len < k ? len : k - 1;
                            while (pos > 0 && lst[pos - 1] < sd) {
                                lst[pos] = lst[pos - 1];
                                --pos;
                            }
                            lst[pos] = sd;
                            if (len < k) ++len;
                            if (len == k) {
                                kth = (double)lst[k - 1];
                                bound = fmax(kth, gb);
                                thr_f = __double2float_rd(bound - qband - fdelta);
                                changed = true;
                            }
                        }
                        // emit what the bound cannot exclude
                        const int p = atomicAdd(&A.cand_cnt[gq], 1);
                        if (p < A.cap) {
                            A.cand_idx[(size_t)gq * A.cap + p] = (int)(i0 + c);
                            A.cand_s[(size_t)gq * A.cap + p] = (float)s;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            um_mbar_arrive(&bars.tempty[b]);
            if (changed && okq) atomicMax(&A.gthr[gq], pf_enc(kth));
        }
        if (saw_nan) atomicOr(A.flags, PF_FLAG_FALLBACK);
        if constexpr (MODE == PF_NEAR) {
            // the two groups of a query merge their (largest, second largest, column) through the list area (unused here)
            float *mrg = lists;   // [UM_TQ][3]
            if (eg == 1) {
                mrg[3 * e] = near1;
                mrg[3 * e + 1] = near2;
                mrg[3 * e + 2] = __int_as_float(near_i);
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (eg == 0) {
                const float o1 = mrg[3 * e], o2 = mrg[3 * e + 1];
                const int oi = __float_as_int(mrg[3 * e + 2]);
                const float hi1 = fmaxf(near1, o1);
                near2 = fmaxf(fminf(near1, o1), fmaxf(near2, o2));
                // equal scores: the lower column, as one group walking all tiles in order would have kept
                near_i = (o1 > near1 || (o1 == near1 && oi >= 0 && (near_i < 0 || oi < near_i))) ? oi : near_i;
                near1 = hi1;
            }
            if (okq && eg == 0) {
                // true d^2 = |q|^2 + |x|^2 - 2 |q| |x| cos: the tile's cos~ is within e_cos of cos (search_pf.cuh), the FP32
                // evaluation of the score within fdelta of the FP64 one, the norms are FP64 sums (1e-13 covers them)
                const double xmax2 = __longlong_as_double((long long)*A.xn2max_bits);
                const double etot = 2.0 * qn * sqrt(xmax2) * A.e_cos * (1.0 + 1e-6) + fdelta + 1e-13 * (lq + xmax2);
                double dlo = sqrt(fmax(-(double)near1 - etot, 0.0)) * (1.0 - 1e-15);
                double dhi = sqrt(fmax(-(double)near1 + etot, 0.0)) * (1.0 + 1e-15);
                double slo = near2 > -INFINITY ? sqrt(fmax(-(double)near2 - etot, 0.0)) * (1.0 - 1e-15) : 0.0;
                if (near_i < 0 || !(dhi == dhi) || !(dlo == dlo) || !(slo == slo) || !(etot < 1e300)) {
                    dlo = 0.0;      // no usable bounds: the chain takes exact steps, the certification fails the row
                    dhi = 1e300;
                    slo = 0.0;
                }
                A.near_idx[gq] = near_i;
                A.near_b[3 * gq] = dlo;
                A.near_b[3 * gq + 1] = dhi;
                A.near_b[3 * gq + 2] = slo;
            }
        }
        }   // eg < ngroups
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) um_cluster_sync();   // no CTA leaves while a peer may still write its shared memory or barriers
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

size_t um_smem_bytes(int k, int mode, int groups = 1) {
    const size_t lists = (size_t)groups * UM_TQ * k * 4;
    return (size_t)UM_RING_BYTES + (size_t)(mode != PF_COSINE ? 4 : 2) * UM_TN * 12 + (lists > 3 * UM_TQ * 4 ? lists : 3 * UM_TQ * 4) + 1024;
}

typedef CUresult (*um_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

um_encode_fn um_encoder() {
    static um_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<um_encode_fn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// rows x fp fp32 plane, boxes of kc features x box_rows rows, swizzle span = box row, out-of-range rows read as zeros
// (bf: the plane holds BF16 values, a box row of 4 kc bytes = 2 kc features)
bool um_make_map(CUtensorMap *map, const float *plane, long long rows, int fp, int box_rows, int kc, bool bf) {
    um_encode_fn enc = um_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)fp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)fp * (bf ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)(bf ? 2 * kc : kc), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(plane), dims,
               strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// option "search_umma" (default 1): 0 keeps the mma.sync tile of search_pf.cuh
bool um_wanted(asb_ctx *ctx) {
    auto it = ctx->options.find("search_umma");
    return it == ctx->options.end() || it->second != 0.0;
}
// option "search_umma_bf16" (default 1): BF16x3 planes on kind::f16 instead of 3xTF32 (0)
bool um_bf16(asb_ctx *ctx) {
    auto it = ctx->options.find("search_umma_bf16");
    return it == ctx->options.end() || it->second != 0.0;
}
// option "search_umma_kc": 32-bit words per stage row, 16 (default: 64-byte rows) or 32 (TF32 planes only)
int um_kc(asb_ctx *ctx) {
    auto it = ctx->options.find("search_umma_kc");
    return (it != ctx->options.end() && it->second == 32.0 && !um_bf16(ctx)) ? 32 : 16;
}
// the cosine error bound of the tile (search_pf.cuh, "E_cos"): FP32 rounding of the unit rows, the split's residual and
// dropped lo * lo term, and the accumulation chain -- 3 fp / K instructions into one FP32 accumulator, each within
// (K + 1) 2^-23 of exact (K products and the accumulator through truncating aligned adders: measured by
// tools/umma_probe.cu for both operand types) -- times 1.1.  The BF16 bound is never below the TF32 one.
double um_e_cos(int fp, bool bf) {
    const double u23 = 1.1920928955078125e-7;
    const double tf = 2.0 * u23 + 3.0 * 9.5367431640625e-7 + (9.0 * (3.0 * fp / 8.0) + 16.0) * u23;
    const double b16 = 2.0 * u23 + 3.0 * 3.814697265625e-6 + (17.0 * (3.0 * fp / 16.0) + 16.0) * u23;
    return 1.1 * (bf ? (b16 > tf ? b16 : tf) : tf);
}

// option "search_umma_cluster": CTAs that share one multicast item stream, 1, 2 (default) or 4
int um_cluster(asb_ctx *ctx, long long nq) {
    auto it = ctx->options.find("search_umma_cluster");
    int cl = it == ctx->options.end() ? 2 : (int)it->second;
    if (cl != 1 && cl != 2 && cl != 4) cl = 2;
    while (cl > 1 && (nq + UM_TQ - 1) / UM_TQ < cl) cl >>= 1;   // fewer query tiles than CTAs per cluster
    return cl;
}

bool um_make_maps(asb_ctx *ctx, UmMaps *maps, const float *qhi, const float *qlo, long long nq, const float *xhi,
                  const float *xlo, long long n, int fp) {
    const int kc = um_kc(ctx);
    const bool bf = um_bf16(ctx);
    const int xbox = um_cluster(ctx, nq) == 4 ? UM_TN / 2 : UM_TN;   // rows per item box (a CTA of 4 fetches half a plane)
    return um_make_map(&maps->qhi, qhi, nq, fp, UM_TQ, kc, bf) && um_make_map(&maps->qlo, qlo, nq, fp, UM_TQ, kc, bf) &&
           um_make_map(&maps->xhi, xhi, n, fp, xbox, kc, bf) && um_make_map(&maps->xlo, xlo, n, fp, xbox, kc, bf);
}

// Slabs of the tcgen05 tile: the CTAs of one slab (all query tiles) stream the same item tiles but start at different
// times, so they only share those tiles through L2 if a whole slab's planes fit there: with 601-tile slabs the 3 GB of
// planes were fetched from HBM 14 times over (profiles/r02_launches_c3.md: 42.6 GB per launch).  Slabs are capped at
// "search_umma_slab_mb" (default 80 MB of BF16 hi + lo planes, 48 MB of TF32 planes: under the 126 MB L2), wave balance
// permitting.
void um_pick_slabs(asb_ctx *ctx, long long qtiles, long long ntiles, int fp, int *nslabs, long long *tps) {
    // (measured at C3: 24 / 48 / 80 MB of BF16 planes per slab -> tile 19.2 / 19.1 / 18.5 ms, 1 624 / 1 309 / 1 095 candidates
    // emitted per query: longer slabs share their bounds sooner and fill the pipeline less often)
    const bool bf = um_bf16(ctx);
    double mb = bf ? 80.0 : 48.0;
    auto it = ctx->options.find("search_umma_slab_mb");
    if (it != ctx->options.end() && it->second > 0.0) mb = it->second;
    const double tile_bytes = (double)UM_TN * fp * (bf ? 4.0 : 8.0);   // hi + lo planes of one item tile
    long long cap = (long long)(mb * 1048576.0 / tile_bytes);
    if (cap < 4) cap = 4;
    long long min_slabs = (ntiles + cap - 1) / cap;
    if (min_slabs < 1) min_slabs = 1;
    // the wave-balancing choice among splits with at least min_slabs slabs (and at most 4x that, 4096 at most)
    long long limit = min_slabs * 4 < 4096 ? min_slabs * 4 : 4096;
    if (limit < 64) limit = 64;
    double best_cost = 0.0;
    long long best_tps = ntiles > 0 ? ntiles : 1;
    int best_ns = 1;
    bool have = false;
    for (long long ns = min_slabs; ns <= limit && ns <= (ntiles > 0 ? ntiles : 1); ++ns) {
        const long long t = (ntiles + ns - 1) / ns;
        const long long real_ns = (ntiles + t - 1) / t;
        const long long waves = (qtiles * real_ns + ctx->sm_count - 1) / ctx->sm_count;
        const double cost = (double)waves * (double)(t + 2) * (1.0 + 5e-4 * (double)real_ns);
        if (!have || cost < best_cost) {
            have = true;
            best_cost = cost;
            best_tps = t;
            best_ns = (int)real_ns;
        }
    }
    *nslabs = best_ns;
    *tps = best_tps;
}

template <int MODE, int KC, int CL, bool BF>
int um_launch_one(asb_ctx *ctx, const UmMaps &maps, const PfArgs &A, int nslabs, size_t usmem) {
    ASB_CUDA(ctx, cudaFuncSetAttribute(search_umma_kernel<MODE, KC, CL, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem));
    const unsigned qtiles = (unsigned)((A.nq + UM_TQ - 1) / UM_TQ);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((qtiles + CL - 1) / CL * CL, (unsigned)nslabs);   // a padding CTA sees no query in range
    cfg.blockDim = dim3(UM_THREADS);
    cfg.dynamicSmemBytes = usmem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ASB_CUDA(ctx, cudaLaunchKernelEx(&cfg, search_umma_kernel<MODE, KC, CL, BF>, maps, A));
    return ASB_OK;
}

template <int MODE>
int um_launch(asb_ctx *ctx, const UmMaps &maps, const PfArgs &A_in, int nslabs, const char *timer) {
    PfArgs A = A_in;
    A.epi_groups = (MODE == PF_NEAR || um_smem_bytes(A.k, MODE, UM_EPI_GROUPS) <= 226 * 1024) ? UM_EPI_GROUPS : 1;
    const size_t usmem = um_smem_bytes(A.k, MODE, A.epi_groups);
    const int kc = um_kc(ctx), cl = um_cluster(ctx, A.nq);
    ctx->kernel_ms["search_umma_cluster"] = (double)cl;
    KernelTimer kt(ctx, timer);
#define ASB_UM_GO(KC, CL) return um_launch_one<MODE, KC, CL, false>(ctx, maps, A, nslabs, usmem)
    if (um_bf16(ctx)) {
        ctx->kernel_ms["search_umma_bf16"] = 1.0;
        if (cl == 4) return um_launch_one<MODE, 16, 4, true>(ctx, maps, A, nslabs, usmem);
        if (cl == 2) return um_launch_one<MODE, 16, 2, true>(ctx, maps, A, nslabs, usmem);
        return um_launch_one<MODE, 16, 1, true>(ctx, maps, A, nslabs, usmem);
    }
    ctx->kernel_ms["search_umma_bf16"] = 0.0;
    if (kc == 32) {
        if (cl == 4) ASB_UM_GO(32, 4);
        if (cl == 2) ASB_UM_GO(32, 2);
        ASB_UM_GO(32, 1);
    }
    if (cl == 4) ASB_UM_GO(16, 4);
    if (cl == 2) ASB_UM_GO(16, 2);
    ASB_UM_GO(16, 1);
#undef ASB_UM_GO
}

}  // namespace
