// cluster_f32p.cuh -- K2, the default variant: the certified 32-row-block walk of cluster_f32.cuh, software-
// pipelined, with the distances on the tensor cores.
//
// cluster_f32_kernel spends more than half of every block outside the distance tile: the cluster barrier of the
// all-to-all, the in-order resolve (one warp, 23 idle) and the tail of the updates
// (profiles/r01_cluster_phases.md).  Here they overlap the distance tile of the NEXT block:
//   * roles (16 warps x 128 registers): warp 0 = control (receive, resolve), warps 1-7 = distances, warps 8-15 =
//     centroid updates, each on its own slice of the features (the update is element-wise);
//   * while warp 0 resolves block b, the compute warps produce q = |c|^2 - 2<x,c> for block b+1 with mma.sync
//     TF32 in 3xTF32 form -- speculatively: block b+1 is assumed to start where block b ends, which holds whenever
//     block b commits whole (virtually always once max_clusters is reached).  The compute warp that finishes the
//     block merges the per-tile arg-mins and sends the CTA's entries to all 16 CTAs with st.async +
//     mbarrier::complete_tx (no cluster barrier; 4 exchange buffers by sequence number);
//   * those distances see centroids that are up to two blocks stale and possibly torn (every element read is
//     some version of that element, so the vector read is within the summed displacement of the outstanding
//     updates of the vector the row will really meet).  The staleness enters the interval arithmetic as one more
//     displacement term P = E(b-1) + E(b-2), the certified displacement bounds of the two preceding blocks, so a
//     decision is still taken from fast distances only when it is CERTIFIED to equal the reference's;
//   * a block that does not commit whole (uncertifiable row, new centroid) drops the speculative tile (its
//     exchange is drained) and the next block is computed from settled centroids (P = 0); a row that cannot be
//     certified goes through the reference-arithmetic FP64 path.
// The FP64 centroids live in global memory (L2-resident; only the update warps and the exact path touch them),
// which frees the shared memory for a 64-row ring of FP32 rows (bulk async copies) and the exchange buffers.
// Outputs are bit-identical to the reference, like every other variant.
#pragma once

namespace {

#define ASB_TICKP(k)                                 \
    do {                                             \
        if (A.phase_times) {                         \
            if (tid == A.tick_tid) {                 \
                const long long _t = clock64();      \
                tphase[k] += _t - tlast;             \
                tlast = _t;                          \
            }                                        \
            __syncwarp(); /* aligned barriers follow */ \
        }                                            \
    } while (0)

constexpr int kXBuf = 4;       // all-to-all buffers: a peer CTA can run up to three exchanges ahead of the slowest reader
constexpr int kPipeGroups = 8;  // ring = up to 8 groups of 8 rows (the block being read + the next one in flight);
                                 // 6 or 4 groups when the centroid shadows need the shared memory (A.ring_groups)

// shared::cta address -> the same location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ unsigned mapa_u32(unsigned saddr, int rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Row pitch (floats) of the FP32 rows / centroid shadows: the feature count padded to the 32-wide k-chunk of
// the MMA tile plus 4 floats, so that the 8 threads of a quarter-warp (2 rows x 4 k-segments) hit 32
// different banks with 16-byte loads.  Padding is zero and contributes nothing to the dot products.
__host__ __device__ inline int f32p_pitch(int f) { return ((f + 31) & ~31) + 4; }

// D(16x8, f32) += A(16x8, tf32, row) * B(8x8, tf32, col).  Operands are FP32 bit patterns (low 13 mantissa bits ignored).
__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 16-byte store into a peer CTA's shared memory that completes 16 tx-bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_16(unsigned raddr, unsigned rmbar, unsigned a, unsigned b, unsigned c,
                                            unsigned d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     raddr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(rmbar)
                 : "memory");
}
// merge two (best, second, best id) triples; ties on the distance go to the lower id
__device__ __forceinline__ void merge_best(float &bd, float &sd, int &bc, float obd, float osd, int obc) {
    if (lex_less_f(obd, obc, bd, bc)) {
        sd = fminf(fminf(sd, osd), bd);
        bd = obd;
        bc = obc;
    } else {
        sd = fminf(fminf(sd, osd), obd);
    }
}
// RN(a / b) from y = RN(1 / b) with two FMA correction steps (Markstein): q0 = RN(a y) is within 2 ulp of a / b,
// q1 = RN(q0 + (a - b q0) y) is faithful, and one more step from a faithful quotient with a correctly rounded
// reciprocal yields the correctly rounded quotient (b is a cluster count, an integer below 2^52, so its significand
// is never all ones -- the one exception of the theorem).  The residuals a - b q are exact in an FMA as long as
// nothing under- or overflows: operands outside [1e-280, 1e280] (and non-finite ones) take the library division.
__device__ __forceinline__ double div_by_count(double a, double b, double y) {
    const double aa = fabs(a);
    if (!(aa >= 1e-280 && aa <= 1e280)) return a == 0.0 ? a / b : __ddiv_rn(a, b);
    const double q0 = __dmul_rn(a, y);
    const double q1 = __fma_rn(__fma_rn(-q0, b, a), y, q0);
    return __fma_rn(__fma_rn(-q1, b, a), y, q1);
}
// x = hi + lo exactly; hi has 10 explicit mantissa bits (a TF32 number), |lo| < 2^-10 |x|
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__global__ void __launch_bounds__(512, 1) cluster_f32p_kernel(ClusterArgs A) {
    constexpr int B = kB32;
    const int NG = A.ring_groups;  // 8, 6 or 4
    cg::cluster_group cluster = cg::this_cluster();
    const int ncta = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int f = A.f;
    const int slots = A.slots_per_cta;
    const int slots4 = (slots + 7) & ~7;  // the MMA tile covers 8 centroids
    const int fp = f32p_pitch(f);
    const int maxk = A.max_k;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);                         // NG * 8 * fp
    float *cent32 = ring + (size_t)NG * kGroup * fp;                           // slots4 * fp
    float *Dall = cent32 + (size_t)slots4 * fp;                                // [2][slots4 * B]
    Xch32 *xch = reinterpret_cast<Xch32 *>(Dall + (size_t)2 * slots4 * B);     // [kXBuf][16][B]
    Xch32 *part_all = xch + kXBuf * 16 * B;                                    // [2][slots4 / 8][B] per-tile arg-min
    Xch *xch_exact = reinterpret_cast<Xch *>(part_all + 2 * (slots4 / 8) * B);  // [2][16]
    GRow32 *G = reinterpret_cast<GRow32 *>(xch_exact + 32);                    // B
    Dec *dec_all = reinterpret_cast<Dec *>(G + B);                             // [2][B]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(dec_all + 2 * B);  // NG mbarriers
    unsigned long long *xbar = full + kPipeGroups;                                      // kXBuf: all-to-all landed (per buffer)
    unsigned long long *cnt = xbar + kXBuf;                                    // maxk (replicated counts)
    double *disp = reinterpret_cast<double *>(cnt + maxk);                     // maxk
    double *wred_d = disp + maxk;                                              // 32
    double *ctld_all = wred_d + 32;                                            // [2]: displacement bound of the block
    double *xrow64 = ctld_all + 2;                                              // f (exact path row)
    int *wred_c = reinterpret_cast<int *>(xrow64 + f);                         // 32
    int *ctl_all = wred_c + 32;                                                // [2][4]
    int *item_ctr = ctl_all + 8;                                               // [2] work counter, [2] items finished
    int *modlist = item_ctr + 4;                                               // B

    const double xmax = sqrt(__longlong_as_double((long long)*A.max_norm2_bits)) * (1.0 + 1e-6);
    const double eta = 3e-7 * xmax;                  // input rounding: 2u(|x| + |c|) <= 2.4e-7 * max|x|
    // Error of q = |c|^2 - 2<x,c> from the tensor-core tile (3xTF32: hi*hi + hi*lo + lo*hi):
    //   * dropped lo*lo term and the TF32 truncation of the lo operands: 3 * 2^-20 |x||c|;
    //   * accumulation: exact products, every mma.sync result within 9 * 2^-23 * max(|acc|, |a_k b_k|) of the
    //     exact sum (aligned, truncating adders of at least 24 bits -- the behaviour reported for NVIDIA tensor
    //     cores since Volta; the `cluster_check_tile` option measures the real error of every tile against
    //     FP64 and tests/test_gpu_parity.py asserts it stays below delta / 4).  The main accumulation is split over 4
    //     independent accumulators, so a chain is nchunk MMAs long, plus 3 FP32 additions and the final FMA;
    //   * |c|^2: 4 chains of 2 * nchunk FMAs + 4 additions.
    // Every centroid is a running mean of rows, so |c| <= max|x|.
    const double nchunk = (double)((f + 31) / 32);
    const double e_dot = 3.0 * 9.5367431640625e-7 + (nchunk * 9.0 + 16.0) * 1.1920928955078125e-7;
    const double e_cn = (2.0 * nchunk + 8.0) * 5.97e-8;
    const double delta = (2.0 * e_dot + e_cn + 5.97e-8 * 3.0) * xmax * xmax * 1.01;

    auto c64 = [&](int slot) -> double * { return A.centroids + ((size_t)slot * ncta + rank) * f; };
    auto c32 = [&](int slot) -> float * { return cent32 + (size_t)slot * fp; };
    auto Dbuf = [&](int which) -> float * { return Dall + (size_t)which * slots4 * B; };
    auto Pbuf = [&](int which) -> Xch32 * { return part_all + (size_t)which * (slots4 / 8) * B; };

    if (tid == 0) {
        for (int s = 0; s < kPipeGroups; ++s) mbar_init(&full[s], 1);
        for (int k = 0; k < kXBuf; ++k) mbar_init(&xbar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        item_ctr[0] = item_ctr[1] = item_ctr[2] = item_ctr[3] = 0;
    }
    for (int c = tid; c < maxk; c += blockDim.x) {
        cnt[c] = (c < A.init_k) ? A.sizes[c] : 0ull;
        disp[c] = 0.0;
    }
    for (int j = tid; j < slots4 * fp; j += blockDim.x) cent32[j] = 0.0f;
    int kc = A.init_k;
    __syncthreads();
    for (int s = warp; s < slots; s += nw) {  // resume: adopt the state left by the previous shard
        const int c = s * ncta + rank;
        if (c >= kc) break;
        const double *src = A.centroids + (size_t)c * f;
        for (int j = lane; j < f; j += 32) c32(s)[j] = (float)src[j];
    }
    const double r_half = A.radius * 0.5, r_full = A.radius, r_relax = A.radius * 1.5;
    const unsigned row_bytes = (unsigned)fp * 4u;
    const long long g_total = (A.n + kGroup - 1) / kGroup;
    long long r0 = 0, n_blocks = 0;
    int n_exact = 0;
    __shared__ long long tphase[8];  // debug phase timers (thread 0 only)
    __shared__ long long arr_t[32], arr_hist[32], arr_late[32];
    if (tid < 32) arr_t[tid] = arr_hist[tid] = arr_late[tid] = 0;
    if (tid == 0)
        for (int k = 0; k < 8; ++k) tphase[k] = 0;
    long long tlast = clock64();

    // Warp roles.  Warp 0 is the control warp (arg-min, all-to-all, resolve): its chain of dependent
    // instructions is the critical path.  The tensor-core tile needs few warps (one per 8 centroids x 16 rows),
    // so the rest apply the centroid updates, each on its own slice of the features.
    // A full block has 2 work items per 8 centroids of this CTA: one compute warp per item where possible
    // (7 of 16 warps for <= 28 centroids per CTA, up to 11 for more), the rest apply updates.
    int ncomp = 2 * ((slots + 7) >> 3);        // warps 1 .. ncomp compute distances in the pipelined steady state
    if (ncomp < (nw - 1) / 2) ncomp = (nw - 1) / 2;
    if (ncomp > nw - 5) ncomp = nw - 5;
    const int first_apply = ncomp + 1;         // warps first_apply .. nw-1 apply updates
    const int napply = nw - first_apply;
    const bool is_compute = warp >= 1 && warp <= ncomp;
    const bool is_apply = warp >= first_apply;
    const int feat_per_apply = (f + napply - 1) / napply;
    // ---- row ring: groups [ring_lo, ring_hi) are resident or in flight.  Every thread tracks, per slot,
    // the parity of the latest copy (issued) and whether it has already observed it landing (waited).
    long long ring_lo = 0, ring_hi = 0;
    unsigned issued = 0u, waited = 0xffffffffu;
    const bool fetcher = (tid == ncomp * 32);  // lane 0 of the last compute warp
    auto need_groups = [&](long long ga, long long gb) {
        if (ga < ring_lo || ga > ring_hi) ring_lo = ring_hi = ga;  // rows were dropped from the ring: restart here
        for (long long g = ring_hi; g < gb; ++g) {
            const int s = (int)((unsigned)g % (unsigned)NG);  // group indices fit 31 bits (n < 2^34 rows)
            if (fetcher) {
                // one copy per slot in flight: the previous one has landed before the slot is reused
                if (!((waited >> s) & 1u)) mbar_wait(&full[s], ((issued >> s) & 1u) ^ 1u);
                const long long r = g * kGroup;
                const long long rows_in = (A.n - r) < kGroup ? (A.n - r) : kGroup;
                const unsigned bytes = (unsigned)rows_in * row_bytes;
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(ring + (size_t)(s * kGroup) * fp, A.rows32 + r * (long long)fp, bytes, &full[s]);
            }
            issued ^= 1u << s;
            waited &= ~(1u << s);
        }
        if (gb > ring_hi) ring_hi = gb;
        if (ring_hi - ring_lo > NG) ring_lo = ring_hi - NG;
        __syncwarp();
    };
    auto wait_group = [&](long long g) {
        const int s = (int)((unsigned)g % (unsigned)NG);
        if (!((waited >> s) & 1u)) {
            mbar_wait(&full[s], ((issued >> s) & 1u) ^ 1u);
            waited |= 1u << s;
        }
    };

    int my_n = 0;
    // (one warp; lane = row) merge the per-tile arg-mins of this CTA, then one 16-byte store per peer CTA and row
    // that also completes 16 tx-bytes on the peer's mbarrier: instruction k carries the 32 rows' entries to peer k
    // (contiguous 512 bytes) -- no staging, no fence, no barrier round trip.  Exchange number `seq` uses buffer
    // seq % kXBuf; every CTA performs the same sequence of exchanges.
    auto send_block = [&](int nbk, const Xch32 *part, int seq) {
        float bd = INFINITY, sd = INFINITY;
        int bc = kNone;
        const int xb = seq & (kXBuf - 1);
        if (lane < nbk) {
            const int ntile = (my_n + 7) >> 3;
            for (int nt = 0; nt < ntile; ++nt) {
                const Xch32 v = part[nt * B + lane];
                merge_best(bd, sd, bc, v.bd, v.sd, v.bc);
            }
            const unsigned dst = smem_u32(xch + ((size_t)xb * 16 + rank) * B + lane);
            const unsigned bar = smem_u32(&xbar[xb]);
            for (int peer = 0; peer < ncta; ++peer)
                st_async_16(mapa_u32(dst, peer), mapa_u32(bar, peer), __float_as_uint(bd), __float_as_uint(sd),
                            (unsigned)bc, 0u);
        }
        __syncwarp();
    };
    // (warp 0) wait until exchange `seq` (16 x nbk entries of 16 bytes) has landed in this CTA
    auto recv_block = [&](int nbk, int seq) {
        const int xb = seq & (kXBuf - 1);
        if (lane == 0) mbar_expect_tx(&xbar[xb], (unsigned)(ncta * nbk * (int)sizeof(Xch32)));
        __syncwarp();
        mbar_wait(&xbar[xb], (unsigned)((seq / kXBuf) & 1));
    };

    // ---- q = |c|^2 - 2 <x, c> for the rows of block [rb, rb + nbk) and this CTA's centroids on the tensor cores
    // (mma.sync m16n8k8 TF32, 3xTF32 split, FP32 accumulate); the squared distance is |x|^2 + q.
    // Work item = (8 centroids) x (2 ring groups = 16 rows) over all features.  Thread (g, t) of the warp loads
    // 8 consecutive features (k0 + 8 t ..) of row g of both groups and of centroid g with 16-byte loads; MMA
    // k-step m pairs the values (2m, 2m+1) of every thread -- A and B use the same feature-to-slot mapping, which
    // is all a dot product needs.  |c|^2 comes from the very registers fed to the MMAs, so a centroid that is
    // being rewritten while it is read still yields the distance to the vector actually read.
    // ctr == nullptr: static round-robin over the warps, else a shared work counter.
    auto phase1 = [&](long long rb, int nbk, float *D, Xch32 *part, int *ctr, int w0, int wn, int seq) {
        const int ntile = (my_n + 7) >> 3;
        const long long g0 = rb / kGroup;
        const int lead = (int)(rb & (kGroup - 1));
        const int nrg = (lead + nbk + kGroup - 1) / kGroup;  // ring groups the block touches (<= 4)
        const int nmt = (nrg + 1) >> 1;
        const int nitems = ntile * nmt;
        const int g = lane >> 2, t = lane & 3;
        int it;
        if (ctr) {
            if (nitems == 0 && warp == w0) send_block(nbk, part, seq);  // this CTA holds no centroid: all-infinite entries
            it = 0;
            if (lane == 0) it = atomicAdd(ctr, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
        } else {
            it = warp - w0;  // static split over warps w0 .. w0 + wn - 1
        }
        while (it < nitems) {
            const int mt = it / ntile, nt = it - mt * ntile;
            const int ga = 2 * mt, gb = (2 * mt + 1 < nrg) ? 2 * mt + 1 : 2 * mt;  // odd tail: group a twice
            wait_group(g0 + ga);
            wait_group(g0 + gb);
            const float *xa = ring + (size_t)(((unsigned)(g0 + ga) % (unsigned)NG) * kGroup + g) * fp + 8 * t;
            const float *xb = ring + (size_t)(((unsigned)(g0 + gb) % (unsigned)NG) * kGroup + g) * fp + 8 * t;
            const float *cb = c32(nt * 8 + g) + 8 * t;
            float cm[4][4], cc[4], cn4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cc[i] = 0.0f;
                cn4[i] = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) cm[i][j] = 0.0f;
            }
            for (int k0 = 0; k0 < fp - 4; k0 += 32) {
                const float4 pa0 = *reinterpret_cast<const float4 *>(xa + k0);
                const float4 pa1 = *reinterpret_cast<const float4 *>(xa + k0 + 4);
                const float4 pb0 = *reinterpret_cast<const float4 *>(xb + k0);
                const float4 pb1 = *reinterpret_cast<const float4 *>(xb + k0 + 4);
                const float4 pc0 = *reinterpret_cast<const float4 *>(cb + k0);
                const float4 pc1 = *reinterpret_cast<const float4 *>(cb + k0 + 4);
                const float va[8] = {pa0.x, pa0.y, pa0.z, pa0.w, pa1.x, pa1.y, pa1.z, pa1.w};
                const float vb[8] = {pb0.x, pb0.y, pb0.z, pb0.w, pb1.x, pb1.y, pb1.z, pb1.w};
                const float vc[8] = {pc0.x, pc0.y, pc0.z, pc0.w, pc1.x, pc1.y, pc1.z, pc1.w};
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    unsigned ah[4], al[4], bh[2], bl[2];
                    split_tf32(va[2 * m], ah[0], al[0]);      // (row g of group a, slot t)
                    split_tf32(vb[2 * m], ah[1], al[1]);      // (row g of group b, slot t)      -> MMA row g + 8
                    split_tf32(va[2 * m + 1], ah[2], al[2]);  // slot t + 4
                    split_tf32(vb[2 * m + 1], ah[3], al[3]);
                    split_tf32(vc[2 * m], bh[0], bl[0]);
                    split_tf32(vc[2 * m + 1], bh[1], bl[1]);
                    cn4[m] = fmaf(vc[2 * m], vc[2 * m], cn4[m]);
                    cn4[m] = fmaf(vc[2 * m + 1], vc[2 * m + 1], cn4[m]);
                    mma_tf32(cm[m], ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
                    mma_tf32(cc, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
                    mma_tf32(cc, al[0], al[1], al[2], al[3], bh[0], bh[1]);
                }
            }
            float cn = (cn4[0] + cn4[1]) + (cn4[2] + cn4[3]);
            cn += __shfl_xor_sync(0xffffffffu, cn, 1);
            cn += __shfl_xor_sync(0xffffffffu, cn, 2);  // |centroid nt*8 + g|^2 in the 4 lanes of group g
            const float cn0 = __shfl_sync(0xffffffffu, cn, 8 * t), cn1 = __shfl_sync(0xffffffffu, cn, 8 * t + 4);
            float qv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // c0:(row g, col 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)
                const float dot = ((cm[0][j] + cm[1][j]) + (cm[2][j] + cm[3][j])) + cc[j];
                float q = fmaf(-2.0f, dot, (j & 1) ? cn1 : cn0);
                const int slot = nt * 8 + 2 * t + (j & 1);
                if (!(fabsf(q) < INFINITY) || slot >= my_n) q = INFINITY;  // non-finite: never wins (exact path decides)
                qv[j] = q;
                const int grp = (j & 2) ? 2 * mt + 1 : 2 * mt;
                const int ri = 8 * grp + g - lead;
                if (slot < my_n && grp < nrg && ri >= 0 && ri < nbk) D[(size_t)slot * B + ri] = q;
                if (A.tile_check && slot < my_n && grp < nrg && ri >= 0 && ri < nbk && q < INFINITY && delta > 0.0) {
                    // debug: the same quantity in FP64 from the very FP32 operands (centroids are settled: no
                    // speculation in this mode); worst |error| / delta over the run, in units of 1e-12
                    const float *xrow = ring + (size_t)(((unsigned)(g0 + grp) % (unsigned)NG) * kGroup + g) * fp;
                    const float *crow = c32(slot);
                    double sc = 0.0, sx = 0.0;
                    for (int kk = 0; kk < f; ++kk) {
                        sc = fma((double)crow[kk], (double)crow[kk], sc);
                        sx = fma((double)xrow[kk], (double)crow[kk], sx);
                    }
                    const double ratio = fabs((double)q - (sc - 2.0 * sx)) / delta;
                    atomicMax(reinterpret_cast<unsigned long long *>(A.phase_times + 47),
                              (unsigned long long)(ratio * 1e12));
                }
            }
            // per row: (best, second, best id) over the 8 centroids of the tile -- 2 in this thread, 4 threads per row
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = (nt * 8 + 2 * t) * ncta + rank;  // centroid id of column 2t; column 2t+1 is c0 + ncta
                float bd = qv[2 * h], sd = qv[2 * h + 1];
                int bc = c0;
                if (sd < bd) {  // strict: the lower id keeps a tie
                    const float tmp = bd;
                    bd = sd;
                    sd = tmp;
                    bc = c0 + ncta;
                }
                if (!(bd < INFINITY)) bc = kNone;
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const float obd = __shfl_xor_sync(0xffffffffu, bd, o);
                    const float osd = __shfl_xor_sync(0xffffffffu, sd, o);
                    const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                    merge_best(bd, sd, bc, obd, osd, obc);
                }
                const int grp = 2 * mt + h;
                const int ri = 8 * grp + g - lead;
                if (t == 0 && grp < nrg && ri >= 0 && ri < nbk) {
                    Xch32 v;
                    v.bd = bd;
                    v.sd = sd;
                    v.bc = bc;
                    v.pad = 0;
                    part[nt * B + ri] = v;
                }
            }
            if (ctr) {
                // the warp that finishes the last item of the block sends the block's entries right away
                __threadfence_block();
                int fin = 0;
                if (lane == 0) {
                    fin = atomicAdd(ctr + 2, 1);
                    it = atomicAdd(ctr, 1);
                }
                fin = __shfl_sync(0xffffffffu, fin, 0);
                it = __shfl_sync(0xffffffffu, it, 0);
                if (fin == nitems - 1) {
                    __threadfence_block();
                    send_block(nbk, part, seq);
                }
            } else {
                it += wn;
            }
        }
    };
    // ---- per row: arg-min over this CTA's centroids (warp 0; lane = row, no shuffles), then ONE 16 x nbk-byte
    // bulk copy per peer CTA that also signals the peer's mbarrier -- 16 packets per block instead of 16 x nbk
    long long tq = 0;
    auto probe = [&](int slot, double dep) {  // debug: accumulate cycles since the previous probe of warp 0
        if (A.phase_times && warp == 0) {
            long long t;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep) : "memory");
            if (lane == 0 && rank == 0 && slot >= 0) A.phase_times[slot] += t - tq;
            tq = t;
        }
    };
    __syncthreads();
    cluster.sync();

    bool have = false;     // the distances + all-to-all of the block at r0 are already in flight (speculated)
    int cur = 0;           // distance buffer / exchange parity of the block at r0
    double nx_ahead = 0.0; // (warp 0) squared norm of row `lane` of the speculated next block
    int xs_next = 0;       // next unused exchange sequence number
    int xs_cur = 0;        // exchange that carries the entries of the block at r0
    int xs_stale = -1, stale_rows = 0;  // a speculative exchange that was sent but will not be used (drained below)
    double E1 = 0.0, E2 = 0.0;  // certified displacement bounds of the two preceding blocks
    while (r0 < A.n) {
        int nb = B - (int)(r0 & (kGroup - 1));  // blocks end on group boundaries
        if ((long long)nb > A.n - r0) nb = (int)(A.n - r0);
        my_n = kc > rank ? (kc - rank + ncta - 1) / ncta : 0;
        Dec *dec = dec_all + cur * B;  // decisions of this block (the owner warps may still read the previous block's)
        int *ctl = ctl_all + cur * 4;  // [0] rows committed [1] exact flag [2] centroid count
        double *ctld = ctld_all + cur;
        double P = 0.0;
        if (!have) {
            __syncthreads();  // every update is applied; nobody reads the ring, D or dec of earlier blocks
            if (is_compute) {  // only the compute warps read the row ring
                const long long ga = r0 / kGroup;
                need_groups(ga, (ga + NG) < g_total ? (ga + NG) : g_total);
            }
            if (is_compute) phase1(r0, nb, Dbuf(cur), Pbuf(cur), nullptr, 1, ncomp, 0);
            if (xs_stale >= 0) {  // consume the exchange of the dropped speculative block (keeps the barrier phases in step)
                if (warp == 0) recv_block(stale_rows, xs_stale);
                xs_stale = -1;
            }
            __syncthreads();
            xs_cur = xs_next++;
            if (warp == 0) send_block(nb, Pbuf(cur), xs_cur);
        } else {
            P = E1 + E2;
        }
        ASB_TICKP(0);
        // ---- speculation: the next block starts where this one ends
        const long long r0n = r0 + nb;
        const int nbn = (A.n - r0n) < (long long)B ? (int)(A.n - r0n) : B;
        const bool spec = !A.force_exact && !A.tile_check && kc == maxk && r0n < A.n;
        if (spec && is_compute) {
            const long long ga = r0n / kGroup;
            need_groups(ga, (ga + NG) < g_total ? (ga + NG) : g_total);
        }
        if (is_compute) {   // L2-prefetch the FP64 rows the updates of the next block will read
            const long long lines_per_row = ((long long)f * 8 + 127) / 128;
            const long long first = r0n * lines_per_row, total = (long long)B * lines_per_row;
            const int na = ncomp * 32, ia = tid - 32;
            for (long long l = (long long)rank * na + ia; l < total; l += (long long)ncta * na) {
                const long long byte = (first + l) * 128;
                if (byte < A.n * (long long)f * 8)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A.rows) + byte));
            }
            __syncwarp();
        }

        // ---- resolve the rows of this block IN ORDER (warp 0, identical in every CTA); lane i holds row i
        if (warp == 0) {
            // |x|^2 of this block's rows: loaded one block ahead when the block was speculated (an HBM round trip
            // would otherwise sit on the critical path)
            const double nx = have ? nx_ahead : ((lane < nb) ? __ldg(A.rows_n2 + r0 + lane) : 0.0);
            if (spec) nx_ahead = (lane < nbn) ? __ldg(A.rows_n2 + r0n + lane) : 0.0;
            recv_block(nb, xs_cur);  // the 16 x nb entries of this block, 16 tx-bytes each
            ASB_TICKP(1);
            probe(-1, 0.0);
            GRow32 g;
            g.bd = g.sb_hi = g.sb_lo = g.ss_lo = INFINITY;
            g.bc = kNone;
            if (lane < nb) {  // reduce the 16 CTA entries of row `lane`, attach the certified distance bounds
                const Xch32 *e = xch + ((size_t)(xs_cur & (kXBuf - 1)) * 16) * B + lane;
                // 4 interleaved merge chains over the 16 CTA entries (depth 4 + 2 instead of 16)
                float b4[4], s4[4];
                int c4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    b4[k] = s4[k] = INFINITY;
                    c4[k] = kNone;
                }
#pragma unroll
                for (int q0 = 0; q0 < 16; q0 += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (q0 + k < ncta) {
                            const Xch32 v = e[(size_t)(q0 + k) * B];
                            merge_best(b4[k], s4[k], c4[k], v.bd, v.sd, v.bc);
                        }
                    }
                }
                merge_best(b4[0], s4[0], c4[0], b4[1], s4[1], c4[1]);
                merge_best(b4[2], s4[2], c4[2], b4[3], s4[3], c4[3]);
                merge_best(b4[0], s4[0], c4[0], b4[2], s4[2], c4[2]);
                const float bd = b4[0], sd = s4[0];
                const int bc = c4[0];
                const double d2b = nx + (double)bd, d2s = nx + (double)sd;  // squared distances, +- delta
                g.bd = fmax(d2b, 0.0);
                // FP32 square roots (2 ulp incl. the conversion) with outward slack
                g.sb_hi = (double)sqrtf((float)(d2b + delta)) * (1.0 + 3e-7) + eta;
                g.sb_lo = fmax((double)sqrtf((float)fmax(d2b - delta, 0.0)) * (1.0 - 3e-7) - eta, 0.0);
                g.ss_lo = fmax((double)sqrtf((float)fmax(d2s - delta, 0.0)) * (1.0 - 3e-7) - eta, 0.0);
                g.bc = bc;
                g.pad = 0;
                G[lane] = g;  // the exact path reads G[0]
            }
            __syncwarp();
            probe(43, g.sb_hi);
            int n_commit = 0, exact = 0, kcl = kc;
            int my_action = 3, my_target = -1;
            double my_knew = 0.0, e_own = 0.0;
            bool done = false;
            if (!A.force_exact && kc == maxk) {  // fast path: certify the whole block at once
                const int b = (g.bc == kNone) ? 0 : g.bc;
                const unsigned long long cb = cnt[b];
                const float inv = lane < nb ? 1.0f / (float)(cb + 1ull) : 0.0f;
                float qsum = inv;
                double esum = lane < nb ? g.sb_hi * (double)inv : 0.0;
                for (int o = 16; o > 0; o >>= 1) {
                    qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
                    esum += __shfl_xor_sync(0xffffffffu, esum, o);
                }
                bool ok = (qsum < 0.9f) && (esum < INFINITY);
                // own displacement of the block: E >= sum_j (sb_hi_j + P + E) / (cnt_j + 1)
                // (upper bound of 1 / (1 - qsum) from directed FP32 roundings instead of an FP64 division)
                const double E =
                    (esum + P * (double)qsum) * (double)__frcp_ru(__fsub_rd(1.0f, qsum)) * 1.001 + 1e-300;
                const double T = P + E;  // staleness + everything earlier rows of the block can add
                const double hi_b = g.sb_hi + T, lo_b = fmax(g.sb_lo - T, 0.0);
                const double lo2 = lo_b * lo_b * (1.0 - 1e-12), hi2 = hi_b * hi_b * (1.0 + 1e-12);
                bool row_ok = (g.bd < INFINITY) && ((g.ss_lo - T) > hi_b * (1.0 + 1e-12)) &&
                              !((r_full >= lo2 && r_full <= hi2) || (r_relax >= lo2 && r_relax <= hi2));
                if (lane >= nb) row_ok = true;
                ok = ok && __all_sync(0xffffffffu, row_ok);
                if (ok) {
                    int action = 3;
                    if (lane < nb) action = (g.bd <= r_full) ? 1 : ((g.bd <= r_relax) ? 2 : 3);
                    const bool counts = (action == 1 || action == 2);
                    const unsigned m = __match_any_sync(0xffffffffu, counts ? b : (-1 - lane));
                    const int lower = __popc(m & ((1u << lane) - 1u));
                    if (counts) {
                        my_knew = (double)(cb + (unsigned long long)lower) + 1.0;
                        if ((m >> lane) == 1u) cnt[b] = cb + (unsigned long long)__popc(m);
                    }
                    my_action = action;
                    my_target = counts ? b : -1;
                    n_commit = nb;
                    e_own = E;
                    done = true;
                }
            }
            probe(44, e_own);
            if (!done) {
                int nmod = 0;
                double dmax = 0.0;
                bool created = false;
                for (int i = 0; i < nb && !created; ++i) {
                    const double bd = __shfl_sync(0xffffffffu, g.bd, i);
                    const double sb_hi = __shfl_sync(0xffffffffu, g.sb_hi, i);
                    const double sb_lo = __shfl_sync(0xffffffffu, g.sb_lo, i);
                    const double ss_lo = __shfl_sync(0xffffffffu, g.ss_lo, i);
                    const int bc = __shfl_sync(0xffffffffu, g.bc, i);
                    int action, target;
                    double knew = 0.0;
                    if (kcl == 0) {
                        action = 0;
                        target = 0;
                    } else {
                        const int b = (bc == kNone) ? 0 : bc;
                        const double mod_b = disp[b];
                        const unsigned long long cb = cnt[b];
                        bool ok = !A.force_exact && (bd < INFINITY);
                        const double hi_b = sb_hi + mod_b + P, lo_b = fmax(sb_lo - mod_b - P, 0.0);
                        if (!((ss_lo - dmax - P) > hi_b * (1.0 + 1e-12))) ok = false;
                        const double lo2 = lo_b * lo_b * (1.0 - 1e-12), hi2 = hi_b * hi_b * (1.0 + 1e-12);
                        if ((kcl < maxk && r_half >= lo2 && r_half <= hi2) || (r_full >= lo2 && r_full <= hi2) ||
                            (r_relax >= lo2 && r_relax <= hi2))
                            ok = false;
                        if (!ok) {
                            if (i == 0) exact = 1;
                            break;
                        }
                        const double d2 = bd;  // lies inside [lo2, hi2]; every value there gives the same outcome
                        if (kcl < maxk && d2 > r_half) {
                            action = 0;
                            target = kcl;
                        } else if (d2 <= r_full) {
                            action = 1;
                            target = b;
                        } else if (d2 <= r_relax) {
                            action = 2;
                            target = b;
                        } else {
                            action = 3;
                            target = -1;
                        }
                        if (action == 1) {
                            knew = (double)cb + 1.0;
                            const double nd = mod_b + hi_b * (double)(1.001f / (float)knew) + 1e-300;
                            if (lane == 0) {
                                if (mod_b == 0.0) modlist[nmod] = b;
                                disp[b] = nd;
                            }
                            if (mod_b == 0.0) nmod++;
                            dmax = fmax(dmax, nd);
                        }
                        if (lane == 0 && (action == 1 || action == 2)) cnt[b] = cb + 1ull;
                    }
                    if (action == 0) {
                        if (lane == 0) cnt[target] = 1ull;
                        kcl++;
                        created = true;
                    }
                    if (lane == i) {
                        my_action = action;
                        my_target = target;
                        my_knew = knew;
                    }
                    n_commit++;
                    __syncwarp();
                }
                __syncwarp();
                if (lane == 0)
                    for (int m = 0; m < nmod; ++m) disp[modlist[m]] = 0.0;
                e_own = dmax;
            }
            probe(45, e_own);
            if (lane < n_commit) {
                Dec dd;
                dd.knew = my_knew;
                dd.action = my_action;
                dd.target = my_target;
                const int tt = my_target < 0 ? 0 : my_target;
                dd.owner = tt % ncta;
                dd.slot = tt / ncta;
                dd.owarp = 0;
                dd.pad = 0;
                dec[lane] = dd;
            }
            if (lane == 0) {
                ctl[0] = n_commit;
                ctl[1] = exact;
                ctl[2] = kcl;
                ctld[0] = e_own;
                item_ctr[(n_blocks + 1) & 1] = 0;  // the work counters of the next iteration
                item_ctr[2 + ((n_blocks + 1) & 1)] = 0;
            }
            __syncwarp();
            probe(46, 0.0);
            ASB_TICKP(2);
        }
        // ---- meanwhile: distances of the next block (warp 0 joins when it has resolved this one)
        const int xs_spec = xs_next;  // exchange of the speculative block (sent by the compute warp that finishes it)
        if (spec) xs_next++;
        if (spec && is_compute)
            phase1(r0n, nbn, Dbuf(cur ^ 1), Pbuf(cur ^ 1), &item_ctr[n_blocks & 1], 1, ncomp, xs_spec);
        ASB_TICKP(3);
        {
            long long tb = 0;
            if (A.phase_times) {
                asm volatile("mov.u64 %0, %%clock64;" : "=l"(tb)::"memory");
                if (lane == 0) arr_t[warp] = tb;
            }
            __syncthreads();
            if (A.phase_times && tid == 0) {  // debug: which warp arrived last, and how long after the control warp
                int last = 0;
                for (int w = 1; w < nw; ++w)
                    if (arr_t[w] > arr_t[last]) last = w;
                arr_hist[last] += 1;
                arr_late[last] += arr_t[last] - arr_t[0];
            }
        }
        ASB_TICKP(4);
        int n_commit = ctl[0];
        const int exact = ctl[1];

        if (exact) {
            // ---- exact path for row r0: FP64 reference arithmetic for every candidate the FP32 bounds allow
            n_exact++;
            const GRow32 g0 = G[0];
            const float *D = Dbuf(cur);
            for (int j = tid; j < f; j += blockDim.x) xrow64[j] = A.rows[r0 * (long long)f + j];
            __syncthreads();
            // candidate: its lower distance bound (staleness included) does not exceed the winner's upper bound
            const double lim = g0.sb_hi + 2.0 * P + eta;
            const double thrq = lim * lim * (1.0 + 1e-12) + delta - A.rows_n2[r0];  // on q = d2 - |x|^2
            double my_d = INFINITY;
            int my_c = kNone;
            for (int s = tid; s < my_n; s += blockDim.x) {
                const int c = s * ncta + rank;
                if (A.force_exact || (double)D[(size_t)s * B] <= thrq || !(thrq < INFINITY)) {
                    const double *cv = c64(s);
                    double d2 = 0.0;
                    for (int j = 0; j < f; ++j) {  // src/clustering.rs:917-921
                        const double diff = __dsub_rn(xrow64[j], cv[j]);
                        d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
                    }
                    if (!(d2 == d2)) d2 = INFINITY;
                    if (lex_less(d2, c, my_d, my_c)) {
                        my_d = d2;
                        my_c = c;
                    }
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, my_d, o);
                const int oc = __shfl_xor_sync(0xffffffffu, my_c, o);
                if (lex_less(od, oc, my_d, my_c)) {
                    my_d = od;
                    my_c = oc;
                }
            }
            if (lane == 0) {
                wred_d[warp] = my_d;
                wred_c[warp] = my_c;
            }
            __syncthreads();
            if (warp == 0) {
                double bd = lane < nw ? wred_d[lane] : INFINITY;
                int bc = lane < nw ? wred_c[lane] : kNone;
                for (int o = 16; o > 0; o >>= 1) {
                    const double obd = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                    if (lex_less(obd, obc, bd, bc)) {
                        bd = obd;
                        bc = obc;
                    }
                }
                if (lane < ncta) {
                    Xch *remote = cluster.map_shared_rank(xch_exact, lane) + cur * 16 + rank;
                    Xch v;
                    v.best_d = bd;
                    v.second_d = INFINITY;
                    v.best_c = bc;
                    v.pad = 0;
                    *remote = v;
                }
            }
            cluster.sync();
            if (tid == 0) {
                double bd = INFINITY;
                int bc = kNone;
                for (int q = 0; q < ncta; ++q) {
                    const Xch e = xch_exact[cur * 16 + q];
                    if (lex_less(e.best_d, e.best_c, bd, bc)) {
                        bd = e.best_d;
                        bc = e.best_c;
                    }
                }
                const int b = (bc == kNone) ? 0 : bc;
                int action, target;
                double knew = 0.0, e_own = 0.0;
                int kcl = ctl[2];
                if (kcl < maxk && bd > r_half) {
                    action = 0;
                    target = kcl;
                    cnt[target] = 1ull;
                    kcl++;
                } else if (bd <= r_full) {
                    action = 1;
                    target = b;
                    knew = (double)cnt[b] + 1.0;
                    cnt[b] += 1ull;
                    e_own = sqrt(bd) * 1.001 / knew + 1e-300;  // the update moves centroid b by |x - c| / k_new
                } else if (bd <= r_relax) {
                    action = 2;
                    target = b;
                    cnt[b] += 1ull;
                } else {
                    action = 3;
                    target = -1;
                }
                Dec dd;
                dd.knew = knew;
                dd.action = action;
                dd.target = target;
                const int tt = target < 0 ? 0 : target;
                dd.owner = tt % ncta;
                dd.slot = tt / ncta;
                dd.owarp = 0;
                dd.pad = 0;
                dec[0] = dd;
                ctl[0] = 1;
                ctl[2] = kcl;
                ctld[0] = e_own;
            }
            __syncthreads();
            n_commit = 1;
        }
        const int kc_new = ctl[2];
        const double e_blk = ctld[0];
        const bool ok_spec = spec && !exact && n_commit == nb && kc_new == kc;
        kc = kc_new;

        if (rank == 0 && tid < n_commit) A.assign[r0 + tid] = (long long)dec[tid].target;
        if (ok_spec) {
            xs_cur = xs_spec;
        } else if (spec) {
            xs_stale = xs_spec;
            stale_rows = nbn;
        }
        ASB_TICKP(5);
        // ---- apply the committed decisions in row order.  The update c += (x - c) / k is element-wise,
        // so the apply warps split the FEATURES: every apply warp handles its slice of every row this CTA owns
        // (FP64 row and centroid from global memory, both centroid copies written; the slice of the centroid
        // stays in registers across consecutive rows of one slot, the next row is loaded ahead).
        if (is_apply) {
            bool mine = false;
            if (lane < n_commit) {
                const Dec dd = dec[lane];
                mine = dd.action != 3 && dd.action != 2 && dd.owner == rank;
            }
            unsigned todo = __ballot_sync(0xffffffffu, mine);
            const int jlo = (warp - first_apply) * feat_per_apply;
            const int jhi = (jlo + feat_per_apply) < f ? (jlo + feat_per_apply) : f;
            constexpr int T = 4;
            if (feat_per_apply <= 32 * T) {
                // software pipeline: the row slice, the centroid slice and the reciprocal of the count of the NEXT row
                // are in flight while the current row is applied (a row that hits the same slot as its predecessor
                // takes the fresh registers)
                double cw[T], xv[T];
                int i = todo ? __ffs(todo) - 1 : -1;
                todo &= todo - 1;
                Dec dd;
                dd.slot = -1;
                dd.action = 3;
                dd.knew = 1.0;
                double yk = 1.0;
                if (i >= 0) {
                    dd = dec[i];
                    const double *row = A.rows + (r0 + i) * (long long)f;
                    const double *cv = c64(dd.slot);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const int j = jlo + lane + 32 * t;
                        xv[t] = j < jhi ? __ldg(row + j) : 0.0;
                        cw[t] = (j < jhi && dd.action != 0) ? cv[j] : 0.0;
                    }
                    yk = __drcp_rn(dd.action != 0 ? dd.knew : 1.0);
                }
                while (i >= 0) {
                    const int inext = todo ? __ffs(todo) - 1 : -1;
                    todo &= todo - 1;
                    double xn[T], cnx[T], yn = 1.0;
                    Dec dn = dd;
                    bool fresh = false;  // the next row needs a centroid slice from memory
                    if (inext >= 0) {
                        dn = dec[inext];
                        fresh = dn.action != 0 && dn.slot != dd.slot;
                        const double *row = A.rows + (r0 + inext) * (long long)f;
                        const double *cvn = c64(dn.slot);
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            const int j = jlo + lane + 32 * t;
                            xn[t] = j < jhi ? __ldg(row + j) : 0.0;
                            cnx[t] = (j < jhi && fresh) ? cvn[j] : 0.0;
                        }
                        yn = __drcp_rn(dn.action != 0 ? dn.knew : 1.0);
                    }
                    double *cv = c64(dd.slot);
                    float *cf = c32(dd.slot);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const int j = jlo + lane + 32 * t;
                        if (j < jhi) {
                            double v = xv[t];
                            if (dd.action != 0)  // c + (x - c) / k, src/clustering.rs:748
                                v = __dadd_rn(cw[t], div_by_count(__dsub_rn(v, cw[t]), dd.knew, yk));
                            cw[t] = v;
                            cv[j] = v;
                            cf[j] = (float)v;
                        }
                    }
                    if (inext >= 0) {
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            xv[t] = xn[t];
                            if (fresh) cw[t] = cnx[t];
                        }
                    }
                    dd = dn;
                    yk = yn;
                    i = inext;
                }
            } else {
                while (todo) {
                    const int i = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const Dec dd = dec[i];
                    double *cv = c64(dd.slot);
                    float *cf = c32(dd.slot);
                    const double *row = A.rows + (r0 + i) * (long long)f;
                    for (int jb = jlo; jb < jhi; jb += 32 * 6) {
                        double xv[6], cw[6];
#pragma unroll
                        for (int t = 0; t < 6; ++t) {
                            const int j = jb + lane + 32 * t;
                            xv[t] = j < jhi ? __ldg(row + j) : 0.0;
                            cw[t] = (j < jhi && dd.action != 0) ? cv[j] : 0.0;
                        }
#pragma unroll
                        for (int t = 0; t < 6; ++t) {
                            const int j = jb + lane + 32 * t;
                            if (j < jhi) {
                                double v = xv[t];
                                if (dd.action != 0)
                                    v = __dadd_rn(cw[t], __ddiv_rn(__dsub_rn(v, cw[t]), dd.knew));  // :748
                                cv[j] = v;
                                cf[j] = (float)v;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        E2 = E1;
        E1 = e_blk;
        r0 += n_commit;
        n_blocks++;
        cur ^= 1;
        have = ok_spec;
        ASB_TICKP(6);
    }
    __syncthreads();
    cluster.sync();
    if (rank == 0) {
        for (int c = tid; c < kc; c += blockDim.x) A.sizes[c] = cnt[c];
        if (tid == 0) {
            A.x_out[0] = kc;
            A.stats[0] = n_exact;
            A.stats[1] = (int)(n_blocks > 0x7fffffff ? 0x7fffffff : n_blocks);
        }
    }
    __syncthreads();
    if (rank == 0 && tid == 0 && A.phase_times)
        for (int k = 0; k < 8; ++k) A.phase_times[k] = tphase[k];
    if (rank == 0 && tid < nw && tid < 24 && A.phase_times) A.phase_times[8 + tid] = arr_hist[tid] * 1000000ll + (arr_hist[tid] ? arr_late[tid] / arr_hist[tid] : 0);
}

size_t cluster_f32p_smem_bytes(int f, int slots, int maxk, int ring_groups) {
    const int slots8 = (slots + 7) & ~7;
    const int fp = f32p_pitch(f);
    size_t b = (size_t)ring_groups * kGroup * fp * 4;  // ring (f32)
    b += (size_t)slots8 * fp * 4;                      // cent32
    b += (size_t)2 * slots8 * kB32 * 4;                // D (two buffers)
    b += (size_t)(kXBuf * 16 + 2 * (slots8 / 8)) * kB32 * sizeof(Xch32);  // xch, per-tile arg-mins
    b += 32 * sizeof(Xch);                             // xch_exact
    b += (size_t)kB32 * sizeof(GRow32);
    b += (size_t)2 * kB32 * sizeof(Dec);
    b += (size_t)kPipeGroups * 8 + kXBuf * 8;          // mbarriers
    b += (size_t)maxk * 8 * 2;                         // cnt, disp
    b += 32 * 8 + 2 * 8;                               // wred_d, ctld
    b += (size_t)f * 8;                                // xrow64
    b += 32 * 4 + 8 * 4 + 4 * 4 + (size_t)kB32 * 4;    // wred_c, ctl, item_ctr, modlist
    return b + 96;
}

}  // namespace
