// cluster_f32.cuh -- K2, third variant: 32 rows per cluster barrier with an FP32 distance prefilter.
//
// Same walk, same certified-decision scheme as cluster_block_kernel, with two changes aimed at the
// per-block fixed latency that dominated it (profiles/r01_cluster_phases.md):
//   * the distance tile runs in FP32 on an FP32 copy of the rows (made by rows_to_f32_kernel, one
//     streaming pass) and an FP32 shadow of the centroids: half the shared-memory bytes and the FP32
//     pipe's 2x rate, which buys a 32-row block in the same shared memory (4 centroids x 8 rows
//     register tile per warp);
//   * every FP32 distance carries a rigorous error interval:  |d_true - sqrt(d2_f32)| <=
//     rho * sqrt(d2_f32) + eta  with  rho = (F + 32) * 2^-24  (accumulation) and  eta = 3e-7 * max|x|
//     (input rounding; every centroid is a running mean of rows, so |c| <= max|x|).  Those intervals
//     feed the same interval arithmetic as the centroid displacements, so a decision is still only
//     taken when it is CERTIFIED; everything else goes through the reference-arithmetic FP64 path.
//     FP64 rows are needed only by the owner warp that applies an update (read straight from
//     global memory, L2-prefetched one block ahead) and by the exact path.  The update of block k
//     overlaps with the distance phase of block k+1 (per-centroid done/need counters instead of a
//     CTA barrier).
// Outputs are bit-identical to the reference for every input, like the other two variants.
#pragma once

namespace {

constexpr int kB32 = 32;          // rows per block
constexpr int kRingGroups = 6;    // ring = 6 groups of 8 rows (4 in use + 2 in flight)

struct __align__(16) Xch32 {
    float bd, sd;
    int bc, pad;
};

struct __align__(16) GRow32 {
    double bd;             // best squared distance (from FP32)
    double sb_hi, sb_lo;   // certified bounds of the best distance
    double ss_lo;          // certified lower bound of the second-best distance
    int bc, pad;
};

// rows (f64) -> rows32 (f32) + max finite squared norm.  One warp per row.
__global__ void __launch_bounds__(256) rows_to_f32_kernel(const double *__restrict__ rows, long long n, int f,
                                                          float *__restrict__ rows32,
                                                          unsigned long long *__restrict__ max_norm2_bits,
                                                          double *__restrict__ rows_n2, int pitch) {
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const double *r = rows + w * (long long)f;
    float *o = rows32 ? rows32 + w * (long long)pitch : nullptr;  // pitch >= f; the padding is zero
    double s = 0.0, s32 = 0.0;
    for (int j = lane; j < f; j += 32) {
        const double v = __ldg(r + j);
        const float vf = (float)v;
        if (o) o[j] = vf;
        s = fma(v, v, s);
        s32 = fma((double)vf, (double)vf, s32);
    }
    if (o)
        for (int j = f + lane; j < pitch; j += 32) o[j] = 0.0f;
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        s32 += __shfl_xor_sync(0xffffffffu, s32, off);
    }
    if (lane == 0 && s < INFINITY) atomicMax(max_norm2_bits, (unsigned long long)__double_as_longlong(s));
    if (lane == 0 && rows_n2) rows_n2[w] = s32;  // squared norm of the FP32-rounded row (FP64, ~1e-16 relative)
}

// acc[c * 4 + i] += |row_i - cent_c|^2 over the 128 features (j0 + 4 lane .. +3), FP32, 16-byte loads:
// register tile of 4 centroids x 4 rows (the 4 rows of an item are consecutive and 4-aligned inside the
// block, so they never straddle a copy group when the block start is a multiple of 4; otherwise `split`).
__device__ __forceinline__ void dist_tile4x4_f32(float (&acc)[16], const float *__restrict__ c0,
                                                 const float *__restrict__ c1, const float *__restrict__ c2,
                                                 const float *__restrict__ c3, const float *ring, int off_a,
                                                 int off_b, int split, int f, int j0, int lane) {
    const int o = j0 + 4 * lane;
    const float4 a0 = *reinterpret_cast<const float4 *>(c0 + o);
    const float4 a1 = *reinterpret_cast<const float4 *>(c1 + o);
    const float4 a2 = *reinterpret_cast<const float4 *>(c2 + o);
    const float4 a3 = *reinterpret_cast<const float4 *>(c3 + o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ro = (i < split) ? off_a + i * f : off_b + (i - split) * f;
        const float4 x = *reinterpret_cast<const float4 *>(ring + ro + o);
        float d;
#define ASB_ACC(A_, IDX)                                   \
        d = x.x - A_.x; acc[IDX] = fmaf(d, d, acc[IDX]);   \
        d = x.y - A_.y; acc[IDX] = fmaf(d, d, acc[IDX]);   \
        d = x.z - A_.z; acc[IDX] = fmaf(d, d, acc[IDX]);   \
        d = x.w - A_.w; acc[IDX] = fmaf(d, d, acc[IDX]);
        ASB_ACC(a0, i)
        ASB_ACC(a1, 4 + i)
        ASB_ACC(a2, 8 + i)
        ASB_ACC(a3, 12 + i)
#undef ASB_ACC
    }
}

// Sum 16 per-lane partial accumulators over the 32 lanes; on return acc[0] of lane L is element L >> 1.
__device__ __forceinline__ void transpose_reduce16f(float (&acc)[16], int lane) {
    int off = 16;
#pragma unroll
    for (int cur = 16; cur > 1; cur >>= 1) {
        const int half = cur >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int h = 0; h < half; ++h) {
            const float send = upper ? acc[h] : acc[h + half];
            const float keep = upper ? acc[h + half] : acc[h];
            acc[h] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
}

__device__ __forceinline__ bool lex_less_f(float d1, int c1, float d2, int c2) {
    return d1 < d2 || (d1 == d2 && c1 < c2);
}

template <bool CSMEM>
__global__ void __launch_bounds__(768, 1) cluster_f32_kernel(ClusterArgs A) {
    constexpr int B = kB32;
    constexpr int RROWS = kRingGroups * kGroup;
    cg::cluster_group cluster = cg::this_cluster();
    const int ncta = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int f = A.f;
    const int cp = block_cent_pitch(f);
    const int slots = A.slots_per_cta;
    const int slots4 = (slots + 3) & ~3;  // the 4-centroid tile may touch up to 3 padding slots
    const int maxk = A.max_k;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);                         // RROWS * f
    float *cent32 = ring + (size_t)RROWS * f;                                  // slots4 * f
    float *D = cent32 + (size_t)slots4 * f;                                    // slots4 * B
    Xch32 *xch = reinterpret_cast<Xch32 *>(D + (size_t)slots4 * B);            // [2][16][B]
    Xch *xch_exact = reinterpret_cast<Xch *>(xch + 2 * 16 * B);                // [2][16]
    GRow32 *G = reinterpret_cast<GRow32 *>(xch_exact + 32);                    // B
    Dec *dec = reinterpret_cast<Dec *>(G + B);                                 // B
    unsigned long long *full = reinterpret_cast<unsigned long long *>(dec + B);  // kRingGroups mbarriers
    unsigned long long *cnt = full + 8;                                        // maxk (replicated counts)
    double *disp = reinterpret_cast<double *>(cnt + maxk);                     // maxk
    double *wred_d = disp + maxk;                                              // 32
    double *xrow64 = wred_d + 32;                                              // f (exact path row)
    int *wred_c = reinterpret_cast<int *>(xrow64 + f);                         // 32
    int *ctl = wred_c + 32;                                                    // 4
    int *modlist = ctl + 4;                                                    // B
    int *upd_need = modlist + B;                                               // slots4: updates committed per own slot
    int *upd_done = upd_need + slots4;                                         // slots4: updates applied per own slot
    double *cent64 = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(upd_done + slots4) + 15) & ~(uintptr_t)15);

    const double xmax = sqrt(__longlong_as_double((long long)*A.max_norm2_bits)) * (1.0 + 1e-6);
    const double eta = 3e-7 * xmax;                  // input rounding: 2u(|x| + |c|) <= 2.4e-7 * max|x|
    const double rho = (double)(f + 32) * 5.97e-8;   // FP32 accumulation + sqrt + conversions

    // FP64 centroids: distributed shared memory when they fit, else the (L2-resident) global array --
    // only the owner warp (updates) and the exact path ever touch them
    auto c64 = [&](int slot) -> double * {
        if (CSMEM) return cent64 + (size_t)slot * cp;
        return A.centroids + ((size_t)slot * ncta + rank) * f;
    };
    auto c32 = [&](int slot) -> float * { return cent32 + (size_t)slot * f; };
    auto ring_off = [&](long long r) -> int { return (int)(((r >> 3) % kRingGroups) * kGroup + (r & 7)) * f; };

    if (tid == 0) {
        for (int s = 0; s < kRingGroups; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = tid; c < maxk; c += blockDim.x) {
        cnt[c] = (c < A.init_k) ? A.sizes[c] : 0ull;
        disp[c] = 0.0;
    }
    for (int j = tid; j < slots4 * f; j += blockDim.x) cent32[j] = 0.0f;
    for (int j = tid; j < 2 * slots4; j += blockDim.x) upd_need[j] = 0;  // upd_need and upd_done are contiguous
    int kc = A.init_k;
    __syncthreads();
    for (int s = warp; s < slots; s += nw) {  // resume: adopt the state left by the previous shard
        const int c = s * ncta + rank;
        if (c >= kc) break;
        const double *src = A.centroids + (size_t)c * f;
        for (int j = lane; j < f; j += 32) {
            const double v = src[j];
            if (CSMEM) c64(s)[j] = v;
            c32(s)[j] = (float)v;
        }
    }
    const double r_half = A.radius * 0.5, r_full = A.radius, r_relax = A.radius * 1.5;
    const unsigned row_bytes = (unsigned)f * 4u;
    long long next_fetch = 0, waited_groups = 0, r0 = 0, n_blocks = 0;
    int n_exact = 0;
    __shared__ long long tphase[8];  // debug phase timers (thread 0 only)
    if (tid == 0)
        for (int k = 0; k < 8; ++k) tphase[k] = 0;
    long long tlast = clock64();
    __syncthreads();
    cluster.sync();

    while (r0 < A.n) {
        const int par = (int)(n_blocks & 1);
        // ---- request whole groups up to group (r0/8 + 5); L2-prefetch the FP64 rows of the next block
        long long fetch_to = (r0 / kGroup + kRingGroups) * kGroup;
        if (fetch_to > A.n) fetch_to = A.n;
        if (tid == blockDim.x - 32) {  // lane 0 of the last warp
            for (long long r = next_fetch; r < fetch_to; r += kGroup) {
                const long long g = r / kGroup;
                const long long rows_in = (A.n - r) < kGroup ? (A.n - r) : kGroup;
                unsigned long long *bar = &full[g % kRingGroups];
                const unsigned bytes = (unsigned)rows_in * row_bytes;
                mbar_expect_tx(bar, bytes);
                bulk_g2s(ring + (size_t)((g % kRingGroups) * kGroup) * f, A.rows32 + r * (long long)f, bytes, bar);
            }
        }
        __syncwarp();
        next_fetch = fetch_to;
        int nb = B - (int)(r0 & (kGroup - 1));  // blocks end on group boundaries
        if ((long long)nb > A.n - r0) nb = (int)(A.n - r0);
        {
            const long long lines_per_row = ((long long)f * 8 + 127) / 128;
            const long long first = (r0 + nb) * lines_per_row, total = (long long)B * lines_per_row;
            for (long long l = (long long)rank * blockDim.x + tid; l < total; l += (long long)ncta * blockDim.x) {
                const long long byte = (first + l) * 128;
                if (byte < A.n * (long long)f * 8)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A.rows) + byte));
            }
        }
        ASB_TICK(0);

        // ---- 1. FP32 distances.  Work item = (4 of this CTA's centroids) x (4 block rows), row groups first.
        const int my_n = kc > rank ? (kc - rank + ncta - 1) / ncta : 0;
        {
            const int nquads = (my_n + 3) >> 2;
            const int nrg = (nb + 3) >> 2;
            const int nitems = nquads * nrg;
            for (int it = warp; it < nitems; it += nw) {
                const int rg = it / nquads, quad = it - rg * nquads;
                const long long rbase = r0 + 4 * rg;
                {  // wait (once per thread) for the copy groups holding these rows
                    long long g_last = (rbase + 3) / kGroup;
                    const long long g_max = (r0 + nb - 1) / kGroup;
                    if (g_last > g_max) g_last = g_max;
                    for (; waited_groups <= g_last; ++waited_groups)
                        mbar_wait(&full[waited_groups % kRingGroups], (unsigned)((waited_groups / kRingGroups) & 1));
                }
                // rows >= nb read stale ring slots; their results are never consumed
                int split = 8 - (int)(rbase & 7);
                if (split > 4) split = 4;
                const int off_a = ring_off(rbase), off_b = ring_off(rbase + split);
                const int s0 = 4 * quad;
                if (lane < 4) {  // the owner warps of these centroids may still be applying the previous block
                    const volatile int *dn = upd_done + s0 + lane;
                    const int need = upd_need[s0 + lane];
                    while (*dn < need) {
                    }
                }
                __syncwarp();
                const float *p0 = c32(s0), *p1 = c32(s0 + 1), *p2 = c32(s0 + 2), *p3 = c32(s0 + 3);
                float acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
                int j0 = 0;
                for (; j0 + 128 <= f; j0 += 128)
                    dist_tile4x4_f32(acc, p0, p1, p2, p3, ring, off_a, off_b, split, f, j0, lane);
                if (j0 < f && j0 + 4 * lane < f) {  // tail (f % 128 != 0, f % 4 == 0)
                    const int o = j0 + 4 * lane;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float *pcc = cc == 0 ? p0 : (cc == 1 ? p1 : (cc == 2 ? p2 : p3));
                        const float4 a = *reinterpret_cast<const float4 *>(pcc + o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int ro = (i < split) ? off_a + i * f : off_b + (i - split) * f;
                            const float4 x = *reinterpret_cast<const float4 *>(ring + ro + o);
                            float d;
                            d = x.x - a.x; acc[cc * 4 + i] = fmaf(d, d, acc[cc * 4 + i]);
                            d = x.y - a.y; acc[cc * 4 + i] = fmaf(d, d, acc[cc * 4 + i]);
                            d = x.z - a.z; acc[cc * 4 + i] = fmaf(d, d, acc[cc * 4 + i]);
                            d = x.w - a.w; acc[cc * 4 + i] = fmaf(d, d, acc[cc * 4 + i]);
                        }
                    }
                }
                transpose_reduce16f(acc, lane);  // lane L holds element L >> 1 = c * 4 + i
                float tot = acc[0];
                if (!(tot == tot)) tot = INFINITY;  // NaN never wins
                const int e = lane >> 1, cidx = e >> 2, i = e & 3;
                if ((lane & 1) == 0 && s0 + cidx < my_n) D[(size_t)(s0 + cidx) * B + 4 * rg + i] = tot;
            }
            // threads that ran no item still have to observe the copies before phase 4 reads nothing of them;
            // keep waited_groups monotone for everybody
            const long long g_max = (r0 + nb - 1) / kGroup;
            for (; waited_groups <= g_max; ++waited_groups)
                mbar_wait(&full[waited_groups % kRingGroups], (unsigned)((waited_groups / kRingGroups) & 1));
        }
        __syncthreads();
        ASB_TICK(1);

        // ---- 2. per row: arg-min over this CTA's centroids, all-to-all through DSMEM
        for (int i = warp; i < nb; i += nw) {
            float bd = INFINITY, sd = INFINITY;
            int bc = kNone;
            for (int s = lane; s < my_n; s += 32) {
                const float d = D[(size_t)s * B + i];
                const int c = s * ncta + rank;
                if (lex_less_f(d, c, bd, bc)) {
                    sd = bd;
                    bd = d;
                    bc = c;
                } else if (d < sd) {
                    sd = d;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float obd = __shfl_xor_sync(0xffffffffu, bd, o);
                const float osd = __shfl_xor_sync(0xffffffffu, sd, o);
                const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (lex_less_f(obd, obc, bd, bc)) {
                    sd = fminf(fminf(sd, osd), bd);
                    bd = obd;
                    bc = obc;
                } else {
                    sd = fminf(fminf(sd, osd), obd);
                }
            }
            if (lane < ncta) {
                Xch32 *remote = cluster.map_shared_rank(xch, lane) + ((size_t)par * 16 + rank) * B + i;
                Xch32 v;
                v.bd = bd;
                v.sd = sd;
                v.bc = bc;
                v.pad = 0;
                *remote = v;
            }
        }
        ASB_TICK(2);
        cluster.sync();
        ASB_TICK(3);

        // ---- 3b. resolve the rows IN ORDER (warp 0, identical in every CTA); lane i holds row i
        if (warp == 0) {
            GRow32 g;
            g.bd = g.sb_hi = g.sb_lo = g.ss_lo = INFINITY;
            g.bc = kNone;
            if (lane < nb) {  // reduce the 16 CTA entries of row `lane`, attach the certified distance bounds
                const Xch32 *e = xch + ((size_t)par * 16) * B + lane;
                float bd = INFINITY, sd = INFINITY;
                int bc = kNone;
                for (int q = 0; q < ncta; ++q) {
                    const Xch32 v = e[(size_t)q * B];
                    if (lex_less_f(v.bd, v.bc, bd, bc)) {
                        sd = fminf(fminf(sd, v.sd), bd);
                        bd = v.bd;
                        bc = v.bc;
                    } else {
                        sd = fminf(fminf(sd, v.sd), v.bd);
                    }
                }
                const double sb = (double)sqrtf(bd), ss = (double)sqrtf(sd);
                g.bd = (double)bd;
                g.sb_hi = sb * (1.0 + rho) + eta;
                g.sb_lo = fmax(sb * (1.0 - rho) - eta, 0.0);
                g.ss_lo = fmax(ss * (1.0 - rho) - eta, 0.0);
                g.bc = bc;
                g.pad = 0;
                G[lane] = g;  // the exact path reads G[0]
            }
            __syncwarp();
            int n_commit = 0, exact = 0, kcl = kc;
            int my_action = 3, my_target = -1;
            double my_knew = 0.0;
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
                const double E = esum / (1.0 - (double)qsum) * 1.001 + 1e-300;  // >= total displacement of the block
                const double hi_b = g.sb_hi + E, lo_b = fmax(g.sb_lo - E, 0.0);
                const double lo2 = lo_b * lo_b * (1.0 - 1e-12), hi2 = hi_b * hi_b * (1.0 + 1e-12);
                bool row_ok = (g.bd < INFINITY) && ((g.ss_lo - E) > hi_b * (1.0 + 1e-12)) &&
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
                    done = true;
                }
            }
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
                        const double hi_b = sb_hi + mod_b, lo_b = fmax(sb_lo - mod_b, 0.0);
                        if (!((ss_lo - dmax) > hi_b * (1.0 + 1e-12))) ok = false;
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
            }
            if (lane < n_commit) {
                Dec dd;
                dd.knew = my_knew;
                dd.action = my_action;
                dd.target = my_target;
                const int tt = my_target < 0 ? 0 : my_target;
                dd.owner = tt % ncta;
                dd.slot = tt / ncta;
                dd.owarp = dd.slot % nw;
                dd.pad = 0;
                dec[lane] = dd;
                if ((my_action == 0 || my_action == 1) && dd.owner == rank) atomicAdd(&upd_need[dd.slot], 1);
            }
            if (lane == 0) {
                ctl[0] = n_commit;
                ctl[1] = exact;
                ctl[2] = kcl;
            }
        }
        __syncthreads();
        ASB_TICK(5);
        int n_commit = ctl[0];
        const int exact = ctl[1];

        if (exact) {
            // ---- exact path for row r0: FP64 reference arithmetic for every candidate the FP32 bounds allow
            n_exact++;
            const GRow32 g0 = G[0];
            for (int j = tid; j < f; j += blockDim.x) xrow64[j] = A.rows[r0 * (long long)f + j];
            __syncthreads();
            // candidate: its lower distance bound does not exceed the winner's upper bound
            const double lim = (g0.sb_hi + eta) / (1.0 - rho);
            const float thr2 = (float)(lim * lim * (1.0 + 1e-6));
            double my_d = INFINITY;
            int my_c = kNone;
            for (int s = tid; s < my_n; s += blockDim.x) {
                const int c = s * ncta + rank;
                if (A.force_exact || D[(size_t)s * B] <= thr2 || !(thr2 < INFINITY)) {
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
                    Xch *remote = cluster.map_shared_rank(xch_exact, lane) + par * 16 + rank;
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
                    const Xch e = xch_exact[par * 16 + q];
                    if (lex_less(e.best_d, e.best_c, bd, bc)) {
                        bd = e.best_d;
                        bc = e.best_c;
                    }
                }
                const int b = (bc == kNone) ? 0 : bc;
                int action, target;
                double knew = 0.0;
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
                dd.owarp = dd.slot % nw;
                dd.pad = 0;
                dec[0] = dd;
                if ((action == 0 || action == 1) && dd.owner == rank) upd_need[dd.slot] += 1;
                ctl[0] = 1;
                ctl[2] = kcl;
            }
            __syncthreads();
            n_commit = 1;
        }
        kc = ctl[2];

        // ---- 4. apply the committed decisions in row order: FP64 row from global, both centroid copies
        {
            bool mine = false;
            if (lane < n_commit) {
                const Dec dd = dec[lane];
                mine = dd.action != 3 && dd.action != 2 && dd.owner == rank && dd.owarp == warp;
            }
            unsigned todo = __ballot_sync(0xffffffffu, mine);
            while (todo) {
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                const Dec dd = dec[i];
                double *cv = c64(dd.slot);
                float *cf = c32(dd.slot);
                const double *row = A.rows + (r0 + i) * (long long)f;
                // the FP64 row comes from global memory (L2-prefetched): issue the loads of a batch first
                for (int jb = 0; jb < f; jb += 32 * 12) {
                    double xv[12];
#pragma unroll
                    for (int t = 0; t < 12; ++t) {
                        const int j = jb + lane + 32 * t;
                        xv[t] = j < f ? __ldg(row + j) : 0.0;
                    }
#pragma unroll
                    for (int t = 0; t < 12; ++t) {
                        const int j = jb + lane + 32 * t;
                        if (j < f) {
                            double v = xv[t];
                            if (dd.action != 0) {
                                const double c0 = cv[j];
                                v = __dadd_rn(c0, __ddiv_rn(__dsub_rn(v, c0), dd.knew));  // :748
                            }
                            cv[j] = v;
                            cf[j] = (float)v;
                        }
                    }
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) *((volatile int *)(upd_done + dd.slot)) = upd_done[dd.slot] + 1;
            }
        }
        if (rank == 0 && tid < n_commit) A.assign[r0 + tid] = (long long)dec[tid].target;
        r0 += n_commit;
        n_blocks++;
        // No CTA barrier here: the owner warps finish their updates while the other warps start the next
        // block; phase 1 waits per centroid (upd_done >= upd_need).  Everything else the next block
        // overwrites (ring slots, D, dec, ctl) is only written after barriers the owner warps also join.
        ASB_TICK(6);
    }
    __syncthreads();
    cluster.sync();
    for (int s = warp; s < slots; s += nw) {
        const int c = s * ncta + rank;
        if (c >= kc) break;
        if (CSMEM) {
            const double *cv = c64(s);
            double *dst = A.centroids + (size_t)c * f;
            for (int j = lane; j < f; j += 32) dst[j] = cv[j];
        }
    }
    if (rank == 0) {
        for (int c = tid; c < kc; c += blockDim.x) A.sizes[c] = cnt[c];
        if (tid == 0) {
            A.x_out[0] = kc;
            A.stats[0] = n_exact;
            A.stats[1] = (int)(n_blocks > 0x7fffffff ? 0x7fffffff : n_blocks);
            if (A.phase_times)
                for (int k = 0; k < 8; ++k) A.phase_times[k] = tphase[k];
        }
    }
}

size_t cluster_f32_smem_bytes(int f, int slots, int maxk, bool cent64_in_smem) {
    const int slots4 = (slots + 3) & ~3;
    size_t b = (size_t)kRingGroups * kGroup * f * 4;   // ring (f32)
    b += (size_t)slots4 * f * 4;                       // cent32
    b += (size_t)slots4 * kB32 * 4;                    // D
    b += (size_t)2 * 16 * kB32 * sizeof(Xch32);        // xch
    b += 32 * sizeof(Xch);                             // xch_exact
    b += (size_t)kB32 * sizeof(GRow32);
    b += (size_t)kB32 * sizeof(Dec);
    b += 8 * 8;                                        // mbarriers
    b += (size_t)maxk * 8 * 2;                         // cnt, disp
    b += 32 * 8;                                       // wred_d
    b += (size_t)f * 8;                                // xrow64
    b += 32 * 4 + 4 * 4 + (size_t)kB32 * 4;            // wred_c, ctl, modlist
    b += (size_t)slots4 * 4 * 2;                       // upd_need, upd_done
    if (cent64_in_smem) b += (size_t)slots * block_cent_pitch(f) * 8;  // cent64
    return b + 96;
}

}  // namespace
