// cluster_block.cuh -- K2, blocked variant: the same order-dependent walk as cluster_rowwise_kernel
// (src/clustering.rs:547-928) but B rows per cluster barrier instead of one.
//
// Per block of B rows (B = 16 or 8):
//   1. every warp computes the fast squared distances from ALL B rows to the centroid(s) it owns
//      (B independent accumulators -> the FP64 pipe is throughput- not latency-bound), reduced
//      over lanes with a transposing butterfly (16 shuffles instead of 80);
//   2. per row: CTA arg-min over its centroids, then the 16 x 16 all-to-all through distributed
//      shared memory -- ONE cluster barrier per block;
//   3. every CTA derives, redundantly and identically, the decisions for the rows IN ORDER.  Row 0
//      sees the exact snapshot.  Row i > 0 sees centroids that rows 0..i-1 of the block may have
//      moved; a running-mean update moves centroid b by exactly |x - c_b| / k_new, so by the
//      triangle inequality the true distance of any later row to b lies within +-disp[b] of its
//      snapshot distance.  The decision for row i is taken from the snapshot only when that
//      interval arithmetic CERTIFIES it (arg-min separated, no threshold inside the interval);
//      the first row that cannot be certified ends the block and becomes row 0 of the next one.
//      A row that creates a centroid also ends the block (later rows need distances to it).
//      Row 0 itself falls back to the reference-arithmetic exact path when its margins are below
//      delta (same rule as the row-wise kernel).
//   4. owner warps apply the committed updates in row order with the reference's own IEEE
//      operations; rank 0 writes the assignments.
// Rows stream through a shared-memory ring filled by 1-D bulk async copies (cp.async.bulk, the
// TMA engine; SASS UBLKCP) completing on per-slot mbarriers.
#pragma once

namespace {

constexpr int kNone = 0x7fffffff;

struct __align__(16) GRow {
    double bd, sd;  // best / second best squared snapshot distance
    double sb, ss;  // their square roots
    int bc;
    int pad;
};

struct __align__(16) Dec {
    double knew;
    int action;  // 0 new centroid, 1 running-mean update, 2 count only, 3 drop
    int target;
    int owner;   // CTA rank that stores the target centroid
    int slot;    // slot inside that CTA
    int owarp;   // warp of that CTA that owns the slot
    int pad;
};

constexpr int kGroup = 8;  // rows per bulk copy / mbarrier

__host__ __device__ inline int block_cent_pitch(int f) {
    const int fpad = (f + 1) & ~1;
    return ((fpad / 2) & 1) ? fpad : fpad + 2;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity));
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Sum B per-lane partial accumulators over the 32 lanes with a transposing butterfly; on return
// acc[0] of lane L holds the total of row (L >> kShift), kShift = 1 (B = 16) or 2 (B = 8).
template <int B>
__device__ __forceinline__ void transpose_reduce(double (&acc)[B], int lane) {
    int off = 16;
#pragma unroll
    for (int cur = B; cur > 1; cur >>= 1) {
        const int half = cur >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int h = 0; h < half; ++h) {
            const double send = upper ? acc[h] : acc[h + half];
            const double keep = upper ? acc[h + half] : acc[h];
            acc[h] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
    for (; off > 0; off >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
}

// acc[i] += sum over T features (j0 + lane + 32 t) of (row_i - c)^2 for the B rows of the block.
template <int B, int T>
__device__ __forceinline__ void dist_chunk(double (&acc)[B], const double *__restrict__ cv, const double *ring,
                                           int s0, int ring_mask, int fpad, int j0, int lane) {
    double cr[T];
#pragma unroll
    for (int t = 0; t < T; ++t) cr[t] = cv[j0 + lane + 32 * t];
#pragma unroll
    for (int i = 0; i < B; ++i) {
        const double *row = ring + (size_t)((s0 + i) & ring_mask) * fpad + j0 + lane;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double df = row[32 * t] - cr[t];
            acc[i] = fma(df, df, acc[i]);
        }
    }
}

// Same with 16-byte shared-memory loads: the lane owns the feature pairs (j0 + 2 lane + 64 t, +1).
template <int B, int T>
__device__ __forceinline__ void dist_chunk2(double (&acc)[B], const double *__restrict__ cv, const double *ring,
                                            int s0, int ring_mask, int fpad, int j0, int lane) {
    double2 cr[T];
#pragma unroll
    for (int t = 0; t < T; ++t) cr[t] = *reinterpret_cast<const double2 *>(cv + j0 + 2 * lane + 64 * t);
#pragma unroll
    for (int i = 0; i < B; ++i) {
        const double *row = ring + (size_t)((s0 + i) & ring_mask) * fpad + j0 + 2 * lane;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double2 x = *reinterpret_cast<const double2 *>(row + 64 * t);
            const double d0 = x.x - cr[t].x, d1 = x.y - cr[t].y;
            acc[i] = fma(d0, d0, acc[i]);
            acc[i] = fma(d1, d1, acc[i]);
        }
    }
}

// Register tile of 2 centroids x 8 rows: every row element fetched from shared memory feeds two
// centroids (the distance loop was shared-memory-bandwidth bound with one centroid per warp,
// profiles/r01_cluster_v4).  acc[c * 8 + i]; 16-byte loads, lane owns features (j0 + 2 lane + 64 t, +1).
template <int T>
__device__ __forceinline__ void dist_tile2x8_v2(double (&acc)[16], const double *__restrict__ cv0,
                                                const double *__restrict__ cv1, const double *ring, int s0,
                                                int ring_mask, int fpad, int j0, int lane) {
    double2 c0[T], c1[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        c0[t] = *reinterpret_cast<const double2 *>(cv0 + j0 + 2 * lane + 64 * t);
        c1[t] = *reinterpret_cast<const double2 *>(cv1 + j0 + 2 * lane + 64 * t);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double *row = ring + (size_t)((s0 + i) & ring_mask) * fpad + j0 + 2 * lane;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double2 x = *reinterpret_cast<const double2 *>(row + 64 * t);
            double d;
            d = x.x - c0[t].x; acc[i] = fma(d, d, acc[i]);
            d = x.y - c0[t].y; acc[i] = fma(d, d, acc[i]);
            d = x.x - c1[t].x; acc[8 + i] = fma(d, d, acc[8 + i]);
            d = x.y - c1[t].y; acc[8 + i] = fma(d, d, acc[8 + i]);
        }
    }
}
// scalar variant: lane owns features j0 + lane + 32 t
template <int T>
__device__ __forceinline__ void dist_tile2x8(double (&acc)[16], const double *__restrict__ cv0,
                                             const double *__restrict__ cv1, const double *ring, int s0,
                                             int ring_mask, int fpad, int j0, int lane) {
    double c0[T], c1[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        c0[t] = cv0[j0 + lane + 32 * t];
        c1[t] = cv1[j0 + lane + 32 * t];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double *row = ring + (size_t)((s0 + i) & ring_mask) * fpad + j0 + lane;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double x = row[32 * t];
            double d;
            d = x - c0[t]; acc[i] = fma(d, d, acc[i]);
            d = x - c1[t]; acc[8 + i] = fma(d, d, acc[8 + i]);
        }
    }
}

template <int B>
__global__ void __launch_bounds__(768, 1) cluster_block_kernel(ClusterArgs A) {
    constexpr int R = 4 * kGroup;                  // ring rows (4 groups of 8; B <= 16 spans <= 3 groups)
    cg::cluster_group cluster = cg::this_cluster();
    const int ncta = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int f = A.f;
    const int fpad = (f + 1) & ~1;
    const int cp = block_cent_pitch(f);  // even (16 B rows) with an odd number of 16 B granules
    const int slots = A.slots_per_cta;
    const int maxk = A.max_k;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);                    // R * fpad
    double *D = ring + (size_t)R * fpad;                                    // slots * B
    double *disp = D + (size_t)slots * B;                                   // maxk (accumulated displacement)
    double *wred_d = disp + ((maxk + 1) & ~1);                              // 32
    Xch *xch = reinterpret_cast<Xch *>(wred_d + 32);                        // [2][16][B]
    Xch *xch_exact = xch + 2 * 16 * B;                                      // [2][16]
    GRow *G = reinterpret_cast<GRow *>(xch_exact + 32);                     // B
    Dec *dec = reinterpret_cast<Dec *>(G + B);                              // B
    unsigned long long *full = reinterpret_cast<unsigned long long *>(dec + B);  // 4 group mbarriers
    unsigned long long *cnt = full + 4;                                     // maxk (replicated counts)
    int *wred_c = reinterpret_cast<int *>(cnt + maxk);                      // 32
    int *ctl = wred_c + 32;                                                 // [0] n_commit [1] exact flag [2] kc
    int *modlist = ctl + 4;                                                 // B
    double *cent_s = reinterpret_cast<double *>(
        (reinterpret_cast<uintptr_t>(modlist + B) + 15) & ~(uintptr_t)15);  // slots * cp, 16 B aligned

    auto cptr = [&](int slot) -> double * {
        return A.cent_in_smem ? cent_s + (size_t)slot * cp : A.centroids + ((size_t)slot * ncta + rank) * f;
    };
    auto rowptr = [&](long long r) -> const double * { return ring + (size_t)(r & (R - 1)) * fpad; };

    // ---- init
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = tid; c < maxk; c += blockDim.x) {
        cnt[c] = (c < A.init_k) ? A.sizes[c] : 0ull;
        disp[c] = 0.0;
    }
    int kc = A.init_k;
    for (int s = warp; s < slots; s += nw) {  // resume: adopt the state left by the previous shard
        const int c = s * ncta + rank;
        if (c >= kc) break;
        if (A.cent_in_smem) {
            double *cv = cent_s + (size_t)s * cp;
            const double *src = A.centroids + (size_t)c * f;
            for (int j = lane; j < f; j += 32) cv[j] = src[j];
        }
    }
    const double r_half = A.radius * 0.5, r_full = A.radius, r_relax = A.radius * 1.5;
    const unsigned row_bytes = (unsigned)f * 8u;
    long long next_fetch = 0;   // rows [0, next_fetch) have been requested (multiple of kGroup, or n)
    long long waited_groups = 0;  // groups [0, waited_groups) are known to have landed
    long long r0 = 0;
    int n_exact = 0;
    long long n_blocks = 0;
    __syncthreads();
    cluster.sync();

    __shared__ long long tphase[8];  // debug phase timers (thread 0 only)
    if (tid == 0)
        for (int k = 0; k < 8; ++k) tphase[k] = 0;
    long long tlast = clock64();
#define ASB_TICK(k)                                  \
    do {                                             \
        if (A.phase_times) {                         \
            if (tid == 0) {                          \
                const long long _t = clock64();      \
                tphase[k] += _t - tlast;             \
                tlast = _t;                          \
            }                                        \
            __syncwarp(); /* aligned barriers follow */ \
        }                                            \
    } while (0)
    while (r0 < A.n) {
        const int par = (int)(n_blocks & 1);
        // ---- fetch whole groups of 8 rows up to group (r0/8 + 3): the slot of that group held group
        //      r0/8 - 1, every row of which is already committed
        long long fetch_to = (r0 / kGroup + 4) * kGroup;
        if (fetch_to > A.n) fetch_to = A.n;
        if (A.vec) {
            if (tid == 0) {
                for (long long r = next_fetch; r < fetch_to; r += kGroup) {
                    const long long g = r / kGroup;
                    const long long rows_in = (A.n - r) < kGroup ? (A.n - r) : kGroup;
                    unsigned long long *bar = &full[g & 3];
                    const unsigned bytes = (unsigned)rows_in * row_bytes;
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(ring + (size_t)(r & (R - 1)) * fpad, A.rows + r * (long long)f, bytes, bar);
                }
            }
        } else {
            for (long long r = next_fetch; r < fetch_to; ++r) {
                double *dst = ring + (size_t)(r & (R - 1)) * fpad;
                const double *src = A.rows + r * (long long)f;
                for (int j = tid; j < f; j += blockDim.x) dst[j] = src[j];
            }
            __syncthreads();
        }
        next_fetch = fetch_to;
        // keep block starts aligned to the copy groups: after a short block the next one stops at a group
        // boundary, so that the following block's rows were all requested one block ahead
        int nb = B - (int)(r0 & (kGroup - 1));
        if ((long long)nb > A.n - r0) nb = (int)(A.n - r0);
        if (A.vec) {
            const long long g_last = (r0 + nb - 1) / kGroup;
            for (; waited_groups <= g_last; ++waited_groups)
                mbar_wait(&full[waited_groups & 3], (unsigned)((waited_groups >> 2) & 1));
        }

        ASB_TICK(0);  // fetch issue + wait for rows
        // ---- 1. fast distances.  Work item = (pair of this CTA's centroids) x (8 consecutive rows of the
        //      block); items go round-robin to the warps.  Rows >= nb read stale ring slots and a pair's
        //      missing second centroid is a duplicate of the first; those results are never consumed.
        const int s0 = (int)(r0 & (R - 1));
        {
            const int my_n = kc > rank ? (kc - rank + ncta - 1) / ncta : 0;  // my centroids < kc
            const int npairs = (my_n + 1) >> 1;
            const int nitems = npairs * (B / 8);
            for (int it = warp; it < nitems; it += nw) {
                const int pair = it / (B / 8), half = it % (B / 8);
                const int sl0 = 2 * pair, sl1 = (2 * pair + 1 < my_n) ? 2 * pair + 1 : 2 * pair;
                const double *cv0 = cptr(sl0), *cv1 = cptr(sl1);
                const int sh = (s0 + 8 * half) & (R - 1);
                double acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0;
                int j0 = 0;
                if (A.vec2) {
                    for (; j0 + 128 <= f; j0 += 128) dist_tile2x8_v2<2>(acc, cv0, cv1, ring, sh, R - 1, fpad, j0, lane);
                    for (; j0 + 64 <= f; j0 += 64) dist_tile2x8_v2<1>(acc, cv0, cv1, ring, sh, R - 1, fpad, j0, lane);
                }
                for (; j0 + 64 <= f; j0 += 64) dist_tile2x8<2>(acc, cv0, cv1, ring, sh, R - 1, fpad, j0, lane);
                for (; j0 + 32 <= f; j0 += 32) dist_tile2x8<1>(acc, cv0, cv1, ring, sh, R - 1, fpad, j0, lane);
                if (j0 + lane < f) {
                    const double a0 = cv0[j0 + lane], a1 = cv1[j0 + lane];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const double x = ring[(size_t)((sh + i) & (R - 1)) * fpad + j0 + lane];
                        double d;
                        d = x - a0; acc[i] = fma(d, d, acc[i]);
                        d = x - a1; acc[8 + i] = fma(d, d, acc[8 + i]);
                    }
                }
                transpose_reduce<16>(acc, lane);  // lane L now holds element L >> 1 = c * 8 + i
                double tot = acc[0];
                if (!(tot == tot)) tot = INFINITY;  // NaN never wins (`d2 < best` is false)
                if ((lane & 1) == 0) {
                    const int e = lane >> 1, cidx = e >> 3, i = e & 7;
                    if (cidx == 0 || sl1 != sl0) D[(size_t)(cidx ? sl1 : sl0) * B + 8 * half + i] = tot;
                }
            }
        }
        __syncthreads();

        ASB_TICK(1);  // phase 1 + barrier
        // ---- 2. per row: arg-min over this CTA's centroids, all-to-all through DSMEM
        const int my_valid = kc > rank ? (kc - rank + ncta - 1) / ncta : 0;  // my centroids < kc
        for (int i = warp; i < nb; i += nw) {
            double bd = INFINITY, sd = INFINITY;
            int bc = kNone;
            for (int s = lane; s < my_valid; s += 32) {
                const double d = D[(size_t)s * B + i];
                const int c = s * ncta + rank;
                if (lex_less(d, c, bd, bc)) {
                    sd = bd;
                    bd = d;
                    bc = c;
                } else if (d < sd) {
                    sd = d;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double obd = __shfl_xor_sync(0xffffffffu, bd, o);
                const double osd = __shfl_xor_sync(0xffffffffu, sd, o);
                const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (lex_less(obd, obc, bd, bc)) {
                    sd = fmin(fmin(sd, osd), bd);
                    bd = obd;
                    bc = obc;
                } else {
                    sd = fmin(fmin(sd, osd), obd);
                }
            }
            if (lane < ncta) {
                Xch *remote = cluster.map_shared_rank(xch, lane) + ((size_t)par * 16 + rank) * B + i;
                Xch v;
                v.best_d = bd;
                v.second_d = sd;
                v.best_c = bc;
                v.pad = 0;
                *remote = v;
            }
        }
        ASB_TICK(2);  // phase 2
        cluster.sync();
        ASB_TICK(3);  // cluster barrier

        // ---- 3a. per row: reduce the 16 CTA entries -> G[i]
        for (int i = warp; i < nb; i += nw) {
            const Xch *e = xch + ((size_t)par * 16) * B + i;
            double bd = lane < ncta ? e[(size_t)lane * B].best_d : INFINITY;
            double sd = lane < ncta ? e[(size_t)lane * B].second_d : INFINITY;
            int bc = lane < ncta ? e[(size_t)lane * B].best_c : kNone;
            for (int o = 8; o > 0; o >>= 1) {
                const double obd = __shfl_xor_sync(0xffffffffu, bd, o);
                const double osd = __shfl_xor_sync(0xffffffffu, sd, o);
                const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (lex_less(obd, obc, bd, bc)) {
                    sd = fmin(fmin(sd, osd), bd);
                    bd = obd;
                    bc = obc;
                } else {
                    sd = fmin(fmin(sd, osd), obd);
                }
            }
            if (lane == 0) {
                GRow g;
                g.bd = bd;
                g.sd = sd;
                g.sb = sqrt(bd);
                g.ss = sqrt(sd);
                g.bc = bc;
                g.pad = 0;
                G[i] = g;
            }
        }
        __syncthreads();

        ASB_TICK(4);  // 3a + barrier
        // ---- 3b. resolve the block's rows IN ORDER (warp 0, identical in every CTA).
        //      Lane i holds row i's summary.  Fast path (saturated state, the steady state of a long
        //      walk): all rows are certified at once against the worst-case total displacement E of
        //      the block, E >= sum_j (sb_j + E) / (cnt_j + 1); duplicates of a target get their k_new
        //      in row order from a match mask.  Otherwise: the sequential interval-certified loop.
        if (warp == 0) {
            GRow g;
            g.bd = g.sd = g.sb = g.ss = INFINITY;
            g.bc = kNone;
            if (lane < nb) g = G[lane];
            int n_commit = 0, exact = 0, kcl = kc;
            int my_action = 3, my_target = -1;
            double my_knew = 0.0;
            bool done = false;
            if (!A.force_exact && kc == maxk) {
                const int b = (g.bc == kNone) ? 0 : g.bc;
                const unsigned long long cb = cnt[b];
                const float inv = lane < nb ? 1.0f / (float)(cb + 1ull) : 0.0f;
                float qsum = inv;
                double esum = lane < nb ? g.sb * (double)inv : 0.0;
                for (int o = 16; o > 0; o >>= 1) {
                    qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
                    esum += __shfl_xor_sync(0xffffffffu, esum, o);
                }
                bool ok = (qsum < 0.25f) && (esum < INFINITY);
                const double E = esum / (1.0 - (double)qsum) * 1.001 + 1e-300;
                const double hi_b = g.sb + E, lo_b = fmax(g.sb - E, 0.0);
                const double lo2 = lo_b * lo_b * (1.0 - kDelta), hi2 = hi_b * hi_b * (1.0 + kDelta);
                bool row_ok = (g.bd < INFINITY) && ((g.ss - E) > hi_b * (1.0 + kDelta)) &&
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
                        if ((m >> lane) == 1u) cnt[b] = cb + (unsigned long long)__popc(m);  // highest lane of the group
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
                    const double sd = __shfl_sync(0xffffffffu, g.sd, i);
                    const double sb = __shfl_sync(0xffffffffu, g.sb, i);
                    const double ss = __shfl_sync(0xffffffffu, g.ss, i);
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
                        double lo2, hi2, hi_b;
                        if (dmax == 0.0) {  // nothing moved yet in this block: the snapshot is the state
                            hi_b = sb;
                            if (!(sd > bd * (1.0 + kDelta))) ok = false;
                            lo2 = bd * (1.0 - kDelta);
                            hi2 = bd * (1.0 + kDelta);
                        } else {
                            hi_b = sb + mod_b;
                            const double lo_b = fmax(sb - mod_b, 0.0);
                            if (!((ss - dmax) > hi_b * (1.0 + kDelta))) ok = false;
                            lo2 = lo_b * lo_b * (1.0 - kDelta);
                            hi2 = hi_b * hi_b * (1.0 + kDelta);
                        }
                        if ((r_half >= lo2 && r_half <= hi2) || (r_full >= lo2 && r_full <= hi2) ||
                            (r_relax >= lo2 && r_relax <= hi2))
                            ok = false;
                        if (!ok) {
                            if (i == 0) exact = 1;
                            break;
                        }
                        const double d2 = bd;  // any value of the certified interval gives the same outcome
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
                            // |c' - c| = |x - c| / k_new; a float reciprocal with 1e-3 slack bounds it from above
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
            if (lane < n_commit) {  // decision records, owner coordinates computed in parallel
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
            }
            if (lane == 0) {
                ctl[0] = n_commit;
                ctl[1] = exact;
                ctl[2] = kcl;
            }
        }
        __syncthreads();
        ASB_TICK(5);  // resolve + barrier
        int n_commit = ctl[0];
        const int exact = ctl[1];

        if (exact) {
            // ---- exact path for row r0: reference arithmetic for every candidate within delta
            n_exact++;
            const GRow g = G[0];
            const double hi = g.bd * (1.0 + kDelta);
            const double *row = rowptr(r0);
            double my_d = INFINITY;
            int my_c = kNone;
            for (int s = tid; s < my_valid; s += blockDim.x) {
                const int c = s * ncta + rank;
                if (A.force_exact || D[(size_t)s * B] <= hi || !(hi < INFINITY)) {
                    const double *cv = cptr(s);
                    double d2 = 0.0;
                    for (int j = 0; j < f; ++j) {  // src/clustering.rs:917-921
                        const double diff = __dsub_rn(row[j], cv[j]);
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
                ctl[0] = 1;
                ctl[2] = kcl;
            }
            __syncthreads();
            n_commit = 1;
        }
        kc = ctl[2];

        // ---- 4. apply the committed decisions in row order (owner warps), write assignments
        {
            bool mine = false;
            if (lane < n_commit) {
                const Dec dd = dec[lane];
                mine = dd.action != 3 && dd.owner == rank && dd.owarp == warp;
            }
            unsigned todo = __ballot_sync(0xffffffffu, mine);
            while (todo) {
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                const Dec dd = dec[i];
                double *cv = cptr(dd.slot);
                const double *row = rowptr(r0 + i);
                if (dd.action == 0) {
                    for (int j = lane; j < f; j += 32) cv[j] = row[j];
                } else if (dd.action == 1) {
                    for (int j = lane; j < f; j += 32) {
                        const double c0 = cv[j];
                        cv[j] = __dadd_rn(c0, __ddiv_rn(__dsub_rn(row[j], c0), dd.knew));  // :748
                    }
                }
                __syncwarp();
            }
        }
        if (rank == 0 && tid < n_commit) A.assign[r0 + tid] = (long long)dec[tid].target;
        r0 += n_commit;
        n_blocks++;
        __syncthreads();  // ring slots, D, G, dec are free for the next block
        ASB_TICK(6);  // exact path (if any) + apply + barrier
    }
    cluster.sync();
    // ---- write back
    for (int s = warp; s < slots; s += nw) {
        const int c = s * ncta + rank;
        if (c >= kc) break;
        if (A.cent_in_smem) {
            const double *cv = cent_s + (size_t)s * cp;
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

template <int B>
size_t cluster_block_smem_bytes(int f, int slots, int maxk, bool cent_in_smem) {
    const int fpad = (f + 1) & ~1;
    size_t b = (size_t)4 * kGroup * fpad * 8;       // ring (R rows, independent of B)
    b += (size_t)slots * B * 8;                     // D
    b += (size_t)((maxk + 1) & ~1) * 8;             // disp
    b += 32 * 8;                                    // wred_d
    b += (size_t)(2 * 16 * B + 32) * sizeof(Xch);   // xch + xch_exact
    b += (size_t)B * sizeof(GRow);
    b += (size_t)B * sizeof(Dec);
    b += (size_t)4 * 8;                             // group mbarriers
    b += (size_t)maxk * 8;                          // cnt
    b += 32 * 4 + 4 * 4 + (size_t)(B + 2) * 4;      // wred_c, ctl, modlist
    if (cent_in_smem) b += (size_t)slots * block_cent_pitch(f) * 8;
    return b + 96;
}

}  // namespace
