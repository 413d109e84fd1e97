// cluster.cu -- K2: incremental leader clustering in row order, bit-exact with the reference's
// deterministic branch.
//
// Replaces run_incremental_clustering_with_sampling + nearest_centroid
// (src/clustering.rs:547-928) for `.with_seed(s).with_inline_sampling(None)` (:842-843).
//
// The algorithm is order dependent (row r sees every centroid move made by rows < r), so the
// rows are walked in order by ONE thread-block cluster of up to 16 CTAs that keeps the whole
// K x F centroid state on chip, distributed over the cluster's shared memory (centroid c lives
// in CTA c mod 16).  Per row:
//   1. every warp computes the squared distance from the row to the centroid(s) it owns with a
//      lane-split FMA reduction (~300 cycles instead of the 8*F-cycle dependent chain);
//   2. CTA-local arg-min, then a 16 x 16 all-to-all of (best, second best) through distributed
//      shared memory and ONE cluster barrier per row (parity double buffering);
//   3. every thread derives the same decision; the owner warp applies it (new centroid /
//      running mean / count only), element-wise with the reference's own IEEE operations.
// The fast distances differ from the reference's sequential non-fused sum by at most
// ~2(F+4)*2^-53 relative, so a decision is taken from them only when it is CERTIFIED: the
// runner-up is more than delta = 1e-11 (relative) away and the winner is not within delta of any
// of the three thresholds (radius/2, radius, 1.5 radius).  Otherwise the row goes through the
// exact path: every candidate within delta of the winner is recomputed by one thread with the
// reference's sequential arithmetic (__dmul_rn/__dadd_rn, strict left-to-right, first strict
// minimum wins) and the decision is taken from those values.  Outputs are therefore identical to
// the reference's for every input; the exact path is also selectable for all rows (tests).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kRing = 8;          // rows prefetched ahead
constexpr double kDelta = 1e-11;  // certification margin (>> 2(F+4)2^-53 for F <= 16384)

struct ClusterArgs {
    const double *rows;
    long long n;
    int f;
    int max_k;
    int init_k;  // centroids already present in `centroids`/`sizes` (pipeline hand-off)
    double radius;
    double *centroids;  // [max_k][f] global: output, and live storage when !cent_in_smem
    long long *assign;
    unsigned long long *sizes;
    int *x_out;
    int *stats;  // [0] rows through the exact path, [1] blocks (blocked kernel)
    int cent_in_smem;
    int slots_per_cta;
    int force_exact;
    int vec;   // rows 16B-copyable
    const float *rows32;                  // FP32 copy of the rows (cluster_f32_kernel)
    const unsigned long long *max_norm2_bits;  // bit pattern of max finite |x|^2 over rows and initial centroids
    long long *phase_times;  // optional: 8 cycle counters of CTA 0 / thread 0 (debug option)
    int vec2;  // centroid storage 16B-loadable (blocked kernel, LDS.128 distance loop)
    const double *rows_n2;  // |fl32(row)|^2 in FP64 (pipelined kernel: distances via dot products)
    int tick_tid;           // thread of CTA 0 that owns the debug phase timers
    int tile_check;         // debug: compare every tensor-core distance with FP64 (pipelined kernel, no speculation)
    int ring_groups;        // pipelined kernel: 8-row groups in the shared-memory row ring (8, 6 or 4)
};

struct __align__(16) Xch {
    double best_d;
    double second_d;
    int best_c;
    int pad;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// lexicographic (d, c) minimum; d is never NaN here
__device__ __forceinline__ bool lex_less(double d1, int c1, double d2, int c2) {
    return d1 < d2 || (d1 == d2 && c1 < c2);
}

}  // namespace

#include "cluster_block.cuh"
#include "cluster_f32.cuh"
#include "cluster_f32p.cuh"

namespace {

__global__ void __launch_bounds__(1024, 1) cluster_rowwise_kernel(ClusterArgs A) {
    cg::cluster_group cluster = cg::this_cluster();
    const int ncta = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int f = A.f;
    const int cp = f | 1;  // odd pitch: conflict-free thread-per-centroid chains
    const int slots = A.slots_per_cta;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);               // kRing * fpad
    const int fpad = (f + 1) & ~1;
    double *dfast = ring + (size_t)kRing * fpad;                       // slots
    double *dexact = dfast + slots;                                    // slots
    double *wbest_d = dexact + slots;                                  // 32
    double *wsec_d = wbest_d + 32;                                     // 32
    Xch *xch = reinterpret_cast<Xch *>(wsec_d + 32);                   // 2 parities x 2 phases x 16
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(xch + 64);  // slots
    int *wbest_c = reinterpret_cast<int *>(cnt + slots);               // 32 (+pad)
    double *cent_s = reinterpret_cast<double *>(wbest_c + 32);         // slots * cp (if in smem)

    auto cptr = [&](int slot) -> double * {
        return A.cent_in_smem ? cent_s + (size_t)slot * cp : A.centroids + ((size_t)slot * ncta + rank) * f;
    };
    auto issue_row = [&](long long r) {
        if (r < A.n) {
            double *dst = ring + (size_t)(r % kRing) * fpad;
            const double *src = A.rows + r * (long long)f;
            if (A.vec) {
                for (int c = tid; c < f / 2; c += blockDim.x) cp_async16(dst + 2 * c, src + 2 * c);
            } else {
                for (int c = tid; c < f; c += blockDim.x) cp_async8(dst + c, src + c);
            }
        }
        cp_async_commit();
    };

    for (int s = tid; s < slots; s += blockDim.x) cnt[s] = 0ull;
    for (long long r = 0; r < kRing - 1; ++r) issue_row(r);
    int kc = A.init_k;
    __syncthreads();
    for (int s = warp; s < slots; s += nw) {  // resume: adopt the state left by the previous shard
        const int c = s * ncta + rank;
        if (c >= kc) break;
        if (A.cent_in_smem) {
            double *cv = cent_s + (size_t)s * cp;
            const double *src = A.centroids + (size_t)c * f;
            for (int j = lane; j < f; j += 32) cv[j] = src[j];
        }
        if (lane == 0) cnt[s] = A.sizes[c];
    }
    int n_exact = 0;
    const double r_half = A.radius * 0.5, r_full = A.radius, r_relax = A.radius * 1.5;
    __syncthreads();
    cluster.sync();

    for (long long r = 0; r < A.n; ++r) {
        const int par = (int)(r & 1);
        cp_async_wait<kRing - 2>();
        __syncthreads();               // row r visible to all; everyone finished row r-1 entirely
        issue_row(r + kRing - 1);      // reuses the slot of row r-1
        const double *row = ring + (size_t)(r % kRing) * fpad;

        // ---- 1. fast distances to the centroids this warp owns
        double wb_d = INFINITY, ws_d = INFINITY;
        int wb_c = 0x7fffffff;
        for (int s = warp; s < slots; s += nw) {
            const int c = s * ncta + rank;
            if (c >= kc) break;
            const double *cv = cptr(s);
            double acc = 0.0;
            for (int j = lane; j < f; j += 32) {
                const double df = row[j] - cv[j];
                acc = fma(df, df, acc);
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (!(acc == acc)) acc = INFINITY;  // NaN never wins (`d2 < best` is false)
            if (lane == 0) dfast[s] = acc;
            if (lex_less(acc, c, wb_d, wb_c)) {
                ws_d = wb_d;
                wb_d = acc;
                wb_c = c;
            } else if (acc < ws_d) {
                ws_d = acc;
            }
        }
        if (lane == 0) {
            wbest_d[warp] = wb_d;
            wsec_d[warp] = ws_d;
            wbest_c[warp] = wb_c;
        }
        __syncthreads();
        // ---- 2. CTA arg-min (warp 0) and all-to-all through distributed shared memory
        if (warp == 0) {
            double bd = lane < nw ? wbest_d[lane] : INFINITY;
            double sd = lane < nw ? wsec_d[lane] : INFINITY;
            int bc = lane < nw ? wbest_c[lane] : 0x7fffffff;
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
                Xch *remote = cluster.map_shared_rank(xch, lane) + (par * 2 + 0) * 16 + rank;
                Xch v;
                v.best_d = bd;
                v.second_d = sd;
                v.best_c = bc;
                v.pad = 0;
                *remote = v;
            }
        }
        cluster.sync();
        // ---- 3. every warp reduces the 16 entries -> identical global (best, second)
        double gb_d, gs_d;
        int gb_c;
        {
            const Xch *e = xch + (par * 2 + 0) * 16;
            double bd = lane < ncta ? e[lane].best_d : INFINITY;
            double sd = lane < ncta ? e[lane].second_d : INFINITY;
            int bc = lane < ncta ? e[lane].best_c : 0x7fffffff;
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
            gb_d = bd;
            gs_d = sd;
            gb_c = bc;
        }
        double dec_d = gb_d;
        int dec_c = (gb_c == 0x7fffffff) ? 0 : gb_c;
        if (kc > 0) {
            bool ambiguous = A.force_exact != 0;
            const double hi = gb_d * (1.0 + kDelta);
            if (!(gs_d > hi)) ambiguous = true;  // runner-up too close (also catches inf/inf)
            const double thr[3] = {r_half, r_full, r_relax};
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const double lo_t = thr[t] * (1.0 - kDelta), hi_t = thr[t] * (1.0 + kDelta);
                if (gb_d >= lo_t && gb_d <= hi_t) ambiguous = true;
            }
            if (ambiguous) {
                // ---- exact path: reference arithmetic for every candidate within delta
                n_exact++;
                double my_d = INFINITY;
                int my_c = 0x7fffffff;
                for (int s = tid; s < slots; s += blockDim.x) {
                    const int c = s * ncta + rank;
                    if (c < kc && (dfast[s] <= hi || !(hi < INFINITY))) {
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
                __syncthreads();  // wbest_* of step 1 fully consumed by warp 0
                if (lane == 0) {
                    wbest_d[warp] = my_d;
                    wbest_c[warp] = my_c;
                }
                __syncthreads();
                if (warp == 0) {
                    double bd = lane < nw ? wbest_d[lane] : INFINITY;
                    int bc = lane < nw ? wbest_c[lane] : 0x7fffffff;
                    for (int o = 16; o > 0; o >>= 1) {
                        const double obd = __shfl_xor_sync(0xffffffffu, bd, o);
                        const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                        if (lex_less(obd, obc, bd, bc)) {
                            bd = obd;
                            bc = obc;
                        }
                    }
                    if (lane < ncta) {
                        Xch *remote = cluster.map_shared_rank(xch, lane) + (par * 2 + 1) * 16 + rank;
                        Xch v;
                        v.best_d = bd;
                        v.second_d = INFINITY;
                        v.best_c = bc;
                        v.pad = 0;
                        *remote = v;
                    }
                }
                cluster.sync();
                const Xch *e = xch + (par * 2 + 1) * 16;
                double bd = lane < ncta ? e[lane].best_d : INFINITY;
                int bc = lane < ncta ? e[lane].best_c : 0x7fffffff;
                for (int o = 16; o > 0; o >>= 1) {
                    const double obd = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int obc = __shfl_xor_sync(0xffffffffu, bc, o);
                    if (lex_less(obd, obc, bd, bc)) {
                        bd = obd;
                        bc = obc;
                    }
                }
                dec_d = bd;
                dec_c = (bc == 0x7fffffff) ? 0 : bc;
            }
        }
        // ---- 4. decision (src/clustering.rs:637-815), identical in every thread of the cluster
        int action;  // 0 new centroid, 1 running-mean update, 2 count only, 3 drop
        int target;
        if (kc == 0) {
            action = 0;
            target = 0;
        } else if (kc < A.max_k && dec_d > r_half) {
            action = 0;
            target = kc;
        } else if (dec_d <= r_full) {
            action = 1;
            target = dec_c;
        } else if (dec_d <= r_relax) {
            action = 2;
            target = dec_c;
        } else {
            action = 3;
            target = -1;
        }
        if (action != 3) {
            const int owner = target % ncta, slot = target / ncta;
            if (owner == rank && warp == slot % nw) {
                double *cv = cptr(slot);
                if (action == 0) {
                    for (int j = lane; j < f; j += 32) cv[j] = row[j];
                    if (lane == 0) cnt[slot] = 1ull;
                } else if (action == 1) {
                    const double k_new = (double)cnt[slot] + 1.0;  // :736-737
                    for (int j = lane; j < f; j += 32) {
                        const double c0 = cv[j];
                        cv[j] = __dadd_rn(c0, __ddiv_rn(__dsub_rn(row[j], c0), k_new));  // :748
                    }
                    __syncwarp();
                    if (lane == 0) cnt[slot] += 1ull;
                } else {
                    if (lane == 0) cnt[slot] += 1ull;  // :780
                }
            }
        }
        if (action == 0) kc++;
        if (rank == 0 && tid == 0) A.assign[r] = (long long)target;
    }
    cp_async_wait<0>();
    __syncthreads();
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
        if (lane == 0) A.sizes[c] = cnt[s];
    }
    if (rank == 0 && tid == 0) {
        A.x_out[0] = kc;
        A.stats[0] = n_exact;
    }
}

size_t cluster_smem_bytes(int f, int slots, bool cent_in_smem) {
    const int fpad = (f + 1) & ~1;
    size_t b = (size_t)kRing * fpad * 8;
    b += (size_t)slots * 8 * 2;  // dfast, dexact
    b += 64 * 8;                 // wbest_d, wsec_d
    b += 64 * sizeof(Xch);
    b += (size_t)slots * 8;      // cnt
    b += 32 * 4;                 // wbest_c
    if (cent_in_smem) b += (size_t)slots * (f | 1) * 8;
    return b + 32;
}

}  // namespace

int asb_dev_cluster_seq(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int64_t max_clusters, double radius,
                    double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d, int64_t *x_out_host,
                    int64_t init_k) {
    if (n <= 0 || f <= 0 || max_clusters <= 0)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster: n=%lld f=%lld max_clusters=%lld", (long long)n, (long long)f,
                 (long long)max_clusters);
    if (f > 16384 || max_clusters > (1 << 20)) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "cluster: f or max_clusters too large");
    if (init_k < 0 || init_k > max_clusters) ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster: init_k=%lld", (long long)init_k);
    asb_wait_rows(ctx, rows_d + n * f);   // (a build from host memory may still be uploading the tail of the matrix)
    const size_t smem_cap = 227 * 1024;
    ClusterArgs A{};
    A.rows = rows_d;
    A.n = n;
    A.f = (int)f;
    A.max_k = (int)max_clusters;
    A.init_k = (int)init_k;
    A.radius = radius;
    A.centroids = centroids_d;
    A.assign = (long long *)assign_d;
    A.sizes = sizes_d;
    {
        auto it = ctx->options.find("cluster_force_exact");
        A.force_exact = (it != ctx->options.end() && it->second != 0.0) ? 1 : 0;
    }
    bool want_rowwise = false;
    {
        auto it = ctx->options.find("cluster_rowwise");
        want_rowwise = (it != ctx->options.end() && it->second != 0.0);
    }
    A.vec = (f % 2 == 0) && (((uintptr_t)rows_d & 15) == 0);
    DevTmp<int> scratch;
    ASB_TRY(scratch.init(ctx, 4));
    ASB_CUDA(ctx, cudaMemsetAsync(scratch.ptr, 0, 4 * sizeof(int), ctx->stream));
    A.x_out = scratch.ptr;
    A.stats = scratch.ptr + 1;
    DevTmp<long long> ptimes;
    bool want_times = false;
    {
        auto it = ctx->options.find("cluster_phase_times");
        want_times = (it != ctx->options.end() && it->second != 0.0);
    }
    {
        auto it = ctx->options.find("cluster_tick_tid");
        A.tick_tid = it != ctx->options.end() ? (int)it->second : 0;
    }
    {
        auto it = ctx->options.find("cluster_check_tile");
        A.tile_check = (it != ctx->options.end() && it->second != 0.0) ? 1 : 0;
        if (A.tile_check) want_times = true;  // the result travels in the debug counters
    }
    if (want_times) {
        ASB_TRY(ptimes.init(ctx, 48));
        ASB_CUDA(ctx, cudaMemsetAsync(ptimes.ptr, 0, 48 * sizeof(long long), ctx->stream));
        A.phase_times = ptimes.ptr;
    }

    // variant -2: pipelined FP32-prefilter kernel (resolve of block b overlaps the distances of b+1);
    // -1: FP32-prefilter kernel, 32 rows per cluster barrier; 0/1: FP64 blocked kernel with B = 16 / 8;
    // 2: row-wise kernel
    bool allow_f32 = (f % 4 == 0) && (((uintptr_t)rows_d & 15) == 0) && !want_rowwise;
    {
        auto it = ctx->options.find("cluster_no_f32");
        if (it != ctx->options.end() && it->second != 0.0) allow_f32 = false;
    }
    DevTmp<float> rows32;
    DevTmp<double> rows_n2;
    DevTmp<unsigned long long> maxn2;
    int rows32_pitch = -1;
    int launched = 0, variant_used = -9;
    bool allow_pipe = !want_rowwise && (((uintptr_t)rows_d & 15) == 0);
    {
        auto it = ctx->options.find("cluster_no_f32");
        if (it != ctx->options.end() && it->second != 0.0) allow_pipe = false;
    }
    int first_variant = want_rowwise ? 2 : (allow_pipe ? -2 : (allow_f32 ? -1 : 0));
    {
        auto it = ctx->options.find("cluster_no_pipeline");
        if (first_variant == -2 && it != ctx->options.end() && it->second != 0.0) first_variant = -1;
    }
    {
        auto it = ctx->options.find("cluster_first_variant");  // diagnostics/tests: skip the faster variants
        if (it != ctx->options.end() && (int)it->second > first_variant && (int)it->second <= 2) first_variant = (int)it->second;
    }
    for (int variant = first_variant; variant < 3 && !launched; ++variant) {
        for (int ncta : {16, 8, 4, 2, 1}) {
            const int slots = (int)((max_clusters + ncta - 1) / ncta);
            int ring_groups = 8;  // pipelined kernel: shrink the row ring (less prefetch) before giving the variant up
            auto bytes = [&](bool in_smem) -> size_t {
                if (variant == -2) return cluster_f32p_smem_bytes((int)f, slots, (int)max_clusters, ring_groups);
                if (variant == -1) return cluster_f32_smem_bytes((int)f, slots, (int)max_clusters, in_smem);
                if (variant == 0) return cluster_block_smem_bytes<16>((int)f, slots, (int)max_clusters, in_smem);
                if (variant == 1) return cluster_block_smem_bytes<8>((int)f, slots, (int)max_clusters, in_smem);
                return cluster_smem_bytes((int)f, slots, in_smem);
            };
            if (variant == -2)
                while (ring_groups > 4 && bytes(false) > smem_cap) ring_groups -= 2;
            A.ring_groups = ring_groups;
            const bool in_smem = variant != -2 && bytes(true) <= smem_cap;
            const size_t smem = bytes(in_smem);
            if (smem > smem_cap) continue;
            if (variant == -1 && !allow_f32) continue;
            const int pitch32 = variant == -2 ? f32p_pitch((int)f) : (int)f;
            if (variant < 0 && rows32_pitch != pitch32) {  // one streaming pass: FP32 copy of the rows + max |x|^2
                ASB_TRY(rows32.init(ctx, (size_t)n * pitch32));
                if (!rows_n2.ptr) ASB_TRY(rows_n2.init(ctx, (size_t)n));
                if (!maxn2.ptr) ASB_TRY(maxn2.init(ctx, 1));
                ASB_CUDA(ctx, cudaMemsetAsync(maxn2.ptr, 0, sizeof(unsigned long long), ctx->stream));
                {
                    KernelTimer kt(ctx, "cluster_prep_kernel");
                    rows_to_f32_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(
                        rows_d, (long long)n, (int)f, rows32.ptr, maxn2.ptr, rows_n2.ptr, pitch32);
                }
                ASB_TRY(asb_check_launch(ctx, "rows_to_f32_kernel"));
                if (init_k > 0) {
                    rows_to_f32_kernel<<<(unsigned)((init_k + 7) / 8), 256, 0, ctx->stream>>>(
                        centroids_d, (long long)init_k, (int)f, nullptr, maxn2.ptr, nullptr, (int)f);
                    ASB_TRY(asb_check_launch(ctx, "rows_to_f32_kernel(centroids)"));
                }
                rows32_pitch = pitch32;
                A.rows32 = rows32.ptr;
                A.rows_n2 = rows_n2.ptr;
                A.max_norm2_bits = maxn2.ptr;
            }
            const int wcap = variant < 2 ? 24 : 32;  // launch bounds of the variants
            const int nwarps = variant == -2 ? 16 : (slots < 4 ? 4 : (slots > wcap ? wcap : slots));
            A.cent_in_smem = in_smem ? 1 : 0;
            A.slots_per_cta = slots;
            A.vec2 = (A.vec && (in_smem || (((uintptr_t)centroids_d & 15) == 0))) ? 1 : 0;
            const void *fn = variant == -2  ? (const void *)cluster_f32p_kernel
                             : variant == -1  ? (in_smem ? (const void *)cluster_f32_kernel<true> : (const void *)cluster_f32_kernel<false>)
                             : variant == 0 ? (const void *)cluster_block_kernel<16>
                             : variant == 1 ? (const void *)cluster_block_kernel<8>
                                            : (const void *)cluster_rowwise_kernel;
            if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ncta);
            cfg.blockDim = dim3(nwarps * 32);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = ncta;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int max_clusters_active = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters_active, fn, &cfg);
            if (e != cudaSuccess || max_clusters_active < 1) {
                cudaGetLastError();
                continue;
            }
            void *args[] = {(void *)&A};
            {
                KernelTimer kt(ctx, "cluster_kernel");
                e = cudaLaunchKernelExC(&cfg, fn, args);
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            launched = ncta;
            variant_used = variant;
            break;
        }
    }
    if (!launched) ASB_FAIL(ctx, ASB_ERR_CUDA, "cluster: no cluster configuration could be launched");
    ctx->launches++;
    int h[3] = {0, 0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(h, scratch.ptr, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *x_out_host = h[0];
    ctx->kernel_ms["cluster_exact_rows"] = (double)h[1];
    if (want_times) {
        long long ht[48];
        ASB_CUDA(ctx, cudaMemcpyAsync(ht, ptimes.ptr, sizeof(ht), cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < 48; ++k) ctx->kernel_ms[std::string("cluster_phase") + std::to_string(k)] = (double)ht[k];
    }
    ctx->kernel_ms["cluster_ncta"] = (double)launched;
    ctx->kernel_ms["cluster_ring_groups"] = variant_used == -2 ? (double)A.ring_groups : 0.0;
    ctx->kernel_ms["cluster_blocks"] = (double)h[2];
    ctx->kernel_ms["cluster_variant"] = (double)variant_used;
    if (h[0] == 0) ASB_FAIL(ctx, ASB_ERR_NO_CLUSTERS, "No clusters created from data");  // clustering.rs:869-874
    return ASB_OK;
}
