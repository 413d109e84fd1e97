// search.cu -- K8 (+K1): dense query x item contraction on the FP64 tensor pipe (DMMA m8n8k4)
// with the blended lambda-aware score and a running top-k fused into the epilogue, so the
// Q x N score matrix never exists in HBM.
//
// Replaces ArrowSpace::search_lambda_aware (src/core.rs:760-798) with the scoring of
// ArrowItem::{cosine_similarity, lambda_component_similarity, lambda_similarity}
// (src/core.rs:135-239) for a batch of queries, and -- in L2 mode -- the distance pass of
// estimate_intrinsic_dimension (src/clustering.rs:118-145).
//
// Tiling: a CTA owns 128 queries x one slab of items and walks the slab in 128-item tiles.
// Per tile the F dimension is streamed in 16-feature chunks (cp.async, double buffered) into
// row-major smem tiles with pitch 20 doubles (conflict-free fragment loads); 16 warps (4 x 4)
// each hold a 32 x 32 accumulator block as 4 x 4 DMMA fragments (4 warps per scheduler keep the
// FP64 tensor pipe issuing: with 2 per scheduler it idled half the time, profiles/r01_v1_search).  The epilogue turns dots
// into scores, parks them in smem (aliasing the operand stages) and one warp per query
// merges the survivors into that query's sorted top-k list (smem, persistent over the slab).
// Per-slab lists are merged by (score desc, index asc): exactly the order of the reference's
// stable sort (src/core.rs:785), ties -> lower index.
#include "common.cuh"

namespace {

constexpr int TQ = 128, TN = 128;
constexpr int kWarps = 16;               // 4 (queries) x 4 (items) warps, 32 x 32 accumulators each
constexpr int kThreads = kWarps * 32;
constexpr int SPITCH = 72;                        // score half-tile pitch (doubles)
constexpr int MODE_COSINE = 0, MODE_L2 = 1, MODE_ENERGY = 2;
constexpr int STATUS_NAN = 1, STATUS_ZERO_LAMBDA = 2;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct SearchArgs {
    const double *items;     // n x f
    const double *lambdas;   // n          (cosine mode)
    const double *norms2;    // n
    const double *queries;   // nq x f
    const double *lambda_q;  // nq         (cosine mode)
    const double *qnorms2;   // nq
    const long long *self_idx;  // nq      (L2 mode: item index to exclude)
    long long n, nq;
    int f, k;
    double alpha;            // cosine mode: alpha; energy mode: w_lambda
    double w_dir;            // energy mode: w_dirichlet
    int nslabs;
    long long tiles_per_slab;
    double *part_score;  // nslabs x nq x k
    int *part_idx;       // nslabs x nq x k
    int *status;
};

// Load one KC-wide chunk of `rows` rows starting at row0 into a smem tile [rows][PITCH].
template <bool VEC, int KC>
__device__ __forceinline__ void load_tile_chunk(double *dst, const double *__restrict__ src, long long row0,
                                                long long nrows_total, int f, int k0, int tid) {
    constexpr int PITCH = KC + 4;
    if (VEC) {
        // 128 rows x 8 chunks of 16 B
        for (int c = tid; c < 128 * (KC / 2); c += kThreads) {
            const int r = c / (KC / 2), ch = c % (KC / 2);
            const long long row = row0 + r;
            const int col = k0 + ch * 2;
            const bool ok = row < nrows_total && col < f;
            const double *g = ok ? src + row * (long long)f + col : src;
            cp_async16(dst + r * PITCH + ch * 2, g, ok ? 16 : 0);
        }
    } else {
        for (int c = tid; c < 128 * KC; c += kThreads) {
            const int r = c / KC, ch = c % KC;
            const long long row = row0 + r;
            const int col = k0 + ch;
            const bool ok = row < nrows_total && col < f;
            const double *g = ok ? src + row * (long long)f + col : src;
            cp_async8(dst + r * PITCH + ch, g, ok ? 8 : 0);
        }
    }
}

template <int MODE, bool VEC, int KC>
__global__ void __launch_bounds__(kThreads, 1) search_kernel(SearchArgs A) {
    constexpr int PITCH = KC + 4;                     // pitch = 4 (mod 16) doubles: conflict-free fragments
    constexpr int STAGE_DOUBLES = (TQ + TN) * PITCH;  // one pipeline stage (both operands)
    constexpr int CPR = KC / 2;                       // 16 B chunks per row
    constexpr int RPP = kThreads / CPR;               // rows covered per pass of the copy threads
    constexpr int NPT = TQ / RPP;                     // chunks per thread per operand
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);     // 2 * STAGE_DOUBLES
    double *S = stages;                                        // aliases stages: TQ x SPITCH
    double *sm_nq = stages + 2 * STAGE_DOUBLES;                // TQ
    double *sm_lq = sm_nq + TQ;                                // TQ
    double *sm_nx = sm_lq + TQ;                                // TN
    double *sm_lx = sm_nx + TN;                                // TN
    double *list_s = sm_lx + TN;                               // TQ * k
    int *list_i = reinterpret_cast<int *>(list_s + (size_t)TQ * A.k);  // TQ * k
    int *list_len = list_i + (size_t)TQ * A.k;                 // TQ
    long long *sm_self = reinterpret_cast<long long *>(list_len + TQ);  // TQ (8B aligned: TQ*k ints + TQ ints even)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int gid = lane >> 2, tig = lane & 3;
    const int k = A.k, f = A.f;

    const long long q0 = (long long)blockIdx.x * TQ;
    const int slab = blockIdx.y;
    const long long ntiles_total = (A.n + TN - 1) / TN;
    const long long t_begin = (long long)slab * A.tiles_per_slab;
    long long t_end = t_begin + A.tiles_per_slab;
    if (t_end > ntiles_total) t_end = ntiles_total;

    for (int q = tid; q < TQ; q += kThreads) {
        const long long gq = q0 + q;
        const bool ok = gq < A.nq;
        const double n2 = ok ? A.qnorms2[gq] : 0.0;
        sm_nq[q] = (MODE == MODE_COSINE) ? sqrt(n2) : n2;
        double lq = 0.0;
        if (MODE == MODE_COSINE) {
            lq = ok ? A.lambda_q[gq] : 1.0;
            if (ok && slab == 0 && lq == 0.0) atomicOr(A.status, STATUS_ZERO_LAMBDA);  // core.rs:773-776
        }
        if (MODE == MODE_ENERGY) lq = ok ? A.lambda_q[gq] : 0.0;
        sm_lq[q] = lq;
        list_len[q] = 0;
        if (MODE == MODE_L2) sm_self[q] = ok ? A.self_idx[gq] : -1;
    }
    const int nchunks = (f + KC - 1) / KC;
    const int lr = tid / CPR, lch = (tid % CPR) * 2;
    const double *gq[NPT];
    bool okq[NPT];
#pragma unroll
    for (int m = 0; m < NPT; ++m) {
        const long long row = q0 + lr + RPP * m;
        okq[m] = row < A.nq;
        gq[m] = A.queries + (okq[m] ? row : 0) * (long long)f + lch;
    }

    for (long long t = t_begin; t < t_end; ++t) {
        const long long i0 = t * TN;
        __syncthreads();  // S (aliasing the stages) and sm_nx/sm_lx of the previous tile are free
        for (int c = tid; c < TN; c += kThreads) {
            const long long gi = i0 + c;
            const bool ok = gi < A.n;
            const double n2 = ok ? A.norms2[gi] : 0.0;
            sm_nx[c] = (MODE == MODE_COSINE) ? sqrt(n2) : n2;
            sm_lx[c] = (MODE != MODE_L2 && ok) ? A.lambdas[gi] : 0.0;
        }
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // Per-thread copy descriptors (VEC path): thread t moves the 16 B chunk (row t/8 + 64 m, cols 2 (t%8))
        // of each operand, m = 0, 1; only the feature offset changes from chunk to chunk.
        const double *gx[NPT];
        bool okx[NPT];
#pragma unroll
        for (int m = 0; m < NPT; ++m) {
            const long long row = i0 + lr + RPP * m;
            okx[m] = row < A.n;
            gx[m] = A.items + (okx[m] ? row : 0) * (long long)f + lch;
        }
        auto issue_chunk = [&](int chunk) {
            double *st = stages + (chunk & 1) * STAGE_DOUBLES;
            const int k0 = chunk * KC;
            if (VEC) {
                const bool colok = k0 + lch < f;
#pragma unroll
                for (int m = 0; m < NPT; ++m) {
                    cp_async16(st + (lr + RPP * m) * PITCH + lch, gq[m] + k0, (okq[m] && colok) ? 16 : 0);
                    cp_async16(st + TQ * PITCH + (lr + RPP * m) * PITCH + lch, gx[m] + k0, (okx[m] && colok) ? 16 : 0);
                }
            } else {
                load_tile_chunk<false, KC>(st, A.queries, q0, A.nq, f, k0, tid);
                load_tile_chunk<false, KC>(st + TQ * PITCH, A.items, i0, A.n, f, k0, tid);
            }
            cp_async_commit();
        };
        issue_chunk(0);
        for (int c = 0; c < nchunks; ++c) {
            if (c + 1 < nchunks) {
                issue_chunk(c + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const double *Qs = stages + (c & 1) * STAGE_DOUBLES;
            const double *Xs = Qs + TQ * PITCH;
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Qs[(wm * 32 + i * 8 + gid) * PITCH + kk + tig];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Xs[(wn * 32 + j * 8 + gid) * PITCH + kk + tig];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncthreads();
        }

        // ---- epilogue: scores -> smem half tile -> per-query top-k merge
        for (int half = 0; half < 2; ++half) {
            if ((wn >> 1) == half) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = wm * 32 + i * 8 + gid;
                    const double nq = sm_nq[r], lq = sm_lq[r];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double2 sv;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int cl = (wn & 1) * 32 + j * 8 + tig * 2 + u;  // column inside this half
                            const int c = half * 64 + cl;
                            const long long gi = i0 + c;
                            const double dot = acc[i][j][u];
                            double s;
                            if (MODE == MODE_COSINE) {
                                const double denom = nq * sm_nx[c];                 // core.rs:230
                                const double cosv = denom > 0.0 ? dot / denom : 0.0;  // :231-236
                                const double ld = fabs(lq - sm_lx[c]);               // :136
                                const double lam = 1.0 - fmin(ld, 1.0);              // :137
                                s = A.alpha * cosv + (1.0 - A.alpha) * lam;          // :165
                                if (gi < A.n && q0 + r < A.nq && s != s) atomicOr(A.status, STATUS_NAN);
                            } else if (MODE == MODE_ENERGY) {
                                // src/energymaps.rs:838-895 (score) and :368-407 (search_energy: rank by -score)
                                const double d = sqrt(fmax(nq + sm_nx[c] - 2.0 * dot, 0.0));   // |q - x|
                                const double e = fmin(d / (1.0 + d), 1.0);                      // bounded_l2_energy
                                s = -(A.alpha * fabs(lq - sm_lx[c]) + A.w_dir * e);
                                if (gi < A.n && q0 + r < A.nq && s != s) atomicOr(A.status, STATUS_NAN);
                            } else {
                                s = -(nq + sm_nx[c] - 2.0 * dot);  // -(|q|^2 + |x|^2 - 2 q.x)
                                if (s != s) s = -INFINITY;
                                if (gi == sm_self[r]) s = -INFINITY;  // j != i, clustering.rs:123
                            }
                            if (gi >= A.n) s = -INFINITY;
                            if (u == 0) sv.x = s; else sv.y = s;
                        }
                        *reinterpret_cast<double2 *>(&S[r * SPITCH + (wn & 1) * 32 + j * 8 + tig * 2]) = sv;
                    }
                }
            }
            __syncthreads();
            for (int qq = 0; qq < TQ / kWarps; ++qq) {
                const int q = warp * (TQ / kWarps) + qq;
                if (q0 + q >= A.nq) break;
                int len = list_len[q];
                double thr = (len == k) ? list_s[(size_t)q * k + k - 1] : -INFINITY;
                bool full = (len == k);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double s = S[q * SPITCH + e * 32 + lane];
                    const int li = (int)(i0 - t_begin * TN) + half * 64 + e * 32 + lane;  // index inside slab
                    // survivors: strictly better than the current k-th (later index loses ties)
                    unsigned mask = __ballot_sync(0xffffffffu, full ? (s > thr) : (s > -INFINITY));
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const double cs = __shfl_sync(0xffffffffu, s, src);
                        const int ci = __shfl_sync(0xffffffffu, li, src);
                        // position = number of entries with score >= cs (sorted descending)
                        int cnt = 0;
                        double e0s = 0.0, e1s = 0.0;
                        int e0i = 0, e1i = 0;
                        const int p0 = lane, p1 = lane + 32;
                        if (p0 < len) {
                            e0s = list_s[(size_t)q * k + p0];
                            e0i = list_i[(size_t)q * k + p0];
                        }
                        if (p1 < len) {
                            e1s = list_s[(size_t)q * k + p1];
                            e1i = list_i[(size_t)q * k + p1];
                        }
                        cnt = __popc(__ballot_sync(0xffffffffu, p0 < len && e0s >= cs)) +
                              __popc(__ballot_sync(0xffffffffu, p1 < len && e1s >= cs));
                        __syncwarp();
                        if (p0 >= cnt && p0 < len && p0 + 1 < k) {
                            list_s[(size_t)q * k + p0 + 1] = e0s;
                            list_i[(size_t)q * k + p0 + 1] = e0i;
                        }
                        if (p1 >= cnt && p1 < len && p1 + 1 < k) {
                            list_s[(size_t)q * k + p1 + 1] = e1s;
                            list_i[(size_t)q * k + p1 + 1] = e1i;
                        }
                        if (lane == 0 && cnt < k) {
                            list_s[(size_t)q * k + cnt] = cs;
                            list_i[(size_t)q * k + cnt] = ci;
                        }
                        if (len < k) len++;
                        __syncwarp();
                        full = (len == k);
                        if (full) {
                            thr = list_s[(size_t)q * k + k - 1];
                            mask &= __ballot_sync(0xffffffffu, s > thr);
                        }
                    }
                }
                __syncwarp();   // every lane has read list_len[q]
                if (lane == 0) list_len[q] = len;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // ---- write this slab's lists
    for (int c = tid; c < TQ * k; c += kThreads) {
        const int q = c / k, p = c % k;
        const long long gq = q0 + q;
        if (gq >= A.nq) continue;
        const bool ok = p < list_len[q];
        const size_t o = ((size_t)slab * A.nq + gq) * k + p;
        A.part_score[o] = ok ? list_s[(size_t)q * k + p] : -INFINITY;
        A.part_idx[o] = ok ? list_i[(size_t)q * k + p] + (int)(t_begin * TN) : -1;
    }
}

// One warp per query: k-way merge of `parts` lists by (score desc, index asc).
template <typename IdxT>
__global__ void __launch_bounds__(256) topk_merge_kernel(const double *__restrict__ in_score,
                                                         const IdxT *__restrict__ in_idx, int parts,
                                                         long long nq, int k, long long index_offset,
                                                         int negate_sqrt, double *__restrict__ out_score,
                                                         long long *__restrict__ out_idx,
                                                         long long *__restrict__ out_count) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    // each part list is sorted: keep a cursor per part (parts <= 32*4 handled in strides)
    // simple selection: k rounds of warp arg-best over all candidates not yet taken.
    const int total = parts * k;
    int taken = 0;
    double last_s = INFINITY;
    long long last_i = -1;
    for (int r = 0; r < k; ++r) {
        double bs = -INFINITY;
        long long bi = -1;
        for (int c = lane; c < total; c += 32) {
            const int p = c / k, e = c % k;
            const size_t o = ((size_t)p * nq + q) * k + e;
            const long long ii = (long long)in_idx[o];
            if (ii < 0) continue;
            const double s = in_score[o];
            // candidate must come strictly after (last_s, last_i) in (score desc, idx asc) order
            const bool after = (s < last_s) || (s == last_s && ii > last_i);
            if (!after) continue;
            if (bi < 0 || s > bs || (s == bs && ii < bi)) {
                bs = s;
                bi = ii;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) {
                bs = os;
                bi = oi;
            }
        }
        if (bi < 0) break;
        if (lane == 0) {
            out_score[q * k + r] = negate_sqrt ? sqrt(fmax(-bs, 0.0)) : bs;
            out_idx[q * k + r] = bi + index_offset;
        }
        last_s = bs;
        last_i = bi;
        taken++;
    }
    if (lane == 0) {
        for (int r = taken; r < k; ++r) {
            out_score[q * k + r] = negate_sqrt ? INFINITY : -INFINITY;
            out_idx[q * k + r] = -1;
        }
        if (out_count) out_count[q] = taken;
    }
}

// Two-NN rescoring: exact direct-form distances for the 4 GEMM-form candidates of each sample
// (src/clustering.rs:125-130), two smallest -> d1, d2.  One warp per sample.
__global__ void __launch_bounds__(128) twonn_rescore_kernel(const double *__restrict__ rows, long long n, int f,
                                                            const double *__restrict__ qrows,
                                                            const long long *__restrict__ cand_idx, int ncand,
                                                            long long s, double *__restrict__ d1,
                                                            double *__restrict__ d2) {
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= s) return;
    const double *ri = qrows + w * (long long)f;   // the sample row itself (it may live on another GPU's shard)
    double m1 = INFINITY, m2 = INFINITY;
    for (int c = 0; c < ncand; ++c) {
        const long long j = cand_idx[w * ncand + c];
        if (j < 0) continue;
        const double *rj = rows + j * (long long)f;
        double acc = 0.0;
        for (int t = lane; t < f; t += 32) {
            const double df = ri[t] - rj[t];
            acc = fma(df, df, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double d = sqrt(acc);
        if (d < m1) {
            m2 = m1;
            m1 = d;
        } else if (d < m2) {
            m2 = d;
        }
    }
    if (lane == 0) {
        d1[w] = m1;
        d2[w] = m2;
    }
}

// Energy search, second pass: exact direct-form scores (src/energymaps.rs:884-894: d_lambda, |q - x| from the
// difference vector, bounded energy) for the kc candidates of the fused pass, then the best k by
// (score desc, index asc) -- the reference's stable descending sort.  One warp per query, kc <= 64.
__global__ void __launch_bounds__(128) energy_rescore_kernel(const double *__restrict__ items, const double *__restrict__ lambdas,
                                                             int f, const double *__restrict__ queries,
                                                             const double *__restrict__ lambda_q, long long nq, int kc,
                                                             int k, long long index_offset, double w_lambda, double w_dir,
                                                             const long long *__restrict__ cand_idx,
                                                             const long long *__restrict__ cand_cnt,
                                                             long long *__restrict__ idx_out, double *__restrict__ score_out,
                                                             long long *__restrict__ count_out) {
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= nq) return;
    const int cnt = (int)cand_cnt[w];
    const double *q = queries + w * (long long)f;
    const double lq = lambda_q[w];
    double sc[2] = {-INFINITY, -INFINITY};   // candidate c lives in lane c % 32, register c / 32
    long long id[2] = {-1, -1};
    for (int c = 0; c < cnt; ++c) {
        const long long gi = cand_idx[w * kc + c];
        const double *x = items + (gi - index_offset) * (long long)f;
        double acc = 0.0;
        for (int t = lane; t < f; t += 32) {
            const double df = q[t] - x[t];
            acc = fma(df, df, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double d = sqrt(acc);
        const double s = -(w_lambda * fabs(lq - lambdas[gi - index_offset]) + w_dir * fmin(d / (1.0 + d), 1.0));
        if ((c & 31) == lane) {
            sc[c >> 5] = s;
            id[c >> 5] = gi;
        }
    }
    const int kout = k < cnt ? k : cnt;
    // rank of every candidate = number of candidates that sort before it
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = h * 32 + lane;
        int rank = 0;
        for (int o = 0; o < cnt; ++o) {
            const double os = __shfl_sync(0xffffffffu, sc[o >> 5], o & 31);
            const long long oi = __shfl_sync(0xffffffffu, id[o >> 5], o & 31);
            if (c < cnt && (os > sc[h] || (os == sc[h] && oi < id[h]))) rank++;
        }
        if (c < cnt && rank < kout) {
            idx_out[w * k + rank] = id[h];
            score_out[w * k + rank] = sc[h];
        }
    }
    for (int r = kout + lane; r < k; r += 32) {
        idx_out[w * k + r] = -1;
        score_out[w * k + r] = 0.0;
    }
    if (lane == 0 && count_out) count_out[w] = kout;
}

size_t search_smem_bytes(int k, int kc) {
    size_t b = (size_t)2 * (TQ + TN) * (kc + 4) * 8;  // stages (S aliases)
    b += (size_t)(2 * TQ + 2 * TN) * 8;           // nq, lq, nx, lx
    b += (size_t)TQ * k * 8 + (size_t)TQ * k * 4; // lists
    b += (size_t)TQ * 4;                          // len
    b = (b + 7) & ~(size_t)7;
    b += (size_t)TQ * 8;                          // self idx
    return b + 16;
}

// Slab split of the item tiles.  A (query tile, slab) unit is one CTA and the CTAs run in waves of sm_count, so
// the time is ~ ceil(units / sm_count) * tiles_per_slab: pick the slab count that minimises it (a single unit
// spilling into an extra wave costs a whole slab time: 79 query tiles x 15 slabs = 8 waves + 1 CTA), each slab
// at least 4 tiles, fewer slabs preferred among near-equal costs (shorter merges, tighter prefilter bounds).
void pick_slabs(int sm_count, long long qtiles, long long ntiles, long long limit, int *nslabs_out, long long *tps_out) {
    long long max_slabs = (ntiles + 3) / 4;
    if (max_slabs > limit) max_slabs = limit;
    if (max_slabs < 1) max_slabs = 1;
    double best_cost = 0.0;
    long long best_tps = ntiles > 0 ? ntiles : 1;
    int best_ns = 1;
    for (long long ns = 1; ns <= max_slabs; ++ns) {
        const long long tps = (ntiles + ns - 1) / ns;
        const long long real_ns = (ntiles + tps - 1) / tps;
        const long long waves = (qtiles * real_ns + sm_count - 1) / sm_count;
        const double cost = (double)waves * (double)(tps + 1) * (1.0 + 5e-4 * (double)real_ns);
        if (ns == 1 || cost < best_cost) {
            best_cost = cost;
            best_tps = tps;
            best_ns = (int)real_ns;
        }
    }
    *nslabs_out = best_ns;
    *tps_out = best_tps;
}

}  // namespace

static int run_search(asb_ctx *ctx, int mode, SearchArgs &A, long long index_offset, int64_t *idx_d,
                      double *score_d, int64_t *count_d, const char *timer_name = nullptr) {
    const int k = A.k;
    if (k < 1 || k > 64) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "search: k=%d outside 1..64", k);
    if (A.n > 0x7fffff00ll) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "search: shard larger than 2^31 items");
    int nslabs = 1;
    pick_slabs(ctx->sm_count, (A.nq + TQ - 1) / TQ, (A.n + TN - 1) / TN, 4096, &nslabs, &A.tiles_per_slab);
    A.nslabs = nslabs;
    DevTmp<double> part_s;
    DevTmp<int> part_i;
    ASB_TRY(part_s.init(ctx, (size_t)nslabs * A.nq * k));
    ASB_TRY(part_i.init(ctx, (size_t)nslabs * A.nq * k));
    A.part_score = part_s.ptr;
    A.part_idx = part_i.ptr;
    int kc = 32;  // deeper chunks halve the barrier count; fall back when the top-k lists need the space
    if (search_smem_bytes(k, kc) > 220 * 1024) kc = 16;
    const size_t smem = search_smem_bytes(k, kc);
    const bool vec = (A.f % 2 == 0) && (((uintptr_t)A.items & 15) == 0) && (((uintptr_t)A.queries & 15) == 0);
    dim3 grid((unsigned)((A.nq + TQ - 1) / TQ), (unsigned)nslabs);
#define LAUNCH(M, V)                                                                                           \
    do {                                                                                                       \
        if (kc == 32) {                                                                                        \
            ASB_CUDA(ctx, cudaFuncSetAttribute(search_kernel<M, V, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem));                                                   \
            search_kernel<M, V, 32><<<grid, kThreads, smem, ctx->stream>>>(A);                                 \
        } else {                                                                                               \
            ASB_CUDA(ctx, cudaFuncSetAttribute(search_kernel<M, V, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem));                                                   \
            search_kernel<M, V, 16><<<grid, kThreads, smem, ctx->stream>>>(A);                                 \
        }                                                                                                      \
    } while (0)
    {
        KernelTimer kt(ctx, timer_name ? timer_name
                                       : (mode == MODE_COSINE ? "search_kernel" : (mode == MODE_ENERGY ? "energy_kernel" : "twonn_kernel")));
        if (mode == MODE_COSINE) {
            if (vec) LAUNCH(MODE_COSINE, true); else LAUNCH(MODE_COSINE, false);
        } else if (mode == MODE_ENERGY) {
            if (vec) LAUNCH(MODE_ENERGY, true); else LAUNCH(MODE_ENERGY, false);
        } else {
            if (vec) LAUNCH(MODE_L2, true); else LAUNCH(MODE_L2, false);
        }
    }
#undef LAUNCH
    ASB_TRY(asb_check_launch(ctx, "search_kernel"));
    const int wpb = 8;
    topk_merge_kernel<int><<<(unsigned)((A.nq + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(
        part_s.ptr, part_i.ptr, nslabs, A.nq, k, index_offset, mode == MODE_L2 ? 1 : 0, score_d,
        (long long *)idx_d, (long long *)count_d);
    ASB_TRY(asb_check_launch(ctx, "topk_merge_kernel"));
    return ASB_OK;
}

// The exact kernel ranks with DMMA-order dot products (1e-15 relative): ids are right except inside near-ties, but
// the scores are not the reference's bits and the order inside a near-tie may differ from the prefilter path, which
// rescoring in the reference's arithmetic.  This kernel gives both paths the same output: the exact kernel keeps a few
// spare candidates (kx >= k), every one is scored as src/core.rs:214-236,:135-165 does (sequential sums, products and
// sums rounded separately -- the same statements as pf_finish_kernel), and the best k by (score desc, index asc)
// are written; unused slots hold -inf / -1.  One warp per query, two candidates per lane (kx <= 64).
__global__ void __launch_bounds__(128) exact_rescore_kernel(const double *__restrict__ items, const double *__restrict__ lambdas,
                                                            const double *__restrict__ queries,
                                                            const double *__restrict__ lambda_q, long long nq, int f,
                                                            double alpha, long long index_offset,
                                                            const long long *__restrict__ in_idx, int kx, int k,
                                                            long long *__restrict__ idx_out, double *__restrict__ score_out,
                                                            long long *__restrict__ count_out, int *status) {
    const long long q = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const double *qr = queries + q * (long long)f;
    const double lq = lambda_q[q];
    double sc[2];
    long long id[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        id[h] = c < kx ? in_idx[q * kx + c] : -1;
        sc[h] = -INFINITY;
        if (id[h] >= 0) {
            const double *x = items + id[h] * (long long)f;
            double nq2 = 0.0, nx2 = 0.0, dot = 0.0;
#pragma unroll 4
            for (int j = 0; j < f; ++j) {
                const double qv = qr[j], xv = x[j];
                nq2 = __dadd_rn(nq2, __dmul_rn(qv, qv));
                nx2 = __dadd_rn(nx2, __dmul_rn(xv, xv));
                dot = __dadd_rn(dot, __dmul_rn(qv, xv));
            }
            const double denom = __dmul_rn(sqrt(nq2), sqrt(nx2));
            const double cosv = denom > 0.0 ? dot / denom : 0.0;
            const double lam = 1.0 - fmin(fabs(lq - lambdas[id[h]]), 1.0);
            const double s = __dadd_rn(__dmul_rn(alpha, cosv), __dmul_rn(1.0 - alpha, lam));
            if (s != s) atomicOr(status, STATUS_NAN);
            sc[h] = (s == s) ? s : -INFINITY;
        }
    }
    int rank[2] = {0, 0}, valid = 0;
    for (int src = 0; src < 64; ++src) {
        const double os = __shfl_sync(0xffffffffu, src < 32 ? sc[0] : sc[1], src & 31);
        const long long oi = __shfl_sync(0xffffffffu, src < 32 ? id[0] : id[1], src & 31);
        if (oi < 0) continue;   // warp-uniform
        ++valid;
#pragma unroll
        for (int h = 0; h < 2; ++h) rank[h] += (os > sc[h] || (os == sc[h] && oi < id[h])) ? 1 : 0;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (id[h] >= 0 && rank[h] < k) {
            score_out[q * k + rank[h]] = sc[h];
            idx_out[q * k + rank[h]] = id[h] + index_offset;
        }
    const int taken = valid < k ? valid : k;
    for (int r = taken + lane; r < k; r += 32) {
        score_out[q * k + r] = -INFINITY;
        idx_out[q * k + r] = -1;
    }
    if (lane == 0 && count_out) count_out[q] = taken;
}

#include "search_pf.cuh"

int asb_dev_search(asb_ctx *ctx, const double *items_d, const double *lambdas_d, const double *norms2_d,
                   int64_t n, int64_t f, const double *queries_d, const double *lambda_q_d, int64_t nq,
                   int64_t k, double alpha, int64_t index_offset, int64_t *idx_d, double *score_d,
                   int64_t *count_d, int *status_d) {
    if (n <= 0 || f <= 0 || nq <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "search: empty input");
    ctx->kernel_ms["search_pf_used"] = 0.0;
    DevTmp<double> qn2, xn2;
    ASB_TRY(qn2.init(ctx, (size_t)nq));
    ASB_TRY(asb_dev_norms2(ctx, queries_d, nq, f, qn2.ptr));
    if (!norms2_d) {
        ASB_TRY(xn2.init(ctx, (size_t)n));
        ASB_TRY(asb_dev_norms2(ctx, items_d, n, f, xn2.ptr));
        norms2_d = xn2.ptr;
    }
    SearchArgs A{};
    A.items = items_d;
    A.lambdas = lambdas_d;
    A.norms2 = norms2_d;
    A.queries = queries_d;
    A.lambda_q = lambda_q_d;
    A.qnorms2 = qn2.ptr;
    A.self_idx = nullptr;
    A.n = n;
    A.nq = nq;
    A.f = (int)f;
    A.k = (int)(k < n ? k : n);
    A.alpha = alpha;
    A.status = status_d;
    {
        auto it = ctx->options.find("search_prefilter");  // default on; 0 = always the exact FP64 kernel
        if (A.k == k && (it == ctx->options.end() || it->second != 0.0)) {
            bool done = false;
            ASB_TRY(run_search_pf(ctx, A, index_offset, idx_d, score_d, count_d, &done));
            if (done) return ASB_OK;
        }
    }
    // exact FP64 kernel with a few spare candidates, then the reference-order rescoring both paths share; the lists are
    // kx wide, the outputs k wide (k > n: the rest is padded with -inf / -1, count = n)
    const int k_eff = A.k;
    int kx = k_eff + 4;
    if (kx > 64) kx = k_eff > 64 ? k_eff : 64;
    if (kx > n) kx = (int)n;
    A.k = kx;
    DevTmp<double> ts;
    DevTmp<int64_t> ti;
    ASB_TRY(ts.init(ctx, (size_t)nq * kx));
    ASB_TRY(ti.init(ctx, (size_t)nq * kx));
    ASB_TRY(run_search(ctx, MODE_COSINE, A, 0, ti.ptr, ts.ptr, nullptr));
    exact_rescore_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, ctx->stream>>>(
        items_d, lambdas_d, queries_d, lambda_q_d, (long long)nq, (int)f, alpha, (long long)index_offset,
        (const long long *)ti.ptr, kx, (int)k, (long long *)idx_d, score_d, (long long *)count_d, status_d);
    return asb_check_launch(ctx, "exact_rescore_kernel");
}

int asb_dev_search_energy(asb_ctx *ctx, const double *items_d, const double *lambdas_d, const double *norms2_d,
                          int64_t n, int64_t f, const double *queries_d, const double *lambda_q_d, int64_t nq,
                          int64_t k, double w_lambda, double w_dirichlet, int64_t index_offset, int64_t *idx_d,
                          double *score_d, int64_t *count_d, int *status_d) {
    if (n <= 0 || f <= 0 || nq <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "search_energy: empty input");
    if (k < 1 || k > 60) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "search_energy: k=%lld outside 1..60", (long long)k);
    DevTmp<double> qn2, xn2, cs;
    DevTmp<int64_t> ci, cc;
    ASB_TRY(qn2.init(ctx, (size_t)nq));
    ASB_TRY(asb_dev_norms2(ctx, queries_d, nq, f, qn2.ptr));
    if (!norms2_d) {
        ASB_TRY(xn2.init(ctx, (size_t)n));
        ASB_TRY(asb_dev_norms2(ctx, items_d, n, f, xn2.ptr));
        norms2_d = xn2.ptr;
    }
    // the fused pass ranks by the GEMM-form distance (|q|^2 + |x|^2 - 2 q.x, absolute error ~1e-13 on d^2, i.e. up
    // to ~3e-7 on a distance near zero): take 4 spare candidates, then rescore exactly
    int64_t kc = k + 4;
    if (kc > n) kc = n;
    SearchArgs A{};
    A.items = items_d;
    A.lambdas = lambdas_d;
    A.norms2 = norms2_d;
    A.queries = queries_d;
    A.lambda_q = lambda_q_d;
    A.qnorms2 = qn2.ptr;
    A.self_idx = nullptr;
    A.n = n;
    A.nq = nq;
    A.f = (int)f;
    A.k = (int)kc;
    A.alpha = w_lambda;
    A.w_dir = w_dirichlet;
    A.status = status_d;
    ASB_TRY(cs.init(ctx, (size_t)nq * kc));
    ASB_TRY(ci.init(ctx, (size_t)nq * kc));
    ASB_TRY(cc.init(ctx, (size_t)nq));
    ASB_TRY(run_search(ctx, MODE_ENERGY, A, index_offset, ci.ptr, cs.ptr, cc.ptr));
    const int wpb = 4;
    energy_rescore_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(
        items_d, lambdas_d, (int)f, queries_d, lambda_q_d, (long long)nq, (int)kc, (int)k, (long long)index_offset,
        w_lambda, w_dirichlet, (const long long *)ci.ptr, (const long long *)cc.ptr, (long long *)idx_d, score_d,
        (long long *)count_d);
    return asb_check_launch(ctx, "energy_rescore_kernel");
}

int asb_dev_twonn(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, const int64_t *sample_d, int64_t s,
                  double *d1_d, double *d2_d) {
    if (n < 2 || f <= 0 || s <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "twonn: n=%lld s=%lld", (long long)n, (long long)s);
    // gather the sample rows (they are the "queries" of the contraction)
    DevTmp<double> q;
    ASB_TRY(q.init(ctx, (size_t)s * f));
    std::vector<int64_t> hs((size_t)s);
    ASB_CUDA(ctx, cudaMemcpyAsync(hs.data(), sample_d, s * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t t = 0; t < s; ++t) {
        if (hs[t] < 0 || hs[t] >= n) ASB_FAIL(ctx, ASB_ERR_INVALID, "twonn: sample index %lld out of range", (long long)hs[t]);
        ASB_CUDA(ctx, cudaMemcpyAsync(q.ptr + t * f, rows_d + hs[t] * f, f * sizeof(double),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return asb_dev_twonn_queries(ctx, q.ptr, sample_d, s, rows_d, n, f, d1_d, d2_d);
}

// The scan proper: for each of s query rows the two smallest Euclidean distances to the n rows, the row self_d[t] (an
// index into rows, or -1: the sample lives elsewhere) left out.  Fewer than two admissible rows leave +inf.
int asb_dev_twonn_queries(asb_ctx *ctx, const double *q_rows_d, const int64_t *self_d, int64_t s, const double *rows_d,
                          int64_t n, int64_t f, double *d1_d, double *d2_d) {
    if (n < 1 || f <= 0 || s <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "twonn: n=%lld s=%lld", (long long)n, (long long)s);
    DevTmp<double> qn2, xn2, cs;
    DevTmp<int64_t> ci, cc;
    DevTmp<int> status;
    struct QView {
        const double *ptr;
    } q{q_rows_d};
    const int64_t *sample_d = self_d;
    ASB_TRY(qn2.init(ctx, (size_t)s));
    ASB_TRY(xn2.init(ctx, (size_t)n));
    const int ncand = (int)(n < 4 ? n : 4);
    ASB_TRY(cs.init(ctx, (size_t)s * ncand));
    ASB_TRY(ci.init(ctx, (size_t)s * ncand));
    ASB_TRY(cc.init(ctx, (size_t)s));
    ASB_TRY(status.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(status.ptr, 0, sizeof(int), ctx->stream));
    ASB_TRY(asb_dev_norms2(ctx, q.ptr, s, f, qn2.ptr));
    ASB_TRY(asb_dev_norms2(ctx, rows_d, n, f, xn2.ptr));
    SearchArgs A{};
    A.items = rows_d;
    A.norms2 = xn2.ptr;
    A.queries = q.ptr;
    A.qnorms2 = qn2.ptr;
    A.self_idx = (const long long *)sample_d;
    A.n = n;
    A.nq = s;
    A.f = (int)f;
    A.k = ncand;
    A.alpha = 0.0;
    A.status = status.ptr;
    ctx->kernel_ms["twonn_pf_used"] = 0.0;
    {
        // default: the certified TF32 ranking on the tcgen05 tile + direct-form distances for the survivors (bit-identical
        // to the reference's arithmetic; 1.5 ms instead of 14.5 ms of FP64 tensor pipe at 1M x 384); 0 = the FP64 kernel
        auto it = ctx->options.find("twonn_prefilter");
        if ((it == ctx->options.end() || it->second != 0.0) && n - 1 >= 2 && um_wanted(ctx)) {
            DevTmp<double> dist;
            DevTmp<int64_t> pi, pc;
            ASB_TRY(dist.init(ctx, (size_t)s * 2));
            ASB_TRY(pi.init(ctx, (size_t)s * 2));
            ASB_TRY(pc.init(ctx, (size_t)s));
            bool done = false;
            ASB_TRY(run_search_pf_l2(ctx, rows_d, n, (int)f, q.ptr, s, xn2.ptr, qn2.ptr, (const long long *)sample_d, 2, pi.ptr,
                                     dist.ptr, pc.ptr, &done));
            if (done) {
                ASB_CUDA(ctx, cudaMemcpy2DAsync(d1_d, sizeof(double), dist.ptr, 2 * sizeof(double), sizeof(double), s,
                                                cudaMemcpyDeviceToDevice, ctx->stream));
                ASB_CUDA(ctx, cudaMemcpy2DAsync(d2_d, sizeof(double), dist.ptr + 1, 2 * sizeof(double), sizeof(double), s,
                                                cudaMemcpyDeviceToDevice, ctx->stream));
                ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the temporaries die with this scope
                ctx->kernel_ms["twonn_pf_used"] = 1.0;
                return ASB_OK;
            }
        }
    }
    ASB_TRY(run_search(ctx, MODE_L2, A, 0, ci.ptr, cs.ptr, cc.ptr));
    const int wpb = 4;
    twonn_rescore_kernel<<<(unsigned)((s + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(
        rows_d, (long long)n, (int)f, q.ptr, (const long long *)ci.ptr, ncand, (long long)s, d1_d, d2_d);
    return asb_check_launch(ctx, "twonn_rescore_kernel");
}

int asb_dev_top2_l2(asb_ctx *ctx, const double *q_d, int64_t m, int64_t f, const double *items_d, int64_t k_items,
                    const double *qn2_d, const double *xn2_d, const int64_t *minus1_d, int64_t *idx_d, double *dist_d,
                    int64_t *cnt_d, int *status_d) {
    if (m < 1 || f < 1 || k_items < 2) ASB_FAIL(ctx, ASB_ERR_INVALID, "top2: m=%lld k_items=%lld", (long long)m, (long long)k_items);
    {
        auto it = ctx->options.find("cluster_replay_tf32");   // opt-in: certified TF32 ranking + direct-form distances
        if (it != ctx->options.end() && it->second != 0.0 && k_items >= 3) {
            bool done = false;
            ASB_TRY(run_search_pf_l2(ctx, items_d, k_items, (int)f, q_d, m, xn2_d, qn2_d, nullptr, 2, idx_d, dist_d, cnt_d, &done));
            if (done) return ASB_OK;
        }
    }
    SearchArgs A{};
    A.items = items_d;
    A.norms2 = xn2_d;
    A.queries = q_d;
    A.qnorms2 = qn2_d;
    A.self_idx = (const long long *)minus1_d;   // -1 everywhere: no row is excluded
    A.n = k_items;
    A.nq = m;
    A.f = (int)f;
    A.k = 2;
    A.alpha = 0.0;
    A.status = status_d;
    return run_search(ctx, MODE_L2, A, 0, idx_d, dist_d, cnt_d, "cluster_top2_kernel");
}

int asb_dev_near_tf32(asb_ctx *ctx, const double *q_d, int64_t m, int64_t f, const double *items_d, int64_t k_items,
                      const double *qn2_d, const double *xn2_d, int64_t *near_idx_d, double *near_b_d, bool *done) {
    return run_near_umma(ctx, items_d, k_items, (int)f, q_d, m, xn2_d, qn2_d, near_idx_d, near_b_d, done);
}

int asb_dev_topk_merge(asb_ctx *ctx, const double *in_score_d, const int64_t *in_idx_d, int64_t parts, int64_t nq,
                       int64_t k, double *out_score_d, int64_t *out_idx_d, int64_t *out_count_d) {
    if (parts < 1 || nq < 1 || k < 1) ASB_FAIL(ctx, ASB_ERR_INVALID, "topk_merge: bad sizes");
    const int wpb = 8;
    topk_merge_kernel<long long><<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(
        in_score_d, (const long long *)in_idx_d, (int)parts, (long long)nq, (int)k, 0, 0, out_score_d,
        (long long *)out_idx_d, (long long *)out_count_d);
    return asb_check_launch(ctx, "topk_merge_kernel");
}

int asb_search_slab_plan(int sm_count, int64_t nq, int64_t n, int64_t max_slabs, int64_t *nslabs,
                         int64_t *tiles_per_slab) {
    if (sm_count < 1 || nq < 1 || n < 1 || max_slabs < 1 || !nslabs || !tiles_per_slab) return ASB_ERR_INVALID;
    int ns = 1;
    long long tps = 1;
    pick_slabs(sm_count, (nq + TQ - 1) / TQ, (n + TN - 1) / TN, max_slabs, &ns, &tps);
    *nslabs = ns;
    *tiles_per_slab = tps;
    return ASB_OK;
}
