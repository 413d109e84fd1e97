// search_pf.cuh -- K8 with a certified tensor-core prefilter (option "search_prefilter", included by search.cu).
//
// The FP64 tensor pipe (DMMA, 34 TFLOP/s measured) is the ceiling of the exact search kernel.  This path ranks
// every (query, item) pair with a cheap score whose distance to the reference's FP64 score is BOUNDED, keeps the
// pairs the bound cannot exclude from the top-k, and rescoring those few in the reference's own arithmetic:
//
//   1. pf_unit_rows_kernel   x^ = fl32(x / |x|) for items and queries (zero rows stay zero: cos = 0, core.rs:231-236).
//   2. search_pf_kernel      cos~ = <q^, x^> on the tensor cores (mma.sync m16n8k8 TF32 in 3xTF32 form: hi*lo + lo*hi +
//                            hi*hi, FP32 accumulate), 128 queries x 128 items per CTA tile, operands streamed in
//                            32-feature chunks through a 3-stage cp.async ring that runs across tile boundaries;
//                            s~ = alpha cos~ + (1 - alpha)(1 - min(|dlambda|, 1)) in FP64 (core.rs:135-165).
//                            Per (CTA, query) a sorted list of the k best s~ of the slab gives a lower bound of the
//                            k-th best approximate score A_k; the bounds of all slabs meet in gthr[q] (atomicMax).
//                            A pair is EMITTED to the query's candidate list when s~ >= bound - 2E.
//   3. pf_finish_kernel      A_k itself = the k-th largest s~ of the query's list (every member of the approximate
//                            top-k was emitted); candidates below A_k - 2E are dropped, the rest are scored exactly as the
//                            reference does (sequential, separately rounded sums: core.rs:214-236; oracle
//                            aso_search_lambda_aware) and the best k by (score desc, index asc) are written.
//
// Why no top-k member is lost: |s~ - s| <= E for every pair (below).  k items have s~ >= A_k, hence true scores
// >= A_k - E, hence the true k-th best S_k >= A_k - E.  A member of the true top-k (ties included) has s >= S_k, so
// s~ >= S_k - E >= A_k - 2E >= bound - 2E for every bound <= A_k -- and every slab-local k-th best, at any time,
// is such a bound.
//
// E = |alpha| E_cos + 1e-10, E_cos =
//     2^-22                      FP32 rounding of q^ and x^ (2 * 2^-24 * sum|q^_j x^_j|, Cauchy-Schwarz: <= 1), doubled
//   + 3 * 2^-20                  dropped lo*lo and the TF32 truncation of the lo operands
//   + (9 * chain + 16) * 2^-23   accumulation: chain = 3 * fp / 8 MMAs into one FP32 accumulator, each within
//                                9 * 2^-23 * max(|acc|, |a b|) <= 9 * 2^-23 of exact -- the tensor-core model of
//                                cluster_f32p.cuh (aligned truncating adders), validated on B200 by its
//                                cluster_check_tile option (worst observed error 0.05 of that bound);
// times 1.1.  The exact kernel is the fallback for everything the bound does not cover: non-finite rows or norms
// outside [1e-145, 1e145], a NaN score (the reference panics there: status from the exact kernel), k > 32, a
// candidate list that overflows.  The exact path ends in the same reference-order rescoring (exact_rescore_kernel
// over its top k + 4), so scores and tie order do not depend on which path ran; the rescored values are
// bit-identical to the oracle's.
#pragma once

namespace {

constexpr int PF_KC = 32;                  // features per chunk
constexpr int PF_PITCH = PF_KC + 4;        // floats: quarter-warp 16-byte fragment loads hit 32 different banks
constexpr int PF_STAGES = 3;
constexpr int PF_STAGE_FLOATS = (TQ + TN) * PF_PITCH;
constexpr int PF_SP = 136;                 // cos~ tile pitch (floats): conflict-free float2 stores and row scans
constexpr int PF_FLAG_FALLBACK = 1, PF_FLAG_OVERFLOW = 2;
constexpr int PF_FLAG_OVERFLOW_SOME = 4;   // cosine mode: the queries marked in PfArgs::ovf go to the exact kernel, the rest stand
constexpr int PF_MAXSEL = 2048;            // rescored candidates per query, at most (PfArgs::maxsel: what the scratch holds)
constexpr int PF_COSINE = 0;               // score = alpha cos + (1 - alpha) lambda proximity (search_lambda_aware)
constexpr int PF_L2 = 1;                   // score = -|q - x|^2 (nearest neighbours: Two-NN scan, replay top-2); opt-in
constexpr int PF_NEAR = 2;                 // tcgen05 tile only: nearest item + certified distance bounds (the replay)

struct PfArgs {
    const float *xf, *qf;   // n x fp, nq x fp unit rows
    int fp;                 // f rounded up to 32 (zero padded)
    const double *lambdas, *lambda_q;
    long long n, nq;
    int k;
    double alpha, band;     // band = 2E
    int nslabs;
    long long tiles_per_slab;
    unsigned long long *gthr;  // nq: order-preserving encoding of the best known lower bound of A_k
    int *cand_cnt;             // nq
    int *cand_idx;             // nq x cap (index inside the shard)
    float *cand_s;             // nq x cap
    int cap;
    int *flags;
    unsigned long long *diag;  // [0] candidates emitted, [1] candidates rescored
    int *status;
    // PF_L2 only: s~ = -(|q|^2 + |x|^2 - 2 |q| |x| cos~), |s~ - s| <= 2 |q| |x| E_cos + rounding
    const double *qn2, *xn2;                 // squared norms
    const double *qnrm, *xnrm;               // norms
    const long long *self_idx;               // per query: item index to leave out (clustering.rs:123), or null
    const unsigned long long *xn2max_bits;   // device scalar: bits of max |x|^2
    double band_rel, band_abs;               // band(q) = band_rel |q| max|x| + band_abs (|q|^2 + max|x|^2)
    // PF_NEAR only: per query the nearest item by the approximate score and certified bounds {dlo, dhi} of its
    // distance and slo, a lower bound of the distance to every OTHER item
    double e_cos;
    int epi_groups;                          // tcgen05 tile: epilogue groups (search_umma.cuh)
    int *ovf;                                // per query (or null): 1 = list or survivor overflow, handled per query
    // survivors of the final bound, per query, in global scratch (L2-resident; in shared memory they cost the finishing
    // kernel half its resident blocks)
    int *sel_idx;                            // nq x maxsel
    double *sel_s;                           // nq x maxsel
    int maxsel;
    long long *near_idx;
    double *near_b;                          // nq x 3
};

__device__ __forceinline__ unsigned long long pf_enc(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double pf_dec(unsigned long long e) {
    const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ void pf_split(float x, unsigned &hi, unsigned &lo) {  // x = hi + lo, hi a TF32 number
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void pf_mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) pf_init_kernel(unsigned long long *gthr, int *cand_cnt, long long nq) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) {
        gthr[q] = pf_enc(-INFINITY);
        cand_cnt[q] = 0;
    }
}

// One warp per row: out = fl32(row / sqrt(norm2)), zero padded to fp.
__global__ void __launch_bounds__(256) pf_unit_rows_kernel(const double *__restrict__ rows, const double *__restrict__ norms2,
                                                           long long n, int f, int fp, float *__restrict__ out,
                                                           int *__restrict__ flags, double *__restrict__ norm_out = nullptr) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double n2 = norms2[r];
    const bool ok = (n2 == 0.0) || (n2 >= 1e-290 && n2 <= 1e290);   // false for NaN / inf as well
    if (!ok && lane == 0) atomicOr(flags, PF_FLAG_FALLBACK);
    const double inv = (ok && n2 > 0.0) ? 1.0 / sqrt(n2) : 0.0;
    if (norm_out && lane == 0) norm_out[r] = ok ? sqrt(n2) : 0.0;
    const double *src = rows + r * (long long)f;
    float *dst = out + r * (long long)fp;
    for (int j = lane; j < fp; j += 32) dst[j] = (ok && j < f) ? (float)(src[j] * inv) : 0.0f;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) search_pf_kernel(PfArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *stages = reinterpret_cast<float *>(smem_raw);                    // PF_STAGES * PF_STAGE_FLOATS
    float *S = stages + PF_STAGES * PF_STAGE_FLOATS;                        // TQ * PF_SP
    double *list_s = reinterpret_cast<double *>(S + TQ * PF_SP);            // TQ * k, sorted descending
    double *sm_lq = list_s + (size_t)TQ * A.k;                              // TQ
    int *list_len = reinterpret_cast<int *>(sm_lq + TQ);                    // TQ
    double *sm_nq = reinterpret_cast<double *>(list_len + TQ);              // TQ   (PF_L2: |q|; sm_lq holds |q|^2)
    double *sm_band = sm_nq + TQ;                                           // TQ   (PF_L2)
    long long *sm_self = reinterpret_cast<long long *>(sm_band + TQ);       // TQ   (PF_L2)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    const int k = A.k, fp = A.fp;
    const long long q0 = (long long)blockIdx.x * TQ;
    const long long ntiles_total = (A.n + TN - 1) / TN;
    const long long t_begin = (long long)blockIdx.y * A.tiles_per_slab;
    long long t_end = t_begin + A.tiles_per_slab;
    if (t_end > ntiles_total) t_end = ntiles_total;

    for (int q = tid; q < TQ; q += kThreads) {
        const long long gq = q0 + q;
        const bool ok = gq < A.nq;
        if constexpr (MODE == PF_COSINE) {
            const double lq = ok ? A.lambda_q[gq] : 1.0;
            if (ok && blockIdx.y == 0 && lq == 0.0) atomicOr(A.status, STATUS_ZERO_LAMBDA);  // core.rs:773-776
            sm_lq[q] = lq;
        } else {
            const double xmax2 = __longlong_as_double((long long)*A.xn2max_bits);
            const double q2 = ok ? A.qn2[gq] : 0.0, qn = ok ? A.qnrm[gq] : 0.0;
            sm_lq[q] = q2;
            sm_nq[q] = qn;
            sm_band[q] = A.band_rel * qn * sqrt(xmax2) + A.band_abs * (q2 + xmax2);
            sm_self[q] = (ok && A.self_idx) ? A.self_idx[gq] : -1ll;
        }
        list_len[q] = 0;
    }
    const int nchunks = fp / PF_KC;
    const long long total = (t_end - t_begin) * nchunks;
    if (total <= 0) return;

    // copy descriptors: thread -> rows (tid / 8) + 64 m, 16-byte column (tid % 8) of the chunk, both operands
    const int lr = tid >> 3, lc = (tid & 7) * 4;
    const float *gqp[2];
    bool okq[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const long long row = q0 + lr + 64 * m;
        okq[m] = row < A.nq;
        gqp[m] = A.qf + (okq[m] ? row : 0) * (long long)fp + lc;
    }
    long long is_tile = t_begin;  // tile / chunk of the next copy to issue
    int is_c = 0, is_stage = 0;
    auto issue = [&]() {
        float *st = stages + is_stage * PF_STAGE_FLOATS;
        const int k0 = is_c * PF_KC;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            cp_async16(st + (lr + 64 * m) * PF_PITCH + lc, gqp[m] + k0, okq[m] ? 16 : 0);
            const long long row = is_tile * TN + lr + 64 * m;
            const bool ok = row < A.n;
            cp_async16(st + (TQ + lr + 64 * m) * PF_PITCH + lc, A.xf + (ok ? row : 0) * (long long)fp + lc + k0, ok ? 16 : 0);
        }
        if (++is_c == nchunks) {
            is_c = 0;
            ++is_tile;
        }
        if (++is_stage == PF_STAGES) is_stage = 0;
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[i][j][u] = 0.0f;

    issue();
    cp_async_commit();
    if (total > 1) issue();
    cp_async_commit();

    bool saw_nan = false;
    long long tile = t_begin;
    int c = 0, stage = 0;
    for (long long gc = 0; gc < total; ++gc) {
        cp_async_wait<1>();  // every group but the newest has landed: chunk gc is in its stage
        __syncthreads();     // ... for all threads, and the stage of chunk gc - 1 is free
        if (gc + 2 < total) issue();
        cp_async_commit();

        const float *qa = stages + stage * PF_STAGE_FLOATS + (wm * 32 + g) * PF_PITCH + 8 * t;
        const float *xb = stages + stage * PF_STAGE_FLOATS + (TQ + wn * 32 + g) * PF_PITCH + 8 * t;
        // thread (g, t) holds features 8t .. 8t+7 of its rows; MMA k-step (h, m) pairs the values (2m, 2m+1) of
        // the h-th 16-byte load -- A and B use the same feature-to-slot mapping, which is all a dot product needs
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float va[2][2][4], vb[4][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float4 p = *reinterpret_cast<const float4 *>(qa + (i * 16 + r * 8) * PF_PITCH + 4 * h);
                    va[i][r][0] = p.x, va[i][r][1] = p.y, va[i][r][2] = p.z, va[i][r][3] = p.w;
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 p = *reinterpret_cast<const float4 *>(xb + (j * 8) * PF_PITCH + 4 * h);
                vb[j][0] = p.x, vb[j][1] = p.y, vb[j][2] = p.z, vb[j][3] = p.w;
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    pf_split(va[i][0][2 * m], ah[i][0], al[i][0]);      // (row g,     slot t)
                    pf_split(va[i][1][2 * m], ah[i][1], al[i][1]);      // (row g + 8, slot t)
                    pf_split(va[i][0][2 * m + 1], ah[i][2], al[i][2]);  // (row g,     slot t + 4)
                    pf_split(va[i][1][2 * m + 1], ah[i][3], al[i][3]);  // (row g + 8, slot t + 4)
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    pf_split(vb[j][2 * m], bh[j][0], bl[j][0]);          // (slot t,     item g)
                    pf_split(vb[j][2 * m + 1], bh[j][1], bl[j][1]);      // (slot t + 4, item g)
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        pf_mma(acc[i][j], ah[i], bl[j]);
                        pf_mma(acc[i][j], al[i], bh[j]);
                        pf_mma(acc[i][j], ah[i], bh[j]);
                    }
            }
        }
        if (++stage == PF_STAGES) stage = 0;
        if (++c < nchunks) continue;
        c = 0;

        // ---- tile finished: cos~ -> S, then one warp per query scans its 128 scores
        const long long i0 = tile * TN;
        ++tile;
        double lx[4], lxn[MODE == PF_L2 ? 4 : 1];
        bool vx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long gi = i0 + 32 * e + lane;
            vx[e] = gi < A.n;
            if constexpr (MODE == PF_COSINE) {
                lx[e] = vx[e] ? A.lambdas[gi] : 0.0;
            } else {
                lx[e] = vx[e] ? A.xn2[gi] : 0.0;
                lxn[e] = vx[e] ? A.xnrm[gi] : 0.0;
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // c0:(row g, col 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)
                const int row = wm * 32 + i * 16 + g, col = wn * 32 + j * 8 + 2 * t;
                *reinterpret_cast<float2 *>(&S[row * PF_SP + col]) = make_float2(acc[i][j][0], acc[i][j][1]);
                *reinterpret_cast<float2 *>(&S[(row + 8) * PF_SP + col]) = make_float2(acc[i][j][2], acc[i][j][3]);
                acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;
            }
        __syncthreads();
        unsigned long long genc = pf_enc(-INFINITY);
        if (lane < TQ / kWarps && q0 + warp * (TQ / kWarps) + lane < A.nq)
            genc = __ldcg(&A.gthr[q0 + warp * (TQ / kWarps) + lane]);
        for (int qq = 0; qq < TQ / kWarps; ++qq) {
            const int q = warp * (TQ / kWarps) + qq;
            const long long gq = q0 + q;
            if (gq >= A.nq) break;
            const double gb = pf_dec(__shfl_sync(0xffffffffu, genc, qq));  // bound published by any slab
            const double lq = sm_lq[q];
            double *lst = list_s + (size_t)q * k;
            int len = list_len[q];
            double kth = (len == k) ? lst[k - 1] : -INFINITY;
            bool changed = false;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double cosv = (double)S[q * PF_SP + 32 * e + lane];
                double s;
                bool valid = vx[e];
                if constexpr (MODE == PF_COSINE) {
                    const double lam = 1.0 - fmin(fabs(lq - lx[e]), 1.0);      // core.rs:136-137
                    s = A.alpha * cosv + (1.0 - A.alpha) * lam;                // core.rs:165 (approximate cos)
                } else {
                    s = -(lq + lx[e] - 2.0 * sm_nq[q] * lxn[e] * cosv);        // -|q - x|^2 from the approximate cos
                    valid = valid && (i0 + 32 * e + lane) != sm_self[q];
                }
                if (valid && s != s) saw_nan = true;
                // keep the slab's k best approximate scores (only values above every known bound matter)
                unsigned mask = __ballot_sync(0xffffffffu, valid && s > fmax(kth, gb));
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const double cs = __shfl_sync(0xffffffffu, s, src);
                    const double es = (lane < len) ? lst[lane] : 0.0;
                    const int pos = __popc(__ballot_sync(0xffffffffu, lane < len && es >= cs));
                    __syncwarp();
                    if (lane >= pos && lane < len && lane + 1 < k) lst[lane + 1] = es;
                    if (lane == 0 && pos < k) lst[pos] = cs;
                    if (len < k) ++len;
                    __syncwarp();
                    changed = true;
                    if (len == k) {
                        kth = lst[k - 1];
                        mask &= __ballot_sync(0xffffffffu, s > fmax(kth, gb));
                    }
                }
                // emit what the bound cannot exclude
                const double bound = fmax(kth, gb);
                const double band = MODE == PF_COSINE ? A.band : sm_band[q];
                const unsigned em = __ballot_sync(0xffffffffu, valid && s >= bound - band);
                if (em) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&A.cand_cnt[gq], __popc(em));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if ((em >> lane) & 1u) {
                        const int p = base + __popc(em & ((1u << lane) - 1u));
                        if (p < A.cap) {
                            A.cand_idx[(size_t)gq * A.cap + p] = (int)(i0 + 32 * e + lane);
                            A.cand_s[(size_t)gq * A.cap + p] = (float)s;
                        }
                    }
                }
            }
            __syncwarp();   // every lane has read list_len[q]
            if (lane == 0) {
                list_len[q] = len;
                if (changed && len == k) atomicMax(&A.gthr[gq], pf_enc(kth));
            }
        }
    }
    if (saw_nan) atomicOr(A.flags, PF_FLAG_FALLBACK);
}

// One block per query: drop the candidates below the final bound, score the rest in the reference's arithmetic
// (src/core.rs:214-236, :135-165 -- sequential sums, products and sums rounded separately), select the best k by
// (score desc, index asc) = the reference's stable descending sort (:785-786).
//
// v2 (ncu of v1, profiles/r02_pf_finish_*: 3.5 ms at C3 for 0.3 ms of DRAM time -- a third of the samples in the 64
// block barriers of a 32-step bisection, a quarter in row reads at one 8-byte word per 32-byte sector, the rest in the
// k selection rounds with two barriers each):
//   * A_k by a radix select over the order-preserving keys, three 8-bit digits = 9 barriers; the low 8 bits of the
//     key are left at zero: a lower bound of A_k 2^-15 relative below it (the band is ~2^-12), which only admits a
//     few more candidates;
//   * the survivors' rows are read by the warp TOGETHER -- 16 features of two candidates per load instruction, two
//     full 128-byte lines -- into a padded shared-memory tile that every lane then walks for its own candidate in
//     feature order: same statements, same order, same bits, an eighth of the sectors;
//   * the best k by counting ranks (one pass over the m <= 1024 survivors per thread), one barrier.
constexpr int PF_FT = 16;                  // features per staged chunk
constexpr int PF_FTP = PF_FT + 1;          // tile pitch (doubles): lane c walks row c conflict-free

__device__ __forceinline__ unsigned pf_fkey(float v) {
    const unsigned b = __float_as_uint(v);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}

// dynamic shared memory: q row (f doubles, padded to PF_FT) | 4 tiles of 32 x PF_FTP doubles
template <int MODE>
__global__ void __launch_bounds__(128) pf_finish_kernel(PfArgs A, const double *__restrict__ items,
                                                        const double *__restrict__ queries, int f, long long index_offset,
                                                        double band_f, long long *__restrict__ idx_out,
                                                        double *__restrict__ score_out, long long *__restrict__ count_out) {
    extern __shared__ double pf_dyn[];
    __shared__ int hist[256];
    __shared__ int nsel, pick_digit, pick_rem;
    const long long q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = A.k;
    const int maxsel = A.maxsel;
    int *sel_idx = A.sel_idx + (size_t)q * maxsel;
    double *sel_s = A.sel_s + (size_t)q * maxsel;
    const int fq = (f + PF_FT - 1) / PF_FT * PF_FT;
    double *qs = pf_dyn;
    double *tile = pf_dyn + fq + (size_t)warp * 32 * PF_FTP;
    if (tid == 0) nsel = 0;
    const int cnt = A.cand_cnt[q];
    if (cnt > A.cap) {
        if (tid == 0) {
            if (A.ovf) {
                A.ovf[q] = 1;
                atomicOr(A.flags, PF_FLAG_OVERFLOW_SOME);
            } else {
                atomicOr(A.flags, PF_FLAG_OVERFLOW);
            }
        }
        return;
    }
    const double *qr = queries + q * (long long)f;
    for (int j = tid; j < fq; j += 128) qs[j] = j < f ? qr[j] : 0.0;
    // Every item of the approximate top-k passed the emission test (its s~ >= A_k >= any bound), so the k-th largest
    // s~ of the list IS A_k.
    const float *cs = A.cand_s + (size_t)q * A.cap;
    if constexpr (MODE == PF_L2) {   // per-query band; the stored floats round |s~| <= 2 (|q|^2 + max|x|^2)
        const double xmax2 = __longlong_as_double((long long)*A.xn2max_bits);
        const double q2 = A.qn2[q];
        band_f = A.band_rel * A.qnrm[q] * sqrt(xmax2) + A.band_abs * (q2 + xmax2) + 2.5e-7 * (q2 + xmax2);
    }
    double thr = pf_dec(A.gthr[q]) - band_f;
    if (cnt >= k) {
        unsigned prefix = 0;   // the digits fixed so far (high bits of the key)
        int rem = k;           // rank still to be found inside the prefix bucket
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int b = tid; b < 256; b += 128) hist[b] = 0;
            __syncthreads();
            // warp-aggregated: the scores of a query share their leading digits (0.9x ...), and 32 shared-memory atomics on
            // ONE address serialise -- the first version of this loop was most of the kernel's 190 us per block
            for (int c0 = 0; c0 < cnt; c0 += 128) {
                const int c = c0 + tid;
                unsigned key = 0u;
                bool in = c < cnt;
                if (in) {
                    key = pf_fkey(cs[c]);
                    in = pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8));
                }
                const unsigned digit = (key >> shift) & 255u;
                const unsigned peers = __match_any_sync(0xffffffffu, in ? digit : 256u + (unsigned)lane);
                if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
            }
            __syncthreads();
            if (warp == 0) {   // largest digit d with #(digit >= d) >= rem; lane l owns digits 8 l .. 8 l + 7
                int loc[8], tot = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    loc[i] = hist[8 * lane + i];
                    tot += loc[i];
                }
                int suf = tot;   // suffix sum over the lanes: items in this lane's digits and all higher ones
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_down_sync(0xffffffffu, suf, o);
                    if (lane + o < 32) suf += v;
                }
                const int above = suf - tot;
                const bool mine = above < rem && suf >= rem;   // exactly one lane
                if (mine) {
                    int acc = above, d = 7;
                    for (; d > 0; --d) {
                        if (acc + loc[d] >= rem) break;
                        acc += loc[d];
                    }
                    pick_digit = 8 * lane + d;
                    pick_rem = rem - acc;
                }
            }
            __syncthreads();
            prefix |= (unsigned)pick_digit << shift;
            rem = pick_rem;
        }
        const float akf = __uint_as_float((prefix >> 31) ? (prefix & 0x7fffffffu) : ~prefix);   // <= A_k, within 2^-15
        thr = fmax(thr, (double)akf - band_f);
    }
    __syncthreads();
    for (int c = tid; c < cnt; c += 128)
        if ((double)cs[c] >= thr) {
            const int p = atomicAdd(&nsel, 1);
            if (p < maxsel) sel_idx[p] = A.cand_idx[(size_t)q * A.cap + c];
        }
    __syncthreads();
    const int m = nsel;
    if (m > maxsel) {
        if (tid == 0) {
            if (A.ovf) {
                A.ovf[q] = 1;
                atomicOr(A.flags, PF_FLAG_OVERFLOW_SOME);
            } else {
                atomicOr(A.flags, PF_FLAG_OVERFLOW);
            }
        }
        return;
    }
    if (tid == 0) {
        atomicAdd(&A.diag[0], (unsigned long long)cnt);
        atomicAdd(&A.diag[1], (unsigned long long)m);
    }
    const double lq = MODE == PF_L2 ? 0.0 : A.lambda_q[q];
    // ---- exact scores: warp w takes candidates 32 (w + 4 r) .. + 31, lane = candidate
    const int half = lane >> 4, l16 = lane & 15;
    for (int base = 32 * warp; base < m; base += 128) {
        const int nc = m - base < 32 ? m - base : 32;
        const int li = lane < nc ? sel_idx[base + lane] : 0;
        // the whole row to L2 with one request (one page lookup, one DRAM page opened once): the 16-feature pieces below
        // then come from L2 instead of 24 separate trips to scattered DRAM pages (ncu, profiles/r02_pf_finish_v2_*: 44 % of
        // the samples waited on those loads at 1 TB/s)
        if (lane < nc && (f & 1) == 0 && ((((size_t)items) & 15) == 0))
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(items + (long long)li * f), "r"((unsigned)(f * 8)) : "memory");
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;   // L2: acc0 = sum (q - x)^2;  cosine: |q|^2, |x|^2, q.x
        for (int j0 = 0; j0 < f; j0 += PF_FT) {
            __syncwarp();
            {   // all 16 loads of the chunk in flight before the first store
                double v[16];
                const int fj = j0 + l16;
#pragma unroll
                for (int ii = 0; ii < 16; ++ii) {
                    const int i = 2 * ii + half;
                    const int ri = __shfl_sync(0xffffffffu, li, i);   // candidate i's row (lane i holds it)
                    v[ii] = (i < nc && fj < f) ? __ldg(items + (long long)ri * f + fj) : 0.0;
                }
#pragma unroll
                for (int ii = 0; ii < 16; ++ii) {
                    const int i = 2 * ii + half;
                    if (i < nc) tile[i * PF_FTP + l16] = v[ii];
                }
            }
            __syncwarp();
            if (lane < nc) {
                const double *tr = tile + lane * PF_FTP;
                const int jn = f - j0 < PF_FT ? f - j0 : PF_FT;
                if (jn == PF_FT) {
#pragma unroll
                    for (int jj = 0; jj < PF_FT; ++jj) {
                        const double qv = qs[j0 + jj], xv = tr[jj];
                        if constexpr (MODE == PF_L2) {
                            const double df = __dsub_rn(qv, xv);                 // src/clustering.rs:125-130
                            acc0 = __dadd_rn(acc0, __dmul_rn(df, df));
                        } else {
                            acc0 = __dadd_rn(acc0, __dmul_rn(qv, qv));
                            acc1 = __dadd_rn(acc1, __dmul_rn(xv, xv));
                            acc2 = __dadd_rn(acc2, __dmul_rn(qv, xv));
                        }
                    }
                } else {
                    for (int jj = 0; jj < jn; ++jj) {
                        const double qv = qs[j0 + jj], xv = tr[jj];
                        if constexpr (MODE == PF_L2) {
                            const double df = __dsub_rn(qv, xv);
                            acc0 = __dadd_rn(acc0, __dmul_rn(df, df));
                        } else {
                            acc0 = __dadd_rn(acc0, __dmul_rn(qv, qv));
                            acc1 = __dadd_rn(acc1, __dmul_rn(xv, xv));
                            acc2 = __dadd_rn(acc2, __dmul_rn(qv, xv));
                        }
                    }
                }
            }
        }
        if (lane < nc) {
            double sres;
            if constexpr (MODE == PF_L2) {
                sres = (acc0 == acc0) ? -acc0 : -INFINITY;
            } else {
                const double denom = __dmul_rn(sqrt(acc0), sqrt(acc1));       // core.rs:230
                const double cosv = denom > 0.0 ? acc2 / denom : 0.0;        // :231-236
                const double lam = 1.0 - fmin(fabs(lq - A.lambdas[li]), 1.0);  // :136-137
                sres = __dadd_rn(__dmul_rn(A.alpha, cosv), __dmul_rn(1.0 - A.alpha, lam));  // :165
                if (sres != sres) {
                    atomicOr(A.status, STATUS_NAN);
                    sres = -INFINITY;
                }
            }
            sel_s[base + lane] = sres;
        }
    }
    __syncthreads();
    // ---- rank by (score desc, index asc): position = number of survivors that come first
    for (int c = tid; c < m; c += 128) {
        const double sc = sel_s[c];
        const int ic = sel_idx[c];
        int rank = 0;
        for (int o = 0; o < m; ++o) {
            const double so = sel_s[o];
            rank += (so > sc || (so == sc && sel_idx[o] < ic)) ? 1 : 0;
        }
        if (rank < k) {
            score_out[q * k + rank] = MODE == PF_L2 ? sqrt(fmax(-sc, 0.0)) : sc;
            idx_out[q * k + rank] = (long long)ic + index_offset;
        }
    }
    const int taken = m < k ? m : k;
    for (int r = taken + tid; r < k; r += 128) {
        score_out[q * k + r] = MODE == PF_L2 ? INFINITY : -INFINITY;
        idx_out[q * k + r] = -1;
    }
    if (tid == 0 && count_out) count_out[q] = taken;
}

// survivors the scratch holds per query: PF_MAXSEL, less when many queries share 512 MB (the replay's PF_L2 ranking of a
// whole chunk: k = 2, a handful of survivors per row)
static int pf_maxsel(long long nq) {
    long long m = (512ll << 20) / (12 * (nq > 0 ? nq : 1));
    if (m > PF_MAXSEL) m = PF_MAXSEL;
    if (m < 64) m = 64;
    return (int)m;
}

static size_t pf_finish_smem(int f) {
    const int fq = (f + PF_FT - 1) / PF_FT * PF_FT;
    return ((size_t)fq + 4 * 32 * PF_FTP) * sizeof(double);
}

__global__ void __launch_bounds__(256) pf_max_kernel(const double *__restrict__ v, long long n, unsigned long long *__restrict__ out_bits) {
    double mx = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) mx = fmax(mx, v[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx == mx) atomicMax(out_bits, (unsigned long long)__double_as_longlong(mx));   // >= 0: bits order
}

size_t pf_smem_bytes(int k, int mode = PF_COSINE) {
    return (size_t)PF_STAGES * PF_STAGE_FLOATS * 4 + (size_t)TQ * PF_SP * 4 + (size_t)TQ * k * 8 + (size_t)TQ * 8 +
           (size_t)TQ * 4 + 16 + (mode == PF_L2 ? (size_t)TQ * 24 : 0);
}

}  // namespace

#include "search_umma.cuh"   // the same tile on tcgen05 / TMA / TMEM (needs PfArgs and the helpers above)

// hi / lo planes for the tcgen05 tile: TF32 planes (fp32 words) or BF16 planes in the same buffers (half used)
#define UM_SPLIT_LAUNCH(grid, rows, norms2, n_, f_, fp_, hi_, lo_, flags_, nout_)                                          \
    do {                                                                                                                 \
        if (bf16)                                                                                                        \
            um_split_rows_bf16_kernel<<<(grid), 256, 0, ctx->stream>>>(rows, norms2, n_, f_, fp_,                        \
                                                                       reinterpret_cast<__nv_bfloat16 *>(hi_),          \
                                                                       reinterpret_cast<__nv_bfloat16 *>(lo_), flags_, nout_); \
        else                                                                                                             \
            um_split_rows_kernel<<<(grid), 256, 0, ctx->stream>>>(rows, norms2, n_, f_, fp_, hi_, lo_, flags_, nout_);   \
    } while (0)

__global__ void __launch_bounds__(128) pf_gather_queries_kernel(const double *__restrict__ queries, const double *__restrict__ lambda_q,
                                                                const double *__restrict__ qn2, int f,
                                                                const long long *__restrict__ list, double *__restrict__ out_q,
                                                                double *__restrict__ out_lq, double *__restrict__ out_qn) {
    const long long s = blockIdx.x, q = list[s];
    for (int j = threadIdx.x; j < f; j += 128) out_q[s * f + j] = queries[q * f + j];
    if (threadIdx.x == 0) {
        out_lq[s] = lambda_q[q];
        out_qn[s] = qn2[q];
    }
}
__global__ void __launch_bounds__(64) pf_scatter_results_kernel(const long long *__restrict__ list, int k,
                                                                const long long *__restrict__ in_i, const double *__restrict__ in_s,
                                                                const long long *__restrict__ in_c, long long *__restrict__ idx_out,
                                                                double *__restrict__ score_out, long long *__restrict__ count_out) {
    const long long s = blockIdx.x, q = list[s];
    for (int r = threadIdx.x; r < k; r += 64) {
        idx_out[q * k + r] = in_i[s * k + r];
        score_out[q * k + r] = in_s[s * k + r];
    }
    if (threadIdx.x == 0 && count_out) count_out[q] = in_c[s];
}

// Tries the prefilter path; *done = true when idx/score/count hold the final answer.  *done = false (with ASB_OK)
// means "not applicable / not certain": the caller runs the exact kernel, which then decides everything.
static int run_search_pf(asb_ctx *ctx, const SearchArgs &SA, long long index_offset, int64_t *idx_d, double *score_d,
                         int64_t *count_d, bool *done) {
    *done = false;
    ctx->kernel_ms["search_pf_used"] = 0.0;
    const int k = SA.k, f = SA.f;
    const long long n = SA.n, nq = SA.nq;
    if (k < 1 || k > 32 || n < 1024 || n > 0x7fffff00ll || !(fabs(SA.alpha) <= 1e6)) return ASB_OK;
    const int fp = (f + 31) & ~31;
    const long long qtiles = (nq + TQ - 1) / TQ, ntiles = (n + TN - 1) / TN;
    // few, long slabs: a slab's own k-th best is the bound its first tiles are held to
    int nslabs = 1;
    long long tps = ntiles;
    pick_slabs(ctx->sm_count, qtiles, ntiles, 64, &nslabs, &tps);
    // candidate capacity: a stream of m items passes a running k-th best about k ln(m / k) times
    const double slab_items = (double)tps * TN;
    double conc = (double)ctx->sm_count / (double)qtiles + 2.0;
    if (conc > nslabs) conc = nslabs;
    const double est = k * (log(fmax((double)n / k, 2.0)) + 2.0) + conc * k * (log(fmax(slab_items / k, 2.0)) + 2.0);
    int cap = 512;
    while (cap < 4.0 * est && cap < 16384) cap *= 2;
    if ((double)nq * cap * 8.0 > 2e9) return ASB_OK;
    const size_t smem = pf_smem_bytes(k);
    if (smem > 227 * 1024) return ASB_OK;

    const bool bf16 = um_wanted(ctx) && um_bf16(ctx);   // BF16x3 planes on the tcgen05 tile (the bound covers both tiles)
    const double e_cos = um_e_cos(fp, bf16);
    const double E = fabs(SA.alpha) * e_cos + 1e-10;
    const double band = 2.0 * E;
    const double band_f = band + 1.1920928955078125e-7 * (fabs(SA.alpha) + fabs(1.0 - SA.alpha) + 1.0);  // cand_s is a float

    DevTmp<float> xf, qf, xlo, qlo, cand_s;
    DevTmp<int> cand_cnt, cand_idx, flags;
    DevTmp<unsigned long long> gthr, diag;
    // a failed allocation only means "use the exact kernel"
    if (xf.init(ctx, (size_t)n * fp) != ASB_OK || qf.init(ctx, (size_t)nq * fp) != ASB_OK ||
        cand_s.init(ctx, (size_t)nq * cap) != ASB_OK || cand_idx.init(ctx, (size_t)nq * cap) != ASB_OK) {
        cudaGetLastError();
        return ASB_OK;
    }
    // Blackwell tensor pipe (search_umma.cuh): needs the lo planes and four TMA descriptors; anything missing falls
    // back to the mma.sync tile above -- same certificate, same lists
    UmMaps maps;
    bool umma = um_wanted(ctx) && um_smem_bytes(k, PF_COSINE) <= 227 * 1024;
    if (umma && (xlo.init(ctx, (size_t)n * fp) != ASB_OK || qlo.init(ctx, (size_t)nq * fp) != ASB_OK)) {
        cudaGetLastError();
        umma = false;
    }
    if (umma) umma = um_make_maps(ctx, &maps, qf.ptr, qlo.ptr, nq, xf.ptr, xlo.ptr, n, fp);
    if (umma) {
        const long long utiles = (n + UM_TN - 1) / UM_TN;
        um_pick_slabs(ctx, (nq + UM_TQ - 1) / UM_TQ, utiles, fp, &nslabs, &tps);
    }
    DevTmp<int> ovf, sel_idx_g;
    DevTmp<double> sel_s_g;
    ASB_TRY(cand_cnt.init(ctx, (size_t)nq));
    ASB_TRY(ovf.init(ctx, (size_t)nq));
    ASB_TRY(flags.init(ctx, 1));
    ASB_TRY(gthr.init(ctx, (size_t)nq));
    ASB_TRY(diag.init(ctx, 2));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(ovf.ptr, 0, (size_t)nq * sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(diag.ptr, 0, 2 * sizeof(unsigned long long), ctx->stream));
    pf_init_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(gthr.ptr, cand_cnt.ptr, nq);
    ASB_TRY(asb_check_launch(ctx, "pf_init_kernel"));
    {
        KernelTimer kt(ctx, "search_pf_prep");
        if (umma) {
            UM_SPLIT_LAUNCH((unsigned)((n + 7) / 8), SA.items, SA.norms2, n, f, fp, xf.ptr, xlo.ptr, flags.ptr, nullptr);
            UM_SPLIT_LAUNCH((unsigned)((nq + 7) / 8), SA.queries, SA.qnorms2, nq, f, fp, qf.ptr, qlo.ptr, flags.ptr, nullptr);
        } else {
            pf_unit_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(SA.items, SA.norms2, n, f, fp, xf.ptr, flags.ptr);
            pf_unit_rows_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, ctx->stream>>>(SA.queries, SA.qnorms2, nq, f, fp, qf.ptr, flags.ptr);
        }
    }
    ASB_TRY(asb_check_launch(ctx, "pf_unit_rows_kernel"));
    ctx->launches++;

    PfArgs A{};
    A.xf = xf.ptr;
    A.qf = qf.ptr;
    A.fp = fp;
    A.lambdas = SA.lambdas;
    A.lambda_q = SA.lambda_q;
    A.n = n;
    A.nq = nq;
    A.k = k;
    A.alpha = SA.alpha;
    A.band = band;
    A.nslabs = nslabs;
    A.tiles_per_slab = tps;
    A.gthr = gthr.ptr;
    A.cand_cnt = cand_cnt.ptr;
    A.cand_idx = cand_idx.ptr;
    A.cand_s = cand_s.ptr;
    A.cap = cap;
    A.flags = flags.ptr;
    A.diag = diag.ptr;
    A.status = SA.status;
    A.ovf = ovf.ptr;
    A.maxsel = pf_maxsel(nq);
    ASB_TRY(sel_idx_g.init(ctx, (size_t)nq * A.maxsel));
    ASB_TRY(sel_s_g.init(ctx, (size_t)nq * A.maxsel));
    A.sel_idx = sel_idx_g.ptr;
    A.sel_s = sel_s_g.ptr;
    ctx->kernel_ms["search_pf_umma"] = umma ? 1.0 : 0.0;
    if (umma) {
        ASB_TRY(um_launch<PF_COSINE>(ctx, maps, A, nslabs, "search_pf_kernel"));
    } else {
        ASB_CUDA(ctx, cudaFuncSetAttribute(search_pf_kernel<PF_COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt(ctx, "search_pf_kernel");
        search_pf_kernel<PF_COSINE><<<dim3((unsigned)qtiles, (unsigned)nslabs), kThreads, smem, ctx->stream>>>(A);
    }
    ASB_TRY(asb_check_launch(ctx, "search_pf_kernel"));
    {
        KernelTimer kt(ctx, "search_pf_finish");
        ASB_CUDA(ctx, cudaFuncSetAttribute(pf_finish_kernel<PF_COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)pf_finish_smem(f)));
        pf_finish_kernel<PF_COSINE><<<(unsigned)nq, 128, pf_finish_smem(f), ctx->stream>>>(A, SA.items, SA.queries, f, index_offset, band_f,
                                                                (long long *)idx_d, score_d, (long long *)count_d);
    }
    ASB_TRY(asb_check_launch(ctx, "pf_finish_kernel"));
    int hflags = 0;
    unsigned long long hdiag[2] = {0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(&hflags, flags.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(hdiag, diag.ptr, sizeof(hdiag), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->kernel_ms["search_pf_flags"] = (double)hflags;
    ctx->kernel_ms["search_pf_cap"] = (double)cap;
    ctx->kernel_ms["search_pf_slabs"] = (double)nslabs;
    ctx->kernel_ms["search_pf_band"] = band;
    ctx->kernel_ms["search_pf_candidates"] = (double)hdiag[0];
    ctx->kernel_ms["search_pf_rescored"] = (double)hdiag[1];
    if ((hflags & ~PF_FLAG_OVERFLOW_SOME) != 0) return ASB_OK;  // the exact kernel decides (and reports NaN scores the way the reference does)
    ctx->kernel_ms["search_pf_overflow_queries"] = 0.0;
    if (hflags & PF_FLAG_OVERFLOW_SOME) {
        // Some queries have more near-ties than the lists hold (tightly packed scores: every item of a blob within the
        // band).  They alone go to the exact FP64 kernel + reference-order rescoring; the other queries' answers stand.
        std::vector<int> hov((size_t)nq);
        ASB_CUDA(ctx, cudaMemcpy(hov.data(), ovf.ptr, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<long long> list;
        for (long long q = 0; q < nq; ++q)
            if (hov[(size_t)q]) list.push_back(q);
        ctx->kernel_ms["search_pf_overflow_queries"] = (double)list.size();
        if ((long long)list.size() * 2 > nq) return ASB_OK;   // most of the batch: one exact pass for all of it
        const long long ns = (long long)list.size();
        DevTmp<long long> lidx, sub_i, out_i, sub_c;
        DevTmp<double> sub_q, sub_lq, sub_qn, sub_s, out_s;
        int kx = k + 4 > 64 ? 64 : k + 4;
        if (kx > n) kx = (int)n;
        ASB_TRY(lidx.init(ctx, (size_t)ns));
        ASB_TRY(sub_q.init(ctx, (size_t)ns * f));
        ASB_TRY(sub_lq.init(ctx, (size_t)ns));
        ASB_TRY(sub_qn.init(ctx, (size_t)ns));
        ASB_TRY(sub_i.init(ctx, (size_t)ns * kx));
        ASB_TRY(sub_s.init(ctx, (size_t)ns * kx));
        ASB_TRY(out_i.init(ctx, (size_t)ns * k));
        ASB_TRY(out_s.init(ctx, (size_t)ns * k));
        ASB_TRY(sub_c.init(ctx, (size_t)ns));
        ASB_CUDA(ctx, cudaMemcpyAsync(lidx.ptr, list.data(), (size_t)ns * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        pf_gather_queries_kernel<<<(unsigned)ns, 128, 0, ctx->stream>>>(SA.queries, SA.lambda_q, SA.qnorms2, f, lidx.ptr, sub_q.ptr,
                                                                       sub_lq.ptr, sub_qn.ptr);
        ASB_TRY(asb_check_launch(ctx, "pf_gather_queries_kernel"));
        SearchArgs S2 = SA;
        S2.queries = sub_q.ptr;
        S2.lambda_q = sub_lq.ptr;
        S2.qnorms2 = sub_qn.ptr;
        S2.nq = ns;
        S2.k = kx;
        ASB_TRY(run_search(ctx, MODE_COSINE, S2, 0, (int64_t *)sub_i.ptr, sub_s.ptr, nullptr, "search_pf_overflow_exact"));
        exact_rescore_kernel<<<(unsigned)((ns + 3) / 4), 128, 0, ctx->stream>>>(SA.items, SA.lambdas, sub_q.ptr, sub_lq.ptr, ns, f,
                                                                               SA.alpha, index_offset, sub_i.ptr, kx, k, out_i.ptr,
                                                                               out_s.ptr, sub_c.ptr, SA.status);
        ASB_TRY(asb_check_launch(ctx, "exact_rescore_kernel"));
        pf_scatter_results_kernel<<<(unsigned)ns, 64, 0, ctx->stream>>>(lidx.ptr, k, out_i.ptr, out_s.ptr, sub_c.ptr, (long long *)idx_d,
                                                                       score_d, (long long *)count_d);
        ASB_TRY(asb_check_launch(ctx, "pf_scatter_results_kernel"));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `list` and the temporaries go out of scope
    }
    ctx->kernel_ms["search_pf_used"] = 1.0;
    *done = true;
    return ASB_OK;
}

// The same scheme for nearest neighbours in Euclidean distance (PF_L2; options "twonn_prefilter" and
// "cluster_replay_tf32", both off by default -- written after the round's GPU budget was spent): the k nearest items of
// every query, ids and distances, the distances computed in the reference's direct form (sequential, separately
// rounded: src/clustering.rs:125-130) for the candidates the certified score cannot exclude.  self_idx (or null) names
// one item per query to leave out.  *done = false: the caller uses the exact FP64 kernel.
static int run_search_pf_l2(asb_ctx *ctx, const double *items_d, long long n, int f, const double *queries_d, long long nq,
                            const double *xn2_d, const double *qn2_d, const long long *self_idx_d, int k,
                            int64_t *idx_d, double *dist_d, int64_t *count_d, bool *done) {
    *done = false;
    if (k < 1 || k > 32 || n < (long long)k + 1 || n > 0x7fffff00ll || nq < 1) return ASB_OK;
    const int fp = (f + 31) & ~31;
    const long long qtiles = (nq + TQ - 1) / TQ, ntiles = (n + TN - 1) / TN;
    int nslabs = 1;
    long long tps = ntiles;
    pick_slabs(ctx->sm_count, qtiles, ntiles, 64, &nslabs, &tps);
    const double slab_items = (double)tps * TN;
    double conc = (double)ctx->sm_count / (double)qtiles + 2.0;
    if (conc > nslabs) conc = nslabs;
    const double est = k * (log(fmax((double)n / k, 2.0)) + 2.0) + conc * k * (log(fmax(slab_items / k, 2.0)) + 2.0);
    int cap = 64;
    while (cap < 4.0 * est && cap < 16384) cap *= 2;
    while (cap / 2 >= n && cap > 64) cap /= 2;            // never more candidates than items
    if ((double)nq * cap * 8.0 > 2e9) return ASB_OK;
    const size_t smem = pf_smem_bytes(k, PF_L2);
    if (smem > 227 * 1024) return ASB_OK;
    const bool bf16 = um_wanted(ctx) && um_bf16(ctx);   // BF16x3 planes on the tcgen05 tile (the bound covers both tiles)
    const double e_cos = um_e_cos(fp, bf16);

    DevTmp<float> xf, qf, xlo, qlo, cand_s;
    DevTmp<double> xnrm, qnrm;
    DevTmp<int> cand_cnt, cand_idx, flags, status;
    DevTmp<unsigned long long> gthr, diag, xmax;
    if (xf.init(ctx, (size_t)n * fp) != ASB_OK || qf.init(ctx, (size_t)nq * fp) != ASB_OK ||
        cand_s.init(ctx, (size_t)nq * cap) != ASB_OK || cand_idx.init(ctx, (size_t)nq * cap) != ASB_OK) {
        cudaGetLastError();
        return ASB_OK;
    }
    UmMaps maps;
    bool umma = um_wanted(ctx) && um_smem_bytes(k, PF_L2) <= 227 * 1024;
    if (umma && (xlo.init(ctx, (size_t)n * fp) != ASB_OK || qlo.init(ctx, (size_t)nq * fp) != ASB_OK)) {
        cudaGetLastError();
        umma = false;
    }
    if (umma) umma = um_make_maps(ctx, &maps, qf.ptr, qlo.ptr, nq, xf.ptr, xlo.ptr, n, fp);
    if (umma) um_pick_slabs(ctx, (nq + UM_TQ - 1) / UM_TQ, (n + UM_TN - 1) / UM_TN, fp, &nslabs, &tps);
    ASB_TRY(xnrm.init(ctx, (size_t)n));
    ASB_TRY(qnrm.init(ctx, (size_t)nq));
    ASB_TRY(cand_cnt.init(ctx, (size_t)nq));
    ASB_TRY(flags.init(ctx, 1));
    ASB_TRY(status.init(ctx, 1));
    ASB_TRY(gthr.init(ctx, (size_t)nq));
    ASB_TRY(diag.init(ctx, 2));
    ASB_TRY(xmax.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(status.ptr, 0, sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(diag.ptr, 0, 2 * sizeof(unsigned long long), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(xmax.ptr, 0, sizeof(unsigned long long), ctx->stream));
    pf_init_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(gthr.ptr, cand_cnt.ptr, nq);
    ASB_TRY(asb_check_launch(ctx, "pf_init_kernel"));
    pf_max_kernel<<<64, 256, 0, ctx->stream>>>(xn2_d, n, xmax.ptr);
    ASB_TRY(asb_check_launch(ctx, "pf_max_kernel"));
    if (umma) {
        UM_SPLIT_LAUNCH((unsigned)((n + 7) / 8), items_d, xn2_d, n, f, fp, xf.ptr, xlo.ptr, flags.ptr, xnrm.ptr);
        UM_SPLIT_LAUNCH((unsigned)((nq + 7) / 8), queries_d, qn2_d, nq, f, fp, qf.ptr, qlo.ptr, flags.ptr, qnrm.ptr);
    } else {
        pf_unit_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(items_d, xn2_d, n, f, fp, xf.ptr, flags.ptr, xnrm.ptr);
        pf_unit_rows_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, ctx->stream>>>(queries_d, qn2_d, nq, f, fp, qf.ptr, flags.ptr, qnrm.ptr);
    }
    ASB_TRY(asb_check_launch(ctx, "pf_unit_rows_kernel"));
    ctx->launches++;

    PfArgs A{};
    A.xf = xf.ptr;
    A.qf = qf.ptr;
    A.fp = fp;
    A.n = n;
    A.nq = nq;
    A.k = k;
    A.nslabs = nslabs;
    A.tiles_per_slab = tps;
    A.gthr = gthr.ptr;
    A.cand_cnt = cand_cnt.ptr;
    A.cand_idx = cand_idx.ptr;
    A.cand_s = cand_s.ptr;
    A.cap = cap;
    A.flags = flags.ptr;
    A.diag = diag.ptr;
    A.status = status.ptr;
    A.qn2 = qn2_d;
    A.xn2 = xn2_d;
    A.qnrm = qnrm.ptr;
    A.xnrm = xnrm.ptr;
    A.self_idx = self_idx_d;
    DevTmp<int> sel_idx_g;
    DevTmp<double> sel_s_g;
    A.maxsel = pf_maxsel(nq);
    ASB_TRY(sel_idx_g.init(ctx, (size_t)nq * A.maxsel));
    ASB_TRY(sel_s_g.init(ctx, (size_t)nq * A.maxsel));
    A.sel_idx = sel_idx_g.ptr;
    A.sel_s = sel_s_g.ptr;
    A.xn2max_bits = xmax.ptr;
    // |s~ - s| <= E(q) = 2 |q| max|x| E_cos + 1e-13 (|q|^2 + max|x|^2)  (norm rounding, the reference's own sum);  band = 2 E
    A.band_rel = 4.0 * e_cos * (1.0 + 1e-6);
    A.band_abs = 2e-13;
    if (umma) {
        ASB_TRY(um_launch<PF_L2>(ctx, maps, A, nslabs, "l2_pf_kernel"));
    } else {
        ASB_CUDA(ctx, cudaFuncSetAttribute(search_pf_kernel<PF_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt(ctx, "l2_pf_kernel");
        search_pf_kernel<PF_L2><<<dim3((unsigned)qtiles, (unsigned)nslabs), kThreads, smem, ctx->stream>>>(A);
    }
    ASB_TRY(asb_check_launch(ctx, "search_pf_kernel<L2>"));
    ASB_CUDA(ctx, cudaFuncSetAttribute(pf_finish_kernel<PF_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pf_finish_smem((int)f)));
    pf_finish_kernel<PF_L2><<<(unsigned)nq, 128, pf_finish_smem((int)f), ctx->stream>>>(A, items_d, queries_d, f, 0, 0.0, (long long *)idx_d, dist_d,
                                                                  (long long *)count_d);
    ASB_TRY(asb_check_launch(ctx, "pf_finish_kernel<L2>"));
    int hflags = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&hflags, flags.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->kernel_ms["l2_pf_flags"] = (double)hflags;
    if (hflags != 0) return ASB_OK;
    *done = true;
    return ASB_OK;
}

// Nearest item and certified distance bounds for every query, straight from the tcgen05 tile (PF_NEAR): the replay of the
// clustering walk needs, per row, a candidate centroid b, bounds dlo <= |x - c_b| <= dhi and a lower bound slo of the
// distance to every OTHER centroid -- not the exact top-2.  One pass, no candidate lists, no rescoring.  near_idx int64[nq],
// near_b f64[nq x 3] = {dlo, dhi, slo}.  *done = false: not applicable (the caller uses the FP64 kernel).
static int run_near_umma(asb_ctx *ctx, const double *items_d, long long n, int f, const double *queries_d, long long nq,
                         const double *xn2_d, const double *qn2_d, int64_t *near_idx_d, double *near_b_d, bool *done) {
    *done = false;
    if (n < 2 || n > 65536 || nq < 1 || nq > 0x7fffff00ll || !um_wanted(ctx)) return ASB_OK;
    const int fp = (f + 31) & ~31;
    if (um_smem_bytes(0, PF_NEAR) > 227 * 1024) return ASB_OK;
    const bool bf16 = um_wanted(ctx) && um_bf16(ctx);   // BF16x3 planes on the tcgen05 tile (the bound covers both tiles)
    const double e_cos = um_e_cos(fp, bf16);
    DevTmp<float> xf, qf, xlo, qlo;
    DevTmp<double> xnrm, qnrm;
    DevTmp<int> flags;
    DevTmp<unsigned long long> xmax;
    if (xf.init(ctx, (size_t)n * fp) != ASB_OK || qf.init(ctx, (size_t)nq * fp) != ASB_OK ||
        xlo.init(ctx, (size_t)n * fp) != ASB_OK || qlo.init(ctx, (size_t)nq * fp) != ASB_OK) {
        cudaGetLastError();
        return ASB_OK;
    }
    ASB_TRY(xnrm.init(ctx, (size_t)n));
    ASB_TRY(qnrm.init(ctx, (size_t)nq));
    ASB_TRY(flags.init(ctx, 1));
    ASB_TRY(xmax.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(xmax.ptr, 0, sizeof(unsigned long long), ctx->stream));
    UmMaps maps;
    if (!um_make_maps(ctx, &maps, qf.ptr, qlo.ptr, nq, xf.ptr, xlo.ptr, n, fp)) return ASB_OK;
    pf_max_kernel<<<64, 256, 0, ctx->stream>>>(xn2_d, n, xmax.ptr);
    ASB_TRY(asb_check_launch(ctx, "pf_max_kernel"));
    UM_SPLIT_LAUNCH((unsigned)((n + 7) / 8), items_d, xn2_d, n, f, fp, xf.ptr, xlo.ptr, flags.ptr, xnrm.ptr);
    UM_SPLIT_LAUNCH((unsigned)((nq + 7) / 8), queries_d, qn2_d, nq, f, fp, qf.ptr, qlo.ptr, flags.ptr, qnrm.ptr);
    ASB_TRY(asb_check_launch(ctx, "um_split_rows_kernel"));
    ctx->launches++;
    PfArgs A{};
    A.fp = fp;
    A.n = n;
    A.nq = nq;
    A.k = 0;
    A.nslabs = 1;
    A.tiles_per_slab = (n + UM_TN - 1) / UM_TN;
    A.flags = flags.ptr;
    A.qn2 = qn2_d;
    A.xn2 = xn2_d;
    A.qnrm = qnrm.ptr;
    A.xnrm = xnrm.ptr;
    A.xn2max_bits = xmax.ptr;
    A.e_cos = e_cos;
    A.near_idx = (long long *)near_idx_d;
    A.near_b = near_b_d;
    ASB_TRY(um_launch<PF_NEAR>(ctx, maps, A, 1, "cluster_top2_kernel"));
    ASB_TRY(asb_check_launch(ctx, "search_umma_kernel<NEAR>"));
    int hflags = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&hflags, flags.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // (the planes die with this scope)
    if (hflags != 0) return ASB_OK;   // a row outside the certified range: the FP64 kernel decides
    *done = true;
    return ASB_OK;
}
