// extras.cu -- "next" rows of SURVEY 8f (rank 1): range_search (src/core.rs:944-976) as an ordered
// stream compaction over the lambda array (HBM bound: 8 B read per item, 16 B written per hit).
#include "common.cuh"

namespace {

constexpr int kItemsPerBlock = 4096;  // 256 threads x 16

// pass 1: hits per block ; pass 2 (WRITE): ordered emit using the scanned block offsets.
template <bool WRITE>
__global__ void __launch_bounds__(256) range_kernel(const double *__restrict__ lam, long long n, double lq, double eps,
                                                    long long *__restrict__ block_counts,
                                                    const long long *__restrict__ block_offsets,
                                                    long long index_offset, long long *__restrict__ idx,
                                                    double *__restrict__ dist) {
    __shared__ int warp_counts[8];
    __shared__ int warp_base[8];
    const long long base = (long long)blockIdx.x * kItemsPerBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long out = WRITE ? block_offsets[blockIdx.x] : 0;
    int total = 0;
    // warp w owns the contiguous 512-item slice [base + 512 w, +512): order is preserved
    for (int it = 0; it < 16; ++it) {
        const long long i = base + (long long)warp * 512 + it * 32 + lane;
        double d = 0.0;
        bool hit = false;
        if (i < n) {
            d = lq - lam[i];       // signed difference, as written in the reference (:962)
            hit = d <= eps;        // :963
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (WRITE) {
            // all warps need the counts of the warps before them: computed once below (first iteration)
            if (it == 0) {
                // count this warp's hits over its whole slice first
                int c = 0;
                for (int jt = 0; jt < 16; ++jt) {
                    const long long j = base + (long long)warp * 512 + jt * 32 + lane;
                    const bool h = (j < n) && ((lq - lam[j]) <= eps);
                    c += __popc(__ballot_sync(0xffffffffu, h));
                }
                if (lane == 0) warp_counts[warp] = c;
                __syncthreads();
                if (threadIdx.x == 0) {
                    int acc = 0;
                    for (int w = 0; w < 8; ++w) {
                        warp_base[w] = acc;
                        acc += warp_counts[w];
                    }
                }
                __syncthreads();
                out += warp_base[warp];
            }
            if (hit) {
                const long long o = out + __popc(m & ((1u << lane) - 1u));
                idx[o] = i + index_offset;
                dist[o] = d;
            }
            out += __popc(m);
        } else {
            total += __popc(m);
        }
    }
    if (!WRITE) {
        if (lane == 0) warp_counts[warp] = total;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long acc = 0;
            for (int w = 0; w < 8; ++w) acc += warp_counts[w];
            block_counts[blockIdx.x] = acc;
        }
    }
}

}  // namespace

// lambdas_d: device; outputs device (capacity entries).  Returns the number of hits in *count_host;
// when it exceeds capacity nothing is written and ASB_ERR_CAPACITY is returned.
int asb_dev_range_search(asb_ctx *ctx, const double *lambdas_d, int64_t n, double lambda_q, double eps,
                         int64_t index_offset, int64_t *idx_d, double *dist_d, int64_t capacity,
                         int64_t *count_host) {
    if (n <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "range_search: n<=0");
    const long long nblocks = (n + kItemsPerBlock - 1) / kItemsPerBlock;
    DevTmp<long long> counts, offsets;
    ASB_TRY(counts.init(ctx, (size_t)nblocks));
    ASB_TRY(offsets.init(ctx, (size_t)nblocks));
    range_kernel<false><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(lambdas_d, (long long)n, lambda_q, eps, counts.ptr,
                                                                   nullptr, 0, nullptr, nullptr);
    ASB_TRY(asb_check_launch(ctx, "range_kernel<count>"));
    std::vector<long long> hc((size_t)nblocks), ho((size_t)nblocks);
    ASB_CUDA(ctx, cudaMemcpyAsync(hc.data(), counts.ptr, nblocks * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    long long acc = 0;
    for (long long b = 0; b < nblocks; ++b) {
        ho[b] = acc;
        acc += hc[b];
    }
    *count_host = acc;
    if (acc > capacity) ASB_FAIL(ctx, ASB_ERR_CAPACITY, "range_search: %lld hits exceed capacity %lld", acc, (long long)capacity);
    if (acc == 0) return ASB_OK;
    ASB_CUDA(ctx, cudaMemcpyAsync(offsets.ptr, ho.data(), nblocks * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    {
        KernelTimer kt(ctx, "range_kernel");
        range_kernel<true><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(lambdas_d, (long long)n, lambda_q, eps, nullptr,
                                                                      offsets.ptr, (long long)index_offset,
                                                                      (long long *)idx_d, dist_d);
    }
    ASB_TRY(asb_check_launch(ctx, "range_kernel<write>"));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host offsets vector goes out of scope
    return ASB_OK;
}

// ---- JL projection with a materialised matrix -------------------------------------------
// ImplicitProjection::project (src/reduction.rs:180-199) regenerates the F x r Gaussian matrix from its seed on
// every call (ChaCha8 + StandardNormal, third-party); the host hands the matrix over instead, in the order the
// reference draws it: G[j * r + k] is the sample used for (feature j, output k).  Every output is the reference's
// own chain  y_k = (...((x_0 g_0k) s + (x_1 g_1k) s) + ...)  with s = 1 / sqrt(r), features in ascending order and
// separately rounded operations, so the result is bit-identical.  One thread per output column, kRowsPerBlock rows
// per CTA staged in shared memory; G (a few hundred kB) is served by L2.
namespace {
constexpr int kProjRows = 8;
__global__ void __launch_bounds__(256) project_kernel(const double *__restrict__ rows, long long n, int f,
                                                      const double *__restrict__ proj, int r, double scale,
                                                      double *__restrict__ out) {
    extern __shared__ double xs[];  // kProjRows x f
    const long long row0 = (long long)blockIdx.x * kProjRows;
    const int nrows = (int)((n - row0) < kProjRows ? (n - row0) : kProjRows);
    for (int e = threadIdx.x; e < nrows * f; e += blockDim.x) xs[e] = rows[row0 * f + e];
    __syncthreads();
    for (int k = threadIdx.x; k < r; k += blockDim.x) {
        double acc[kProjRows];
#pragma unroll
        for (int i = 0; i < kProjRows; ++i) acc[i] = 0.0;
        for (int j = 0; j < f; ++j) {
            const double g = proj[(size_t)j * r + k];
#pragma unroll
            for (int i = 0; i < kProjRows; ++i)
                if (i < nrows) acc[i] = __dadd_rn(acc[i], __dmul_rn(__dmul_rn(xs[i * f + j], g), scale));  // :195
        }
#pragma unroll
        for (int i = 0; i < kProjRows; ++i)
            if (i < nrows) out[(row0 + i) * r + k] = acc[i];
    }
}
}  // namespace

int asb_dev_project(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, const double *proj_d, int64_t r,
                    double *out_d) {
    if (n <= 0 || f <= 0 || r <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "project: bad sizes");
    const size_t smem = (size_t)kProjRows * f * sizeof(double);
    if (smem > 200 * 1024) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "project: f=%lld too wide", (long long)f);
    ASB_CUDA(ctx, cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double scale = 1.0 / sqrt((double)r);  // src/reduction.rs:184
    {
        KernelTimer kt(ctx, "project_kernel");
        project_kernel<<<(unsigned)((n + kProjRows - 1) / kProjRows), 256, smem, ctx->stream>>>(
            rows_d, (long long)n, (int)f, proj_d, (int)r, scale, out_d);
    }
    return asb_check_launch(ctx, "project_kernel");
}

// ---- helpers of the row-sharded entry points (api.cu) ------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) mark_tail_kernel(long long *__restrict__ idx, const long long *__restrict__ cnt,
                                                        long long nq, int k) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * k) return;
    const long long q = t / k;
    if ((t - q * k) >= cnt[q]) idx[t] = -1;
}
// sample t: the row (when this shard owns it, zeros otherwise) and its local index (or -1)
__global__ void __launch_bounds__(128) twonn_gather_kernel(const double *__restrict__ rows, long long n_local, int f,
                                                           long long offset, const long long *__restrict__ sample,
                                                           long long s, double *__restrict__ q, long long *__restrict__ self) {
    const long long t = blockIdx.x;
    if (t >= s) return;
    const long long g = sample[t] - offset;
    const bool mine = g >= 0 && g < n_local;
    for (int j = threadIdx.x; j < f; j += blockDim.x) q[t * f + j] = mine ? rows[g * f + j] : 0.0;
    if (threadIdx.x == 0) self[t] = mine ? g : -1;
}
// two smallest of the 2 R per-shard distances of every sample
__global__ void __launch_bounds__(256) twonn_merge_kernel(const double *__restrict__ all, int parts, long long s,
                                                          double *__restrict__ d1, double *__restrict__ d2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s) return;
    double m1 = INFINITY, m2 = INFINITY;
    for (int p = 0; p < parts; ++p)
        for (int w = 0; w < 2; ++w) {
            const double d = all[(size_t)p * 2 * s + (size_t)w * s + t];
            if (d < m1) {
                m2 = m1;
                m1 = d;
            } else if (d < m2) {
                m2 = d;
            }
        }
    d1[t] = m1;
    d2[t] = m2;
}
}  // namespace

int asb_dev_mark_tail(asb_ctx *ctx, int64_t *idx_d, const int64_t *cnt_d, int64_t nq, int64_t k) {
    mark_tail_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, ctx->stream>>>((long long *)idx_d, (const long long *)cnt_d,
                                                                              (long long)nq, (int)k);
    return asb_check_launch(ctx, "mark_tail_kernel");
}
int asb_dev_twonn_gather(asb_ctx *ctx, const double *rows_d, int64_t n_local, int64_t f, int64_t offset,
                         const int64_t *sample_d, int64_t s, double *q_d, int64_t *self_d) {
    twonn_gather_kernel<<<(unsigned)s, 128, 0, ctx->stream>>>(rows_d, (long long)n_local, (int)f, (long long)offset,
                                                              (const long long *)sample_d, (long long)s, q_d,
                                                              (long long *)self_d);
    return asb_check_launch(ctx, "twonn_gather_kernel");
}
int asb_dev_twonn_merge(asb_ctx *ctx, const double *all_d, int parts, int64_t s, double *d1_d, double *d2_d) {
    twonn_merge_kernel<<<(unsigned)((s + 255) / 256), 256, 0, ctx->stream>>>(all_d, parts, (long long)s, d1_d, d2_d);
    return asb_check_launch(ctx, "twonn_merge_kernel");
}

// ---- energy search through a projection / the spectral signals (api.cu: asb_index_search_energy) ------------------------
namespace {
__global__ void __launch_bounds__(256) nonfinite_rows_kernel(const double *__restrict__ rows, long long total, int *flag) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) bad |= !(fabs(rows[i]) < INFINITY);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
// out[i] = S x_i for every row x_i (S: d x d CSR); one warp per row, the row staged in shared memory, lanes over the
// rows of S, each summing val * x[col] in stored order (the order of projected_dirichlet, src/energymaps.rs:869-876)
__global__ void __launch_bounds__(128) csr_apply_rows_kernel(const long long *__restrict__ indptr,
                                                             const long long *__restrict__ indices,
                                                             const double *__restrict__ data, int d,
                                                             const double *__restrict__ rows, long long n,
                                                             double *__restrict__ out) {
    extern __shared__ double xs_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *xs = xs_all + (size_t)warp * d;
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (i >= n) return;
    for (int j = lane; j < d; j += 32) xs[j] = rows[i * d + j];
    __syncwarp();
    for (int r = lane; r < d; r += 32) {
        double sum = 0.0;
        for (long long e = indptr[r]; e < indptr[r + 1]; ++e) sum = __dadd_rn(sum, __dmul_rn(data[e], xs[indices[e]]));
        out[i * d + r] = sum;
    }
}
// second pass of the energy search: exact scores of the kc candidates from the DIFFERENCE vector (ProjectedEnergy::score,
// src/energymaps.rs:884-894): diff = q' - x' element-wise, then |S diff| (signals) or |diff|, bounded, blended with
// |lambda_q - lambda_i|; best k by (score desc, index asc).  One warp per query, kc <= 64.
__global__ void __launch_bounds__(128) energy_rescore_ex_kernel(
    const double *__restrict__ items, const double *__restrict__ lambdas, int d, const double *__restrict__ queries,
    const double *__restrict__ lambda_q, long long nq, int kc, int k, double w_lambda, double w_dir,
    const long long *__restrict__ sig_indptr, const long long *__restrict__ sig_indices, const double *__restrict__ sig_data,
    const long long *__restrict__ cand_idx, const long long *__restrict__ cand_cnt, long long *__restrict__ idx_out,
    double *__restrict__ score_out, long long *__restrict__ count_out) {
    extern __shared__ double diff_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *diff = diff_all + (size_t)warp * d;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (w >= nq) return;
    const int cnt = (int)cand_cnt[w];
    const double *q = queries + w * (long long)d;
    const double lq = lambda_q[w];
    double sc[2] = {-INFINITY, -INFINITY};
    long long id[2] = {-1, -1};
    for (int c = 0; c < cnt; ++c) {
        const long long gi = cand_idx[w * kc + c];
        const double *x = items + gi * (long long)d;
        __syncwarp();
        double acc = 0.0;
        for (int t = lane; t < d; t += 32) {
            const double df = q[t] - x[t];
            diff[t] = df;
            acc = fma(df, df, acc);
        }
        __syncwarp();
        if (sig_indptr) {
            acc = 0.0;
            for (int r = lane; r < d; r += 32) {
                double sum = 0.0;
                for (long long e = sig_indptr[r]; e < sig_indptr[r + 1]; ++e)
                    sum = __dadd_rn(sum, __dmul_rn(sig_data[e], diff[sig_indices[e]]));
                acc = fma(sum, sum, acc);
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double dn = sqrt(acc);
        const double s = -(w_lambda * fabs(lq - lambdas[gi]) + w_dir * fmin(dn / (1.0 + dn), 1.0));
        if ((c & 31) == lane) {
            sc[c >> 5] = s;
            id[c >> 5] = gi;
        }
    }
    const int kout = k < cnt ? k : cnt;
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
}  // namespace

int asb_dev_nonfinite_rows(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int *flag_d) {
    nonfinite_rows_kernel<<<256, 256, 0, ctx->stream>>>(rows_d, (long long)n * f, flag_d);
    return asb_check_launch(ctx, "nonfinite_rows_kernel");
}

int asb_dev_csr_apply_rows(asb_ctx *ctx, const int64_t *indptr_d, const int64_t *indices_d, const double *data_d, int64_t d,
                           const double *rows_d, int64_t n, double *out_d) {
    if (d * 8 * 4 > 200 * 1024) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "csr_apply_rows: dimension %lld too large", (long long)d);
    const size_t smem = (size_t)4 * d * sizeof(double);
    ASB_CUDA(ctx, cudaFuncSetAttribute(csr_apply_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_apply_rows_kernel<<<(unsigned)((n + 3) / 4), 128, smem, ctx->stream>>>((const long long *)indptr_d,
                                                                               (const long long *)indices_d, data_d, (int)d,
                                                                               rows_d, (long long)n, out_d);
    return asb_check_launch(ctx, "csr_apply_rows_kernel");
}

// rank_items / rank_queries: the rows the fused pass measures distances between (S x' / S q' with signals, else x' / q');
// items / queries: the (projected) vectors themselves, for the exact second pass
int asb_dev_search_energy_ex(asb_ctx *ctx, const double *rank_items_d, const double *items_d, const double *lambdas_d,
                             int64_t n, int64_t d, const double *rank_queries_d, const double *queries_d,
                             const double *lambda_q_d, int64_t nq, int64_t k, double w_lambda, double w_dirichlet,
                             const int64_t *sig_indptr_d, const int64_t *sig_indices_d, const double *sig_data_d,
                             int64_t *idx_d, double *score_d, int64_t *count_d, int *status_d) {
    if (n <= 0 || d <= 0 || nq <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "search_energy: empty input");
    if (k < 1 || k > 56) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "search_energy: k=%lld outside 1..56", (long long)k);
    if (d * 8 * 4 > 200 * 1024) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "search_energy: dimension %lld too large", (long long)d);
    int64_t kc = k + 4;
    if (kc > n) kc = n;
    DevTmp<int64_t> ci, cc;
    DevTmp<double> cs;
    ASB_TRY(ci.init(ctx, (size_t)nq * kc));
    ASB_TRY(cs.init(ctx, (size_t)nq * kc));
    ASB_TRY(cc.init(ctx, (size_t)nq));
    // first pass: kc candidates by the fused kernel's own (GEMM-form) energy on the ranking rows
    ASB_TRY(asb_dev_search_energy(ctx, rank_items_d, lambdas_d, nullptr, n, d, rank_queries_d, lambda_q_d, nq, kc, w_lambda,
                                  w_dirichlet, 0, ci.ptr, cs.ptr, cc.ptr, status_d));
    const size_t smem = (size_t)4 * d * sizeof(double);
    ASB_CUDA(ctx, cudaFuncSetAttribute(energy_rescore_ex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    energy_rescore_ex_kernel<<<(unsigned)((nq + 3) / 4), 128, smem, ctx->stream>>>(
        items_d, lambdas_d, (int)d, queries_d, lambda_q_d, (long long)nq, (int)kc, (int)k, w_lambda, w_dirichlet,
        (const long long *)sig_indptr_d, (const long long *)sig_indices_d, sig_data_d, (const long long *)ci.ptr,
        (const long long *)cc.ptr, (long long *)idx_d, score_d, (long long *)count_d);
    return asb_check_launch(ctx, "energy_rescore_ex_kernel");
}
