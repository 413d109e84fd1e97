// laplacian.cu -- K3+K4: feature-graph Laplacian of the centroid matrix, emitted as CSR with a
// bit-exact structure.
//
// Replaces GraphFactory::build_laplacian_matrix_from_k_cluster (src/graph.rs:149-204) ->
// build_laplacian_matrix / _build_adjacency / _symmetrise_adjancency / _build_sparse_laplacian
// (src/laplacian.rs:122-417).  Nodes are the F feature columns; node i's vector is column i of
// the X x F centroid matrix (the transpose at src/graph.rs:172).
//
// The graph is tiny (F <= a few thousand nodes) but its STRUCTURE is a parity contract, so every
// comparison that decides structure (dist <= eps, kNN order) is computed with the reference's
// arithmetic: sequential sums, separate multiply and add (__dmul_rn/__dadd_rn, never fused), IEEE
// sqrt and divide.  The cosine kNN itself lives in smartcore 0.4.5 (un-vendored):
// dist = 1 - dot/(|a||b|), self excluded, ties by lower index; both are switches.
//
// Pipeline (all on device): feature norms -> F x F distance matrix -> per-row top-(topk+1)
// selection (warp arg-min rounds) + degrees -> weights / inline sparsification -> symmetric
// adjacency bitmap (atomicOr) -> row counts + scan -> CSR emit (warp per row, ballot/popc
// positions, diagonal always stored).
#include "common.cuh"

namespace {

constexpr int kMaxT1 = 64;  // topk + 1 upper bound held in registers/local arrays

// one thread per centroid: sequential, separately rounded sums over its F values (the oracle's order)
__global__ void __launch_bounds__(128) standardise_rows_kernel(const double *__restrict__ cent, int x, int f, int ddof,
                                                               double *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= x) return;
    const double *row = cent + (size_t)c * f;
    double s = 0.0, s2 = 0.0;
    for (int j = 0; j < f; ++j) {
        const double v = row[j];
        s = __dadd_rn(s, v);
        s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    const double n = (double)f;
    const double mean = __ddiv_rn(s, n);
    double var = __dsub_rn(__ddiv_rn(s2, n), __dmul_rn(mean, mean));
    if (ddof) var = __dmul_rn(var, __ddiv_rn(n, n - 1.0));
    double sd = sqrt(var);
    if (!(sd > 0.0)) sd = 1.0;
    for (int j = 0; j < f; ++j) out[(size_t)c * f + j] = __ddiv_rn(__dsub_rn(row[j], mean), sd);
}

__global__ void feat_norms_kernel(const double *__restrict__ cent, int x, int f, double *__restrict__ mag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f) return;
    double s = 0.0;
    for (int c = 0; c < x; ++c) {
        const double v = cent[(size_t)c * f + i];
        s = __dadd_rn(s, __dmul_rn(v, v));
    }
    mag[i] = __dsqrt_rn(s);
}

__global__ void feat_dist_kernel(const double *__restrict__ cent, int x, int f, const double *__restrict__ mag,
                                 int rectified, double *__restrict__ D, int *__restrict__ err) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= f) return;
    double dot = 0.0;
    for (int c = 0; c < x; ++c) {
        const double a = cent[(size_t)c * f + i];
        const double b = cent[(size_t)c * f + j];
        dot = __dadd_rn(dot, __dmul_rn(a, b));
    }
    double cs = __ddiv_rn(dot, __dmul_rn(mag[i], mag[j]));
    if (cs != cs) atomicOr(err, 1);  // zero-magnitude feature (smartcore would panic)
    if (rectified && cs < 0.0) cs = 0.0;
    D[(size_t)i * f + j] = __dsub_rn(1.0, cs);
}

// One warp per node: the t1 smallest (dist, j) in ascending order, by t1 rounds of warp arg-min.
__global__ void __launch_bounds__(256) knn_select_kernel(const double *__restrict__ D, int f, int t1,
                                                         int self_included, double eps,
                                                         int *__restrict__ knn_j, double *__restrict__ knn_d,
                                                         int *__restrict__ knn_n, int *__restrict__ deg) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= f) return;
    const double *row = D + (size_t)i * f;
    double last_d = -INFINITY;
    int last_j = -1;
    int found = 0, dg = 0;
    for (int r = 0; r < t1; ++r) {
        double bd = INFINITY;
        int bj = -1;
        for (int j = lane; j < f; j += 32) {
            if (j == i && !self_included) continue;
            const double d = row[j];
            const bool after = (d > last_d) || (d == last_d && j > last_j);
            if (!after) continue;
            if (bj < 0 || d < bd || (d == bd && j < bj)) {
                bd = d;
                bj = j;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (oj >= 0 && (bj < 0 || od < bd || (od == bd && oj < bj))) {
                bd = od;
                bj = oj;
            }
        }
        if (bj < 0) break;
        if (lane == 0) {
            knn_j[(size_t)i * t1 + r] = bj;
            knn_d[(size_t)i * t1 + r] = bd;
        }
        if (bj != i && bd <= eps) dg++;  // src/laplacian.rs:217-227
        last_d = bd;
        last_j = bj;
        found++;
    }
    if (lane == 0) {
        knn_n[i] = found;
        deg[i] = dg;
    }
}

// Single CTA: mean degree -> sparsify flag; per node the kept out-edges; symmetric bitmap.
__global__ void __launch_bounds__(1024) adjacency_kernel(int f, int t1, double eps, double sigma, double p,
                                                         const int *__restrict__ knn_j,
                                                         const double *__restrict__ knn_d,
                                                         const int *__restrict__ knn_n,
                                                         const int *__restrict__ deg, unsigned *__restrict__ bitmap,
                                                         int wpr, int *__restrict__ sparsify_out) {
    __shared__ long long sdeg[32];
    __shared__ int s_sparsify;
    long long local = 0;
    for (int i = threadIdx.x; i < f; i += blockDim.x) local += deg[i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) sdeg[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sdeg[w];
        const double avg = (double)tot / (double)f;  // src/laplacian.rs:229
        s_sparsify = avg > 10.0 ? 1 : 0;             // :230
        *sparsify_out = s_sparsify;
    }
    __syncthreads();
    const int sparsify = s_sparsify;
    for (int i = threadIdx.x; i < f; i += blockDim.x) {
        int vj[kMaxT1];
        double vs[kMaxT1];
        int nv = 0;
        const int cnt = knn_n[i];
        for (int q = 0; q < cnt; ++q) {  // :249-271
            const int j = knn_j[(size_t)i * t1 + q];
            const double d = knn_d[(size_t)i * t1 + q];
            if (j != i && d <= eps) {
                const double w = 1.0 / (1.0 + pow(d / sigma, p));
                if (w > 1e-12) {
                    vj[nv] = j;
                    vs[nv] = sparsify ? w * sqrt((double)((long long)deg[i] * (long long)deg[j])) : w;
                    nv++;
                }
            }
        }
        if (sparsify && nv > 2) {  // :274-280 ; stable insertion sort, score descending
            for (int a = 1; a < nv; ++a) {
                const int tj = vj[a];
                const double ts = vs[a];
                int b = a - 1;
                while (b >= 0 && vs[b] < ts) {
                    vj[b + 1] = vj[b];
                    vs[b + 1] = vs[b];
                    --b;
                }
                vj[b + 1] = tj;
                vs[b + 1] = ts;
            }
            int keep = nv / 2;
            if (keep < 1) keep = 1;
            nv = keep;
        }
        for (int q = 0; q < nv; ++q) {  // :317-320 union symmetrisation
            const int j = vj[q];
            atomicOr(&bitmap[(size_t)i * wpr + (j >> 5)], 1u << (j & 31));
            atomicOr(&bitmap[(size_t)j * wpr + (i >> 5)], 1u << (i & 31));
        }
    }
}

// Single CTA: row lengths (neighbours + the always-stored diagonal) -> exclusive scan -> indptr.
__global__ void __launch_bounds__(1024) row_scan_kernel(int f, const unsigned *__restrict__ bitmap, int wpr,
                                                        long long *__restrict__ indptr,
                                                        long long *__restrict__ nnz_out) {
    __shared__ long long carry;
    __shared__ long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < f; base += blockDim.x) {
        const int i = base + threadIdx.x;
        long long c = 0;
        if (i < f) {
            c = 1;  // diagonal, src/laplacian.rs:370
            for (int w = 0; w < wpr; ++w) c += __popc(bitmap[(size_t)i * wpr + w]);
        }
        long long incl = c;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            long long v = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
            long long inc2 = v;
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += u;
            }
            wsum[lane] = inc2 - v;  // exclusive prefix of warp sums
        }
        __syncthreads();
        const long long excl = carry + wsum[wid] + incl - c;
        if (i < f) indptr[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        indptr[f] = carry;
        *nnz_out = carry;
    }
}

// One warp per node: emit the CSR row in ascending column order (src/laplacian.rs:349-417):
// (i,j) = -w_ij, (i,i) = sum_j w_ij accumulated in ascending j, stored even when 0.
__global__ void __launch_bounds__(256) emit_kernel(int f, const unsigned *__restrict__ bitmap, int wpr,
                                                   const double *__restrict__ D, double sigma, double p,
                                                   const long long *__restrict__ indptr,
                                                   long long *__restrict__ indices, double *__restrict__ data) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= f) return;
    const long long pos0 = indptr[i];
    long long running = 0;  // neighbours emitted so far (all lanes agree)
    for (int wb = 0; wb < wpr; wb += 32) {
        const int w = wb + lane;
        unsigned bits = (w < wpr) ? bitmap[(size_t)i * wpr + w] : 0u;
        const int c = __popc(bits);
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        long long rank = running + incl - c;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int j = w * 32 + b;
            const double wgt = 1.0 / (1.0 + pow(D[(size_t)i * f + j] / sigma, p));
            const long long o = pos0 + rank + (j > i ? 1 : 0);  // the diagonal sits before the first j > i
            indices[o] = j;
            data[o] = -wgt;
            rank++;
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    if (lane == 0) {
        const long long len = indptr[i + 1] - pos0;
        // diagonal position = number of neighbours with j < i
        long long dpos = 0;
        for (int w = 0; w < wpr; ++w) {
            unsigned bits = bitmap[(size_t)i * wpr + w];
            if (w * 32 + 31 < i) dpos += __popc(bits);
            else if (w * 32 <= i) dpos += __popc(bits & ((1u << (i & 31)) - 1u));
        }
        double degree = 0.0;
        for (long long e = 0; e < len; ++e) {
            if (e == dpos) continue;
            degree = __dadd_rn(degree, -data[pos0 + e]);  // ascending j, :369
        }
        indices[pos0 + dpos] = i;
        data[pos0 + dpos] = degree;
    }
}

}  // namespace

int asb_dev_laplacian(asb_ctx *ctx, const double *centroids_d, int64_t x, int64_t f, const asb_graph_params &gp,
                      int64_t *indptr_d, int64_t *indices_d, double *data_d, int64_t capacity,
                      int64_t *nnz_host) {
    if (x < 2 || f < 2)
        ASB_FAIL(ctx, ASB_ERR_SHAPE, "items should be at least of shape (2,2): (%lld,%lld)", (long long)f,
                 (long long)x);  // src/laplacian.rs:129-134
    if (gp.topk < 0 || gp.topk + 1 > kMaxT1)
        ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "topk=%lld outside 0..%d", (long long)gp.topk, kMaxT1 - 1);
    if (f > 32768) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "feature graph with %lld nodes", (long long)f);
    const int fi = (int)f, xi = (int)x;
    const int t1 = (int)gp.topk + 1;  // src/laplacian.rs:211
    const double sigma = gp.has_sigma ? gp.sigma : 1.0;  // :254
    const int wpr = (fi + 31) / 32;

    DevTmp<double> mag, D, knn_d;
    DevTmp<int> knn_j, knn_n, deg, flags;
    DevTmp<unsigned> bitmap;
    DevTmp<long long> nnz_d;
    ASB_TRY(mag.init(ctx, f));
    ASB_TRY(D.init(ctx, (size_t)f * f));
    ASB_TRY(knn_d.init(ctx, (size_t)f * t1));
    ASB_TRY(knn_j.init(ctx, (size_t)f * t1));
    ASB_TRY(knn_n.init(ctx, f));
    ASB_TRY(deg.init(ctx, f));
    ASB_TRY(flags.init(ctx, 2));
    ASB_TRY(bitmap.init(ctx, (size_t)f * wpr));
    ASB_TRY(nnz_d.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, 2 * sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(bitmap.ptr, 0, (size_t)f * wpr * sizeof(unsigned), ctx->stream));

    // normalise (src/laplacian.rs:146-151): StandardScaler over the F x X matrix handed to build_laplacian_matrix, i.e.
    // every COLUMN of it -- every centroid, across its F feature values -- is shifted to mean 0 and scaled to unit
    // standard deviation.  smartcore 0.4.5 is not in the mount: mean = sum / F (features ascending), variance =
    // sum(x^2) / F - mean^2 (normalise = 1, population form) or the same times F / (F - 1) (normalise = 2), a zero
    // deviation leaves the column unscaled -- the oracle restates exactly this, the convention itself is "unpinned".
    DevTmp<double> scaled;
    if (gp.normalise) {
        ASB_TRY(scaled.init(ctx, (size_t)x * f));
        standardise_rows_kernel<<<(xi + 127) / 128, 128, 0, ctx->stream>>>(centroids_d, xi, fi, gp.normalise == 2 ? 1 : 0,
                                                                          scaled.ptr);
        ASB_TRY(asb_check_launch(ctx, "standardise_rows_kernel"));
        centroids_d = scaled.ptr;
    }
    feat_norms_kernel<<<(fi + 127) / 128, 128, 0, ctx->stream>>>(centroids_d, xi, fi, mag.ptr);
    ASB_TRY(asb_check_launch(ctx, "feat_norms_kernel"));
    feat_dist_kernel<<<dim3((fi + 127) / 128, fi), 128, 0, ctx->stream>>>(centroids_d, xi, fi, mag.ptr,
                                                                          gp.rectified, D.ptr, flags.ptr);
    ASB_TRY(asb_check_launch(ctx, "feat_dist_kernel"));
    knn_select_kernel<<<(fi + 7) / 8, 256, 0, ctx->stream>>>(D.ptr, fi, t1, gp.self_included, gp.eps, knn_j.ptr,
                                                             knn_d.ptr, knn_n.ptr, deg.ptr);
    ASB_TRY(asb_check_launch(ctx, "knn_select_kernel"));
    adjacency_kernel<<<1, 1024, 0, ctx->stream>>>(fi, t1, gp.eps, sigma, gp.p, knn_j.ptr, knn_d.ptr, knn_n.ptr,
                                                  deg.ptr, bitmap.ptr, wpr, flags.ptr + 1);
    ASB_TRY(asb_check_launch(ctx, "adjacency_kernel"));
    row_scan_kernel<<<1, 1024, 0, ctx->stream>>>(fi, bitmap.ptr, wpr, (long long *)indptr_d, nnz_d.ptr);
    ASB_TRY(asb_check_launch(ctx, "row_scan_kernel"));
    long long nnz = 0;
    int hflags[2] = {0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(&nnz, nnz_d.ptr, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(hflags, flags.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hflags[0])
        ASB_FAIL(ctx, ASB_ERR_ZERO_NORM, "cosine kNN: a feature column has zero magnitude across all centroids");
    *nnz_host = nnz;
    if (nnz > capacity)
        ASB_FAIL(ctx, ASB_ERR_CAPACITY, "laplacian: nnz=%lld exceeds capacity=%lld", nnz, (long long)capacity);
    emit_kernel<<<(fi + 7) / 8, 256, 0, ctx->stream>>>(fi, bitmap.ptr, wpr, D.ptr, sigma, gp.p,
                                                       (const long long *)indptr_d, (long long *)indices_d, data_d);
    ASB_TRY(asb_check_launch(ctx, "emit_kernel"));
    if (gp.sparsity_check) {  // src/graph.rs:185-193
        const double sparsity = 1.0 - (double)nnz / ((double)f * (double)f);
        if (sparsity > 0.95) {
            ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ASB_FAIL(ctx, ASB_ERR_TOO_SPARSE, "Resulting laplacian matrix is too sparse %g", sparsity);
        }
    }
    return ASB_OK;
}
