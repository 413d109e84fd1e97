// api.cu -- the C ABI (include/arrowspace_b200.h): context, host/device staging, the stage
// entry points and the device-resident index (ArrowSpaceBuilder::build, src/builder.rs:249-455).
#include <algorithm>
#include <chrono>

#include "comm.cuh"

#define ASB_VERSION_STRING "arrowspace_b200 0.1.0 (sm_100a)"

struct asb_index {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t n = 0, f = 0, x = 0, nnz = 0, max_clusters = 0;
    double radius = 0.0;
    const double *items = nullptr;  // borrowed or owned
    double *items_owned = nullptr;
    double *lambdas = nullptr, *norms2 = nullptr, *centroids = nullptr, *stats = nullptr;
    int64_t *assign = nullptr;
    unsigned long long *sizes = nullptr;
    int64_t *indptr = nullptr, *indices = nullptr;
    double *data = nullptr;
    GraphPlan plan;       // feature Laplacian: query lambdas (and item lambdas unless spectral)
    GraphPlan plan_sig;   // spectral signals (Laplacian-of-Laplacian): item lambdas when spectral
    int64_t *sig_indptr = nullptr, *sig_indices = nullptr;
    double *sig_data = nullptr;
    int64_t sig_nnz = 0;
    int tau_mode = ASB_TAU_MEDIAN;
    double tau_value = 0.0;
    double h_stats[3] = {0, 0, 0};
    double ms_cluster = 0, ms_laplacian = 0, ms_taumode = 0, ms_total = 0;
    int64_t shard_offset = 0, n_global = 0;   // row-sharded build: global index of the first local row, global row count
    // with_dims_reduction: the materialised F x r projection, r, and (lazily, for the energy search) the projected items
    double *proj = nullptr, *items_proj = nullptr;
    int64_t r = 0;
    // lazily built for the energy search with spectral signals: S (x'_i) for every item
    double *items_sig = nullptr;
};

// {min, max, sum} <-> {min, -max, sum}: one MIN all-reduce covers the first two
__global__ void stats_pack_kernel(const double *in, double *out, int /*unused*/) {
    if (threadIdx.x == 0) {
        out[0] = in[0];
        out[1] = -in[1];
        out[2] = in[2];
    }
}

extern "C" {

const char *asb_version(void) { return ASB_VERSION_STRING; }

const char *asb_status_string(int s) {
    switch (s) {
        case ASB_OK: return "ok";
        case ASB_ERR_INVALID: return "invalid argument";
        case ASB_ERR_CUDA: return "CUDA error";
        case ASB_ERR_NCCL: return "NCCL error";
        case ASB_ERR_NONFINITE_QUERY:
            return "Query item contains invalid values (NaN or infinity). All values must be finite.";
        case ASB_ERR_ZERO_LAMBDA: return "Lambda of the item is 0.0, prepare the item before searching";
        case ASB_ERR_SHAPE: return "items should be at least of shape (2,2)";
        case ASB_ERR_TOO_SPARSE: return "Resulting laplacian matrix is too sparse";
        case ASB_ERR_NO_CLUSTERS: return "No clusters created from data";
        case ASB_ERR_NAN_SCORE: return "NaN score in search (partial_cmp unwrap)";
        case ASB_ERR_ZERO_NORM: return "zero-magnitude feature column in cosine kNN";
        case ASB_ERR_EMPTY: return "items cannot be empty / cannot create a arrowspace of one arrow only";
        case ASB_ERR_DIM: return "dimension mismatch";
        case ASB_ERR_CAPACITY: return "output buffer too small";
        case ASB_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

int asb_ctx_create(int device, void *stream, asb_ctx **out) {
    if (!out) return ASB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return ASB_ERR_CUDA;  // no CPU fallback: fail loudly
    }
    if (cudaSetDevice(device) != cudaSuccess) return ASB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ASB_ERR_CUDA;
    if (prop.major < 10) return ASB_ERR_CUDA;  // built for sm_100a only
    asb_ctx *c = new asb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return ASB_ERR_CUDA;
        }
        c->own_stream = true;
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    // keep freed stream-ordered allocations cached in the pool (no per-call cudaMalloc cost)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return ASB_OK;
}

void asb_ctx_destroy(asb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->ktimers) {
        if (kv.second.a) cudaEventDestroy(kv.second.a);
        if (kv.second.b) cudaEventDestroy(kv.second.b);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *asb_last_error(asb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
int64_t asb_kernel_launches(asb_ctx *ctx) { return ctx ? ctx->launches : 0; }
double asb_last_kernel_ms(asb_ctx *ctx, const char *which) {
    if (!ctx || !which) return 0.0;
    auto kt = ctx->ktimers.find(which);
    if (kt != ctx->ktimers.end() && kt->second.b) {
        float ms = 0.f;
        if (cudaEventSynchronize(kt->second.b) == cudaSuccess &&
            cudaEventElapsedTime(&ms, kt->second.a, kt->second.b) == cudaSuccess)
            return ms;
        cudaGetLastError();
        return 0.0;
    }
    auto it = ctx->kernel_ms.find(which);
    return it == ctx->kernel_ms.end() ? 0.0 : it->second;
}

int asb_ctx_set_option(asb_ctx *ctx, const char *key, double value) {
    if (!ctx || !key) return ASB_ERR_INVALID;
    ctx->options[key] = value;
    return ASB_OK;
}

int64_t asb_laplacian_max_nnz(int64_t f, int64_t topk) { return f * (1 + 2 * (topk + 1)); }

}  // extern "C"

// ---- helpers --------------------------------------------------------------------------------

static int set_device(asb_ctx *ctx) {
    if (!ctx) return ASB_ERR_INVALID;
    ASB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ASB_OK;
}

// Fetch a CSR given as host or device pointers into host vectors (the graph is tiny).
static int csr_to_host(asb_ctx *ctx, const int64_t *indptr, const int64_t *indices, const double *data, int64_t f,
                       std::vector<int64_t> &hp, std::vector<int64_t> &hi, std::vector<double> &hd) {
    if (!indptr || f <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "csr: null indptr or f<=0");
    hp.resize((size_t)f + 1);
    if (asb_is_device_ptr(indptr)) {
        ASB_CUDA(ctx, cudaMemcpyAsync(hp.data(), indptr, (f + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        memcpy(hp.data(), indptr, (f + 1) * sizeof(int64_t));
    }
    const int64_t nnz = hp[f];
    if (nnz < 0 || nnz > ((int64_t)1 << 30)) ASB_FAIL(ctx, ASB_ERR_INVALID, "csr: bad nnz %lld", (long long)nnz);
    hi.resize((size_t)nnz);
    hd.resize((size_t)nnz);
    if (nnz > 0) {
        if (!indices || !data) ASB_FAIL(ctx, ASB_ERR_INVALID, "csr: null indices/data");
        if (asb_is_device_ptr(indices))
            ASB_CUDA(ctx, cudaMemcpyAsync(hi.data(), indices, nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        else
            memcpy(hi.data(), indices, nnz * sizeof(int64_t));
        if (asb_is_device_ptr(data))
            ASB_CUDA(ctx, cudaMemcpyAsync(hd.data(), data, nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        else
            memcpy(hd.data(), data, nnz * sizeof(double));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ASB_OK;
}

struct PlanGuard {
    GraphPlan plan;
    ~PlanGuard() { plan.release(); }
};

static int search_status_to_rc(asb_ctx *ctx, int st) {
    if (st & 2) ASB_FAIL(ctx, ASB_ERR_ZERO_LAMBDA, "Lambda of the item is 0.0, prepare the item before searching");
    if (st & 1) ASB_FAIL(ctx, ASB_ERR_NAN_SCORE, "NaN score encountered while ranking (reference panics in sort_by)");
    return ASB_OK;
}

extern "C" {

// ---- stage 1 --------------------------------------------------------------------------------

int asb_twonn_distances(asb_ctx *ctx, const double *rows, int64_t n, int64_t f, const int64_t *sample_idx,
                        int64_t s, double *d1, double *d2) {
    ASB_TRY(set_device(ctx));
    if (!rows || !sample_idx || !d1 || !d2 || n < 2 || f <= 0 || s <= 0)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "twonn: bad arguments");
    DevIn<double> r;
    DevIn<int64_t> si;
    DevOut<double> o1, o2;
    ASB_TRY(r.init(ctx, rows, (size_t)n * f));
    ASB_TRY(si.init(ctx, sample_idx, (size_t)s));
    ASB_TRY(o1.init(ctx, d1, (size_t)s));
    ASB_TRY(o2.init(ctx, d2, (size_t)s));
    StageTimer t(ctx, "twonn");
    ASB_TRY(asb_dev_twonn(ctx, r.ptr, n, f, si.ptr, s, o1.ptr, o2.ptr));
    t.stop();
    ASB_TRY(o1.finish(ctx));
    ASB_TRY(o2.finish(ctx));
    return asb_sync(ctx);
}

int asb_cluster_incremental(asb_ctx *ctx, const double *rows, int64_t n, int64_t f, int64_t max_clusters,
                            double radius, double *centroids, int64_t *assignments, uint64_t *sizes,
                            int64_t *x_out) {
    ASB_TRY(set_device(ctx));
    if (!rows || !centroids || !assignments || !sizes || !x_out)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster: null pointer");
    if (n <= 0 || f <= 0 || max_clusters <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster: bad sizes");
    DevIn<double> r;
    DevOut<double> c;
    DevOut<int64_t> a;
    DevOut<unsigned long long> sz;
    ASB_TRY(r.init(ctx, rows, (size_t)n * f));
    ASB_TRY(c.init(ctx, centroids, (size_t)max_clusters * f));
    ASB_TRY(a.init(ctx, assignments, (size_t)n));
    ASB_TRY(sz.init(ctx, (unsigned long long *)sizes, (size_t)max_clusters));
    ASB_CUDA(ctx, cudaMemsetAsync(c.ptr, 0, (size_t)max_clusters * f * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(sz.ptr, 0, (size_t)max_clusters * sizeof(unsigned long long), ctx->stream));
    StageTimer t(ctx, "cluster");
    int64_t x = 0;
    int rc = asb_dev_cluster(ctx, r.ptr, n, f, max_clusters, radius, c.ptr, a.ptr, sz.ptr, &x);
    t.stop();
    ASB_TRY(rc);
    *x_out = x;
    ASB_TRY(c.finish(ctx));
    ASB_TRY(a.finish(ctx));
    ASB_TRY(sz.finish(ctx));
    return asb_sync(ctx);
}

int asb_cluster_incremental_resume(asb_ctx *ctx, const double *rows, int64_t n, int64_t f, int64_t max_clusters,
                                   double radius, double *centroids, int64_t *assignments, uint64_t *sizes,
                                   int64_t *x_inout) {
    ASB_TRY(set_device(ctx));
    if (!rows || !centroids || !assignments || !sizes || !x_inout)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster_resume: null pointer");
    if (n <= 0 || f <= 0 || max_clusters <= 0 || *x_inout < 0 || *x_inout > max_clusters)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster_resume: bad sizes");
    DevIn<double> r;
    DevOut<int64_t> a;
    ASB_TRY(r.init(ctx, rows, (size_t)n * f));
    ASB_TRY(a.init(ctx, assignments, (size_t)n));
    // centroids / sizes are in-out: stage them both ways when they live on the host
    const bool cdev = asb_is_device_ptr(centroids), sdev = asb_is_device_ptr(sizes);
    DevTmp<double> ctmp;
    DevTmp<unsigned long long> stmp;
    double *c_d = centroids;
    unsigned long long *s_d = (unsigned long long *)sizes;
    if (!cdev) {
        ASB_TRY(ctmp.init(ctx, (size_t)max_clusters * f));
        ASB_CUDA(ctx, cudaMemcpyAsync(ctmp.ptr, centroids, (size_t)max_clusters * f * sizeof(double),
                                      cudaMemcpyHostToDevice, ctx->stream));
        c_d = ctmp.ptr;
    }
    if (!sdev) {
        ASB_TRY(stmp.init(ctx, (size_t)max_clusters));
        ASB_CUDA(ctx, cudaMemcpyAsync(stmp.ptr, sizes, (size_t)max_clusters * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                      ctx->stream));
        s_d = stmp.ptr;
    }
    StageTimer t(ctx, "cluster");
    int64_t x = 0;
    int rc = asb_dev_cluster(ctx, r.ptr, n, f, max_clusters, radius, c_d, a.ptr, s_d, &x, *x_inout);
    t.stop();
    ASB_TRY(rc);
    *x_inout = x;
    if (!cdev)
        ASB_CUDA(ctx, cudaMemcpyAsync(centroids, c_d, (size_t)max_clusters * f * sizeof(double), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    if (!sdev)
        ASB_CUDA(ctx, cudaMemcpyAsync(sizes, s_d, (size_t)max_clusters * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    ASB_TRY(a.finish(ctx));
    return asb_sync(ctx);
}

// ---- stage 2 --------------------------------------------------------------------------------

int asb_build_feature_laplacian(asb_ctx *ctx, const double *centroids, int64_t x, int64_t f,
                                const asb_graph_params *params, int64_t *indptr, int64_t *indices, double *data,
                                int64_t capacity, int64_t *nnz_out) {
    ASB_TRY(set_device(ctx));
    if (!centroids || !params || !indptr || !indices || !data || !nnz_out)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "laplacian: null pointer");
    if (x < 2 || f < 2)
        ASB_FAIL(ctx, ASB_ERR_SHAPE, "items should be at least of shape (2,2): (%lld,%lld)", (long long)f, (long long)x);
    if (capacity < f) ASB_FAIL(ctx, ASB_ERR_CAPACITY, "laplacian: capacity < f");
    DevIn<double> c;
    DevOut<int64_t> ip, ii;
    DevOut<double> dd;
    ASB_TRY(c.init(ctx, centroids, (size_t)x * f));
    ASB_TRY(ip.init(ctx, indptr, (size_t)f + 1));
    ASB_TRY(ii.init(ctx, indices, (size_t)capacity));
    ASB_TRY(dd.init(ctx, data, (size_t)capacity));
    StageTimer t(ctx, "laplacian");
    int64_t nnz = 0;
    int rc = asb_dev_laplacian(ctx, c.ptr, x, f, *params, ip.ptr, ii.ptr, dd.ptr, capacity, &nnz);
    t.stop();
    *nnz_out = nnz;
    ASB_TRY(rc);
    ASB_TRY(ip.finish(ctx));
    ASB_TRY(ii.finish(ctx, (size_t)nnz));
    ASB_TRY(dd.finish(ctx, (size_t)nnz));
    return asb_sync(ctx);
}

// ---- stage 3 --------------------------------------------------------------------------------

int asb_compute_taumode(asb_ctx *ctx, const double *items, int64_t n, int64_t f, const int64_t *indptr,
                        const int64_t *indices, const double *data, int32_t tau_mode, double tau_value,
                        double *lambdas, double *norms2, double *stats) {
    ASB_TRY(set_device(ctx));
    if (!items || !lambdas) ASB_FAIL(ctx, ASB_ERR_INVALID, "taumode: null pointer");
    if (n <= 0 || f <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "taumode: bad sizes");
    std::vector<int64_t> hp, hi;
    std::vector<double> hd;
    ASB_TRY(csr_to_host(ctx, indptr, indices, data, f, hp, hi, hd));
    PlanGuard pg;
    ASB_TRY(asb_graph_plan_from_host(ctx, hp.data(), hi.data(), hd.data(), f, &pg.plan));
    DevIn<double> it;
    DevOut<double> lam, n2;
    DevTmp<double> st;
    ASB_TRY(it.init(ctx, items, (size_t)n * f));
    ASB_TRY(lam.init(ctx, lambdas, (size_t)n));
    ASB_TRY(n2.init(ctx, norms2, norms2 ? (size_t)n : 0));
    ASB_TRY(st.init(ctx, 3));
    StageTimer t(ctx, "taumode");
    int rc = asb_dev_taumode(ctx, it.ptr, n, f, pg.plan, tau_mode, tau_value, lam.ptr, n2.ptr,
                             stats ? st.ptr : nullptr, nullptr);
    t.stop();
    ASB_TRY(rc);
    ASB_TRY(lam.finish(ctx));
    ASB_TRY(n2.finish(ctx));
    if (stats) {
        if (asb_is_device_ptr(stats))
            ASB_CUDA(ctx, cudaMemcpyAsync(stats, st.ptr, 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        else
            ASB_CUDA(ctx, cudaMemcpyAsync(stats, st.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return asb_sync(ctx);
}

int asb_prepare_query_lambdas(asb_ctx *ctx, const double *queries, int64_t nq, int64_t f, const int64_t *indptr,
                              const int64_t *indices, const double *data, int32_t tau_mode, double tau_value,
                              double *lambda_q) {
    ASB_TRY(set_device(ctx));
    if (!queries || !lambda_q) ASB_FAIL(ctx, ASB_ERR_INVALID, "prepare_query: null pointer");
    if (nq <= 0 || f <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "prepare_query: bad sizes");
    std::vector<int64_t> hp, hi;
    std::vector<double> hd;
    ASB_TRY(csr_to_host(ctx, indptr, indices, data, f, hp, hi, hd));
    PlanGuard pg;
    ASB_TRY(asb_graph_plan_from_host(ctx, hp.data(), hi.data(), hd.data(), f, &pg.plan));
    DevIn<double> q;
    DevOut<double> lam;
    DevTmp<int> flag;
    ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
    ASB_TRY(lam.init(ctx, lambda_q, (size_t)nq));
    ASB_TRY(flag.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(flag.ptr, 0, sizeof(int), ctx->stream));
    ASB_TRY(asb_dev_taumode(ctx, q.ptr, nq, f, pg.plan, tau_mode, tau_value, lam.ptr, nullptr, nullptr, flag.ptr));
    int hflag = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&hflag, flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hflag)  // src/core.rs:534-537
        ASB_FAIL(ctx, ASB_ERR_NONFINITE_QUERY,
                 "Query item contains invalid values (NaN or infinity). All values must be finite.");
    ASB_TRY(lam.finish(ctx));
    return asb_sync(ctx);
}

// ---- stage 5 --------------------------------------------------------------------------------

int asb_search_lambda_aware_batch(asb_ctx *ctx, const double *items, const double *lambdas, const double *norms2,
                                  int64_t n, int64_t f, const double *queries, const double *lambda_q, int64_t nq,
                                  int64_t k, double alpha, int64_t index_offset, int64_t *idx, double *score,
                                  int64_t *count) {
    ASB_TRY(set_device(ctx));
    if (!items || !lambdas || !queries || !lambda_q || !idx || !score)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "search: null pointer");
    if (n <= 0 || f <= 0 || nq <= 0 || k < 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "search: bad sizes");
    if (k == 0) {
        if (count) {
            if (asb_is_device_ptr(count)) ASB_CUDA(ctx, cudaMemsetAsync(count, 0, nq * sizeof(int64_t), ctx->stream));
            else memset(count, 0, nq * sizeof(int64_t));
        }
        return asb_sync(ctx);
    }
    DevIn<double> it, lam, n2, q, lq;
    DevOut<int64_t> oi, oc;
    DevOut<double> os;
    DevTmp<int64_t> cnt_tmp;
    DevTmp<int> status;
    ASB_TRY(it.init(ctx, items, (size_t)n * f));
    ASB_TRY(lam.init(ctx, lambdas, (size_t)n));
    ASB_TRY(n2.init(ctx, norms2, norms2 ? (size_t)n : 0));
    ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
    ASB_TRY(lq.init(ctx, lambda_q, (size_t)nq));
    ASB_TRY(oi.init(ctx, idx, (size_t)nq * k));
    ASB_TRY(os.init(ctx, score, (size_t)nq * k));
    ASB_TRY(oc.init(ctx, count, count ? (size_t)nq : 0));
    ASB_TRY(cnt_tmp.init(ctx, (size_t)nq));
    ASB_TRY(status.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(status.ptr, 0, sizeof(int), ctx->stream));
    StageTimer t(ctx, "search");
    int rc = asb_dev_search(ctx, it.ptr, lam.ptr, n2.ptr, n, f, q.ptr, lq.ptr, nq, k, alpha, index_offset, oi.ptr,
                            os.ptr, count ? oc.ptr : cnt_tmp.ptr, status.ptr);
    t.stop();
    ASB_TRY(rc);
    int hst = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&hst, status.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASB_TRY(search_status_to_rc(ctx, hst));
    ASB_TRY(oi.finish(ctx));
    ASB_TRY(os.finish(ctx));
    ASB_TRY(oc.finish(ctx));
    return asb_sync(ctx);
}

int asb_search_energy_batch(asb_ctx *ctx, const double *items, const double *lambdas, const double *norms2, int64_t n,
                            int64_t f, const double *queries, const double *lambda_q, int64_t nq, int64_t k,
                            double w_lambda, double w_dirichlet, int64_t index_offset, int64_t *idx, double *score,
                            int64_t *count) {
    ASB_TRY(set_device(ctx));
    if (!items || !lambdas || !queries || !lambda_q || !idx || !score)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "search_energy: null pointer");
    if (n <= 0 || f <= 0 || nq <= 0 || k < 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "search_energy: bad sizes");
    if (k == 0) {  // scored.truncate(0), src/energymaps.rs:398
        if (count) {
            if (asb_is_device_ptr(count)) ASB_CUDA(ctx, cudaMemsetAsync(count, 0, nq * sizeof(int64_t), ctx->stream));
            else memset(count, 0, nq * sizeof(int64_t));
        }
        return asb_sync(ctx);
    }
    DevIn<double> it, lam, n2, q, lq;
    DevOut<int64_t> oi, oc;
    DevOut<double> os;
    DevTmp<int64_t> cnt_tmp;
    DevTmp<int> status;
    ASB_TRY(it.init(ctx, items, (size_t)n * f));
    ASB_TRY(lam.init(ctx, lambdas, (size_t)n));
    ASB_TRY(n2.init(ctx, norms2, norms2 ? (size_t)n : 0));
    ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
    ASB_TRY(lq.init(ctx, lambda_q, (size_t)nq));
    ASB_TRY(oi.init(ctx, idx, (size_t)nq * k));
    ASB_TRY(os.init(ctx, score, (size_t)nq * k));
    ASB_TRY(oc.init(ctx, count, count ? (size_t)nq : 0));
    ASB_TRY(cnt_tmp.init(ctx, (size_t)nq));
    ASB_TRY(status.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(status.ptr, 0, sizeof(int), ctx->stream));
    StageTimer t(ctx, "search_energy");
    int rc = asb_dev_search_energy(ctx, it.ptr, lam.ptr, n2.ptr, n, f, q.ptr, lq.ptr, nq, k, w_lambda, w_dirichlet,
                                   index_offset, oi.ptr, os.ptr, count ? oc.ptr : cnt_tmp.ptr, status.ptr);
    t.stop();
    ASB_TRY(rc);
    int hst = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&hst, status.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASB_TRY(search_status_to_rc(ctx, hst));
    ASB_TRY(oi.finish(ctx));
    ASB_TRY(os.finish(ctx));
    ASB_TRY(oc.finish(ctx));
    return asb_sync(ctx);
}

int asb_search_lambda_aware_hybrid_batch(asb_ctx *ctx, const double *items, const double *lambdas, const double *norms2,
                                         int64_t n, int64_t f, const double *queries, const double *lambda_q,
                                         int64_t nq, int64_t k, double alpha, int64_t *idx, double *score,
                                         int64_t *count) {
    ASB_TRY(set_device(ctx));
    if (!items || !lambdas || !queries || !lambda_q || !idx || !score)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "hybrid search: null pointer");
    if (n <= 0 || f <= 0 || nq <= 0 || k < 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "hybrid search: bad sizes");
    std::vector<int64_t> h_idx((size_t)nq * (k ? k : 1), -1), h_cnt((size_t)nq, 0);
    std::vector<double> h_sc((size_t)nq * (k ? k : 1), 0.0);
    if (k > 0) {  // k == 0 -> empty (src/core.rs:810-812)
        DevIn<double> it, lam, n2, q, lq;
        DevTmp<int64_t> i1, i2, c1, c2;
        DevTmp<double> s1, s2, xn2;
        DevTmp<int> status;
        ASB_TRY(it.init(ctx, items, (size_t)n * f));
        ASB_TRY(lam.init(ctx, lambdas, (size_t)n));
        ASB_TRY(n2.init(ctx, norms2, norms2 ? (size_t)n : 0));
        ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
        ASB_TRY(lq.init(ctx, lambda_q, (size_t)nq));
        ASB_TRY(i1.init(ctx, (size_t)nq * k));
        ASB_TRY(i2.init(ctx, (size_t)nq * k));
        ASB_TRY(s1.init(ctx, (size_t)nq * k));
        ASB_TRY(s2.init(ctx, (size_t)nq * k));
        ASB_TRY(c1.init(ctx, (size_t)nq));
        ASB_TRY(c2.init(ctx, (size_t)nq));
        ASB_TRY(status.init(ctx, 1));
        ASB_CUDA(ctx, cudaMemsetAsync(status.ptr, 0, sizeof(int), ctx->stream));
        const double *n2p = n2.ptr;
        if (!n2p) {
            ASB_TRY(xn2.init(ctx, (size_t)n));
            ASB_TRY(asb_dev_norms2(ctx, it.ptr, n, f, xn2.ptr));
            n2p = xn2.ptr;
        }
        StageTimer t(ctx, "hybrid_search");
        ASB_TRY(asb_dev_search(ctx, it.ptr, lam.ptr, n2p, n, f, q.ptr, lq.ptr, nq, k, alpha, 0, i1.ptr, s1.ptr, c1.ptr,
                               status.ptr));
        ASB_TRY(asb_dev_search(ctx, it.ptr, lam.ptr, n2p, n, f, q.ptr, lq.ptr, nq, k, 1.0, 0, i2.ptr, s2.ptr, c2.ptr,
                               status.ptr));
        t.stop();
        std::vector<int64_t> a_i((size_t)nq * k), b_i((size_t)nq * k), a_c((size_t)nq), b_c((size_t)nq);
        std::vector<double> a_s((size_t)nq * k), b_s((size_t)nq * k);
        int hst = 0;
        ASB_CUDA(ctx, cudaMemcpyAsync(a_i.data(), i1.ptr, a_i.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(b_i.data(), i2.ptr, b_i.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(a_s.data(), s1.ptr, a_s.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(b_s.data(), s2.ptr, b_s.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(a_c.data(), c1.ptr, a_c.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(b_c.data(), c2.ptr, b_c.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(&hst, status.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (hst & 1) ASB_FAIL(ctx, ASB_ERR_NAN_SCORE, "NaN score encountered while ranking");
        // union + re-rank on the host: 2k candidates per query (src/core.rs:897-924)
        struct Cand { int64_t i; double s; };
        std::vector<Cand> u;
        for (int64_t qi = 0; qi < nq; ++qi) {
            u.clear();
            const int64_t ca = a_c[qi], cb = b_c[qi];
            auto has = [&](int64_t id) {
                for (auto &c : u) if (c.i == id) return true;
                return false;
            };
            for (int64_t r = 0; r < cb; ++r)  // high semantic matches, scored by cosine
                if (b_s[qi * k + r] > 0.9999) u.push_back({b_i[qi * k + r], b_s[qi * k + r]});
            for (int64_t r = 0; r < ca; ++r)  // lambda top-k, or_insert
                if (!has(a_i[qi * k + r])) u.push_back({a_i[qi * k + r], a_s[qi * k + r]});
            if (cb > 0 && !has(b_i[qi * k])) u.push_back({b_i[qi * k], b_s[qi * k]});  // semantic top-1
            std::sort(u.begin(), u.end(), [](const Cand &x, const Cand &y) { return x.s > y.s || (x.s == y.s && x.i < y.i); });
            const int64_t outn = (int64_t)u.size() < k ? (int64_t)u.size() : k;
            for (int64_t r = 0; r < outn; ++r) {
                h_idx[qi * k + r] = u[r].i;
                h_sc[qi * k + r] = u[r].s;
            }
            h_cnt[qi] = outn;
        }
    }
    auto put = [&](void *dst, const void *src, size_t bytes) -> int {
        if (!dst || bytes == 0) return ASB_OK;
        if (asb_is_device_ptr(dst)) {
            ASB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
            ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            memcpy(dst, src, bytes);
        }
        return ASB_OK;
    };
    if (k > 0) {
        ASB_TRY(put(idx, h_idx.data(), (size_t)nq * k * 8));
        ASB_TRY(put(score, h_sc.data(), (size_t)nq * k * 8));
    }
    ASB_TRY(put(count, h_cnt.data(), (size_t)nq * 8));
    return ASB_OK;
}

int asb_range_search(asb_ctx *ctx, const double *lambdas, int64_t n, double lambda_q, double eps, int64_t index_offset,
                     int64_t *idx, double *dist, int64_t capacity, int64_t *count_out) {
    ASB_TRY(set_device(ctx));
    if (!lambdas || !idx || !dist || !count_out) ASB_FAIL(ctx, ASB_ERR_INVALID, "range_search: null pointer");
    if (n <= 0 || capacity < 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "range_search: bad sizes");
    DevIn<double> lam;
    DevOut<int64_t> oi;
    DevOut<double> od;
    ASB_TRY(lam.init(ctx, lambdas, (size_t)n));
    ASB_TRY(oi.init(ctx, idx, (size_t)capacity));
    ASB_TRY(od.init(ctx, dist, (size_t)capacity));
    int64_t cnt = 0;
    int rc = asb_dev_range_search(ctx, lam.ptr, n, lambda_q, eps, index_offset, oi.ptr, od.ptr, capacity, &cnt);
    *count_out = cnt;
    ASB_TRY(rc);
    ASB_TRY(oi.finish(ctx, (size_t)cnt));
    ASB_TRY(od.finish(ctx, (size_t)cnt));
    return asb_sync(ctx);
}

int64_t asb_jl_dimension(int64_t n_points, double epsilon) {  // compute_jl_dimension, src/reduction.rs:127-141
    const double log_n = log((double)n_points);
    const double eps_sq = pow(epsilon, 2.0);
    const double v = ceil(8.0 * log_n / eps_sq);
    int64_t jl = (v != v || v < 0.0) ? 0 : (v > 9.0e18 ? INT64_MAX : (int64_t)v);  // `as usize` saturates
    return jl > 32 ? jl : 32;
}

int asb_project_matrix(asb_ctx *ctx, const double *rows, int64_t n, int64_t f, const double *projection, int64_t r,
                       double *out) {
    ASB_TRY(set_device(ctx));
    if (!rows || !projection || !out) ASB_FAIL(ctx, ASB_ERR_INVALID, "project_matrix: null pointer");
    if (n <= 0 || f <= 0 || r <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "project_matrix: bad sizes");
    DevIn<double> x, g;
    DevOut<double> y;
    ASB_TRY(x.init(ctx, rows, (size_t)n * f));
    ASB_TRY(g.init(ctx, projection, (size_t)f * r));
    ASB_TRY(y.init(ctx, out, (size_t)n * r));
    StageTimer t(ctx, "project");
    int rc = asb_dev_project(ctx, x.ptr, n, f, g.ptr, r, y.ptr);
    t.stop();
    ASB_TRY(rc);
    ASB_TRY(y.finish(ctx));
    return asb_sync(ctx);
}

int asb_topk_merge(asb_ctx *ctx, const double *in_score, const int64_t *in_idx, int64_t parts, int64_t nq,
                   int64_t k, double *out_score, int64_t *out_idx, int64_t *out_count) {
    ASB_TRY(set_device(ctx));
    if (!in_score || !in_idx || !out_score || !out_idx) ASB_FAIL(ctx, ASB_ERR_INVALID, "topk_merge: null pointer");
    if (parts < 1 || nq < 1 || k < 1) ASB_FAIL(ctx, ASB_ERR_INVALID, "topk_merge: bad sizes");
    DevIn<double> is;
    DevIn<int64_t> ii;
    DevOut<double> os;
    DevOut<int64_t> oi, oc;
    DevTmp<int64_t> cnt_tmp;
    ASB_TRY(is.init(ctx, in_score, (size_t)parts * nq * k));
    ASB_TRY(ii.init(ctx, in_idx, (size_t)parts * nq * k));
    ASB_TRY(os.init(ctx, out_score, (size_t)nq * k));
    ASB_TRY(oi.init(ctx, out_idx, (size_t)nq * k));
    ASB_TRY(oc.init(ctx, out_count, out_count ? (size_t)nq : 0));
    ASB_TRY(cnt_tmp.init(ctx, (size_t)nq));
    ASB_TRY(asb_dev_topk_merge(ctx, is.ptr, ii.ptr, parts, nq, k, os.ptr, oi.ptr, out_count ? oc.ptr : cnt_tmp.ptr));
    ASB_TRY(os.finish(ctx));
    ASB_TRY(oi.finish(ctx));
    ASB_TRY(oc.finish(ctx));
    return asb_sync(ctx);
}

// ---- whole build ----------------------------------------------------------------------------

void asb_index_destroy(asb_index *ix) {
    if (!ix) return;
    // Independent of the creating context (it may already be gone at interpreter shutdown):
    // cudaFree synchronises with outstanding work on the memory and accepts the stream-ordered
    // (pooled) allocations asb_index_build makes -- they go back to the device's memory pool, so the
    // next build does not pay the driver for them again.
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();
    cudaFree(ix->items_owned);
    cudaFree(ix->lambdas);
    cudaFree(ix->norms2);
    cudaFree(ix->centroids);
    cudaFree(ix->stats);
    cudaFree(ix->assign);
    cudaFree(ix->sizes);
    cudaFree(ix->indptr);
    cudaFree(ix->indices);
    cudaFree(ix->data);
    cudaFree(ix->plan.entries);
    cudaFree(ix->plan.row_ptr);
    cudaFree(ix->plan.sym_edges);
    cudaFree(ix->plan.resid);
    cudaFree(ix->plan_sig.entries);
    cudaFree(ix->plan_sig.row_ptr);
    cudaFree(ix->plan_sig.sym_edges);
    cudaFree(ix->plan_sig.resid);
    cudaFree(ix->sig_indptr);
    cudaFree(ix->sig_indices);
    cudaFree(ix->sig_data);
    cudaFree(ix->proj);
    cudaFree(ix->items_proj);
    cudaFree(ix->items_sig);
    cudaGetLastError();
    delete ix;
}

}  // extern "C"

// stages 1-3 over the local rows; comm == nullptr (or one rank): the whole dataset is local
static int index_build_impl(asb_ctx *ctx, asb_comm *comm, const double *rows, int64_t n, int64_t f, int64_t shard_offset,
                            int64_t n_global, const asb_build_params *bp, asb_index **out) {
    ASB_TRY(set_device(ctx));
    if (!rows || !bp || !out) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_build: null pointer");
    *out = nullptr;
    const bool sharded = comm && comm->nranks > 1;
    if (n_global <= 0) ASB_FAIL(ctx, ASB_ERR_EMPTY, "items cannot be empty");  // src/core.rs:416
    if (n_global <= 1) ASB_FAIL(ctx, ASB_ERR_EMPTY, "cannot create a arrowspace of one arrow only");  // :417-420
    if (n <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_build: a shard needs at least one row");
    if (f <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_build: f<=0");
    if (bp->max_clusters <= 0 || !(bp->radius >= 0.0))
        ASB_FAIL(ctx, ASB_ERR_INVALID, "index_build: max_clusters/radius must come from the host heuristic");
    asb_graph_params gp = bp->graph;
    if (bp->apply_define_result_k) {  // src/builder.rs:225-233
        if (gp.k <= 5) gp.topk = 3;
        else if (gp.k < 10) gp.topk = 4;
    }
    asb_index *ix = new asb_index();
    ix->device = ctx->device;
    ix->stream = ctx->stream;
    ix->n = n;
    ix->f = f;
    ix->shard_offset = shard_offset;
    ix->n_global = n_global;
    ix->max_clusters = bp->max_clusters;
    ix->radius = bp->radius;
    ix->tau_mode = bp->tau_mode;
    ix->tau_value = bp->tau_value;
    struct Guard {
        asb_index *p;
        ~Guard() { if (p) asb_index_destroy(p); }
    } guard{ix};

    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    struct EvGuard {
        cudaEvent_t a, b, c, d;
        ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); cudaEventDestroy(d); }
    } evg{e0, e1, e2, e3};

    // Rows in host memory travel on a second stream in chunks while stage 1 already works on the head of the matrix
    // (the replay's chunks double in size, the upload runs ahead of them): a build from host memory costs about the
    // copy itself instead of copy + build.  Pinned host memory makes the copies truly asynchronous.
    RowsInFlight upload;
    struct UploadGuard {
        asb_ctx *ctx;
        RowsInFlight *u;
        cudaStream_t s = nullptr;
        cudaEvent_t ready = nullptr;
        ~UploadGuard() {
            ctx->rows_in_flight = nullptr;
            if (s) cudaStreamSynchronize(s);
            for (cudaEvent_t e : u->events) cudaEventDestroy(e);
            if (ready) cudaEventDestroy(ready);
            if (s) cudaStreamDestroy(s);
        }
    } upg{ctx, &upload};
    if (asb_is_device_ptr(rows)) {
        ix->items = rows;
    } else {
        ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->items_owned, (size_t)n * f * sizeof(double), ctx->stream));
        ix->items = ix->items_owned;
        const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)((16u << 20) / ((size_t)f * sizeof(double))));
        bool overlap = false;
        {
            auto it = ctx->options.find("build_overlap_upload");
            overlap = (it == ctx->options.end() || it->second != 0.0) && n > 4 * rows_per_chunk;
        }
        if (overlap && cudaStreamCreateWithFlags(&upg.s, cudaStreamNonBlocking) == cudaSuccess) {
            // the allocation above is stream-ordered on the context's stream: the copies wait for it
            ASB_CUDA(ctx, cudaEventCreateWithFlags(&upg.ready, cudaEventDisableTiming));
            ASB_CUDA(ctx, cudaEventRecord(upg.ready, ctx->stream));
            ASB_CUDA(ctx, cudaStreamWaitEvent(upg.s, upg.ready, 0));
            upload.base = ix->items_owned;
            upload.f = f;
            upload.rows_per_event = rows_per_chunk;
            for (int64_t r0 = 0; r0 < n; r0 += rows_per_chunk) {
                const int64_t cnt = std::min<int64_t>(rows_per_chunk, n - r0);
                ASB_CUDA(ctx, cudaMemcpyAsync(ix->items_owned + r0 * f, rows + r0 * f, (size_t)cnt * f * sizeof(double),
                                              cudaMemcpyHostToDevice, upg.s));
                cudaEvent_t ev;
                ASB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                ASB_CUDA(ctx, cudaEventRecord(ev, upg.s));
                upload.events.push_back(ev);
            }
            ctx->rows_in_flight = &upload;
        } else {
            ASB_CUDA(ctx, cudaMemcpyAsync(ix->items_owned, rows, (size_t)n * f * sizeof(double), cudaMemcpyHostToDevice,
                                          ctx->stream));
        }
    }
    const int64_t cap = asb_laplacian_max_nnz(f, gp.topk);
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->lambdas, (size_t)n * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->norms2, (size_t)n * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->centroids, (size_t)bp->max_clusters * f * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->stats, 3 * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->assign, (size_t)n * sizeof(int64_t), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->sizes, (size_t)bp->max_clusters * sizeof(unsigned long long), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->indptr, (size_t)(f + 1) * sizeof(int64_t), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->indices, (size_t)cap * sizeof(int64_t), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->data, (size_t)cap * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(ix->centroids, 0, (size_t)bp->max_clusters * f * sizeof(double), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(ix->sizes, 0, (size_t)bp->max_clusters * sizeof(unsigned long long), ctx->stream));

    // stage 1: clustering (src/eigenmaps.rs:224-232)
    cudaEventRecord(e0, ctx->stream);
    int64_t x = 0;
    if (sharded)
        ASB_TRY(asb_dev_cluster_sharded(ctx, comm, ix->items, n, f, bp->max_clusters, bp->radius, ix->centroids, ix->assign,
                                        ix->sizes, &x));
    else
        ASB_TRY(asb_dev_cluster(ctx, ix->items, n, f, bp->max_clusters, bp->radius, ix->centroids, ix->assign, ix->sizes, &x));
    ix->x = x;
    asb_wait_rows(ctx, ix->items + n * f);   // stage 3 reads every row
    cudaEventRecord(e1, ctx->stream);
    // stage 2: feature Laplacian (src/eigenmaps.rs:313-323); assert clustered.shape().0 <= n_items holds.  Row-sharded:
    // rank 0 builds it once, the CSR is broadcast (a few hundred kB) -- every rank ends with the same bytes
    int64_t nnz = 0;
    // with_dims_reduction (src/eigenmaps.rs:248-269): the graph is built on the projected centroids, X x r -> r x r
    const int64_t fg = (bp->projection && bp->reduced_dim > 0) ? bp->reduced_dim : f;   // nodes of the feature graph
    DevTmp<double> cent_proj;
    const double *lap_in = ix->centroids;
    if (fg != f) {
        if (fg < 2 || fg > f) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_build: reduced_dim=%lld outside 2..F", (long long)fg);
        ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->proj, (size_t)f * fg * sizeof(double), ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(ix->proj, bp->projection, (size_t)f * fg * sizeof(double),
                                      asb_is_device_ptr(bp->projection) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                      ctx->stream));
        ix->r = fg;
        ASB_TRY(cent_proj.init(ctx, (size_t)x * fg));
        ASB_TRY(asb_dev_project(ctx, ix->centroids, x, f, ix->proj, fg, cent_proj.ptr));
        lap_in = cent_proj.ptr;
    }
    if (!sharded || comm->rank == 0) {
        int rc_lap = asb_dev_laplacian(ctx, lap_in, x, fg, gp, ix->indptr, ix->indices, ix->data, cap, &nnz);
        if (!sharded) ASB_TRY(rc_lap);
        else if (rc_lap != ASB_OK) nnz = -(int64_t)rc_lap;   // the status travels with the broadcast: no rank hangs
    }
    if (sharded) {
        DevTmp<long long> hdr;
        ASB_TRY(hdr.init(ctx, 1));
        long long h_nnz = nnz;
        ASB_CUDA(ctx, cudaMemcpyAsync(hdr.ptr, &h_nnz, 8, cudaMemcpyHostToDevice, ctx->stream));
        ASB_TRY(asb_comm_bcast_bytes(ctx, comm, hdr.ptr, 8, 0));
        ASB_CUDA(ctx, cudaMemcpyAsync(&h_nnz, hdr.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_nnz < 0) {
            if (comm->rank != 0) ctx->last_error = "feature Laplacian failed on rank 0";
            return (int)(-h_nnz);
        }
        nnz = h_nnz;
        ASB_TRY(asb_comm_bcast_bytes(ctx, comm, ix->indptr, (size_t)(fg + 1) * sizeof(int64_t), 0));
        ASB_TRY(asb_comm_bcast_bytes(ctx, comm, ix->indices, (size_t)nnz * sizeof(int64_t), 0));
        ASB_TRY(asb_comm_bcast_bytes(ctx, comm, ix->data, (size_t)nnz * sizeof(double), 0));
    }
    ix->nnz = nnz;
    {
        std::vector<int64_t> hp, hi;
        std::vector<double> hd;
        ASB_TRY(csr_to_host(ctx, ix->indptr, ix->indices, ix->data, fg, hp, hi, hd));
        ASB_TRY(asb_graph_plan_from_host(ctx, hp.data(), hi.data(), hd.data(), fg, &ix->plan));
        if (bp->spectral) {
            const int64_t f = fg;   // (the signals graph lives on the feature graph's nodes)
            // optional stage (src/eigenmaps.rs:325-345 -> src/graph.rs:211-231): signals = the same Laplacian
            // construction run on dense(L)^T, i.e. with the F rows of L as the "items" and its F columns as nodes
            std::vector<double> dense((size_t)f * f, 0.0);
            for (int64_t r = 0; r < f; ++r)
                for (int64_t e = hp[r]; e < hp[r + 1]; ++e) dense[(size_t)r * f + hi[e]] = hd[e];
            DevTmp<double> dense_d;
            ASB_TRY(dense_d.init(ctx, (size_t)f * f));
            ASB_CUDA(ctx, cudaMemcpyAsync(dense_d.ptr, dense.data(), (size_t)f * f * sizeof(double), cudaMemcpyHostToDevice,
                                          ctx->stream));
            ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->sig_indptr, (size_t)(f + 1) * sizeof(int64_t), ctx->stream));
            ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->sig_indices, (size_t)cap * sizeof(int64_t), ctx->stream));
            ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->sig_data, (size_t)cap * sizeof(double), ctx->stream));
            int64_t snnz = 0;
            ASB_TRY(asb_dev_laplacian(ctx, dense_d.ptr, f, f, gp, ix->sig_indptr, ix->sig_indices, ix->sig_data, cap, &snnz));
            ix->sig_nnz = snnz;
            std::vector<int64_t> sp, si;
            std::vector<double> sd;
            ASB_TRY(csr_to_host(ctx, ix->sig_indptr, ix->sig_indices, ix->sig_data, f, sp, si, sd));
            ASB_TRY(asb_graph_plan_from_host(ctx, sp.data(), si.data(), sd.data(), f, &ix->plan_sig));
        }
    }
    cudaEventRecord(e2, ctx->stream);
    // stage 3: taumode (src/eigenmaps.rs:358-383); the items read the signals graph when it exists
    // (src/taumode.rs:195-200), queries never do (src/core.rs:548)
    ASB_TRY(asb_dev_taumode(ctx, ix->items, n, f, bp->spectral ? ix->plan_sig : ix->plan, bp->tau_mode, bp->tau_value,
                            ix->lambdas, ix->norms2, ix->stats, nullptr));
    cudaEventRecord(e3, ctx->stream);
    if (sharded) {   // global lambda statistics (src/eigenmaps.rs:372-382): min, max, sum over all shards
        DevTmp<double> st;
        ASB_TRY(st.init(ctx, 3));
        stats_pack_kernel<<<1, 32, 0, ctx->stream>>>(ix->stats, st.ptr, 0);
        ASB_TRY(asb_comm_allreduce_f64(ctx, comm, st.ptr, 2, ASB_RED_MIN));      // {min, -max}
        ASB_TRY(asb_comm_allreduce_f64(ctx, comm, st.ptr + 2, 1, ASB_RED_SUM));
        stats_pack_kernel<<<1, 32, 0, ctx->stream>>>(st.ptr, ix->stats, 1);
        ASB_TRY(asb_check_launch(ctx, "stats_pack_kernel"));
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // st dies with this scope
    }
    ASB_CUDA(ctx, cudaMemcpyAsync(ix->h_stats, ix->stats, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1); ix->ms_cluster = ms;
    cudaEventElapsedTime(&ms, e1, e2); ix->ms_laplacian = ms;
    cudaEventElapsedTime(&ms, e2, e3); ix->ms_taumode = ms;
    cudaEventElapsedTime(&ms, e0, e3); ix->ms_total = ms;
    ctx->kernel_ms["build_cluster"] = ix->ms_cluster;
    ctx->kernel_ms["build_laplacian"] = ix->ms_laplacian;
    ctx->kernel_ms["build_taumode"] = ix->ms_taumode;
    ctx->kernel_ms["build_total"] = ix->ms_total;
    guard.p = nullptr;
    *out = ix;
    return ASB_OK;
}

extern "C" {

int asb_index_build(asb_ctx *ctx, const double *rows, int64_t n, int64_t f, const asb_build_params *bp,
                    asb_index **out) {
    return index_build_impl(ctx, nullptr, rows, n, f, 0, n, bp, out);
}

int asb_index_build_sharded(asb_ctx *ctx, asb_comm *comm, const double *rows_local, int64_t n_local, int64_t f,
                            int64_t shard_offset, int64_t n_global, const asb_build_params *bp, asb_index **out) {
    if (!ctx || !comm) return ASB_ERR_INVALID;
    return index_build_impl(ctx, comm, rows_local, n_local, f, shard_offset, n_global, bp, out);
}

int asb_index_info_get(const asb_index *ix, asb_index_info *info) {
    if (!ix || !info) return ASB_ERR_INVALID;
    info->n_items = ix->n;
    info->n_features = ix->f;
    info->n_clusters = ix->x;
    info->nnz = ix->nnz;
    info->lambda_min = ix->h_stats[0];
    info->lambda_max = ix->h_stats[1];
    info->lambda_sum = ix->h_stats[2];
    info->radius = ix->radius;
    info->max_clusters = ix->max_clusters;
    info->ms_cluster = ix->ms_cluster;
    info->ms_laplacian = ix->ms_laplacian;
    info->ms_taumode = ix->ms_taumode;
    info->ms_total = ix->ms_total;
    info->nnz_signals = ix->sig_nnz;
    return ASB_OK;
}

static int copy_out(asb_ctx *ctx, void *dst, const void *src_d, size_t bytes) {
    if (!dst) ASB_FAIL(ctx, ASB_ERR_INVALID, "copy_out: null destination");
    if (bytes == 0) return ASB_OK;
    ASB_CUDA(ctx, cudaMemcpyAsync(dst, src_d, bytes, asb_is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                  ctx->stream));
    return asb_sync(ctx);
}

int asb_index_lambdas(asb_ctx *ctx, const asb_index *ix, double *dst) {
    ASB_TRY(set_device(ctx));
    if (!ix) ASB_FAIL(ctx, ASB_ERR_INVALID, "null index");
    return copy_out(ctx, dst, ix->lambdas, (size_t)ix->n * sizeof(double));
}
int asb_index_centroids(asb_ctx *ctx, const asb_index *ix, double *dst) {
    ASB_TRY(set_device(ctx));
    if (!ix) ASB_FAIL(ctx, ASB_ERR_INVALID, "null index");
    return copy_out(ctx, dst, ix->centroids, (size_t)ix->x * ix->f * sizeof(double));
}
int asb_index_assignments(asb_ctx *ctx, const asb_index *ix, int64_t *dst) {
    ASB_TRY(set_device(ctx));
    if (!ix) ASB_FAIL(ctx, ASB_ERR_INVALID, "null index");
    return copy_out(ctx, dst, ix->assign, (size_t)ix->n * sizeof(int64_t));
}
int asb_index_cluster_sizes(asb_ctx *ctx, const asb_index *ix, uint64_t *dst) {
    ASB_TRY(set_device(ctx));
    if (!ix) ASB_FAIL(ctx, ASB_ERR_INVALID, "null index");
    return copy_out(ctx, dst, ix->sizes, (size_t)ix->x * sizeof(uint64_t));
}
int asb_index_laplacian(asb_ctx *ctx, const asb_index *ix, int64_t *indptr, int64_t *indices, double *data) {
    ASB_TRY(set_device(ctx));
    if (!ix) ASB_FAIL(ctx, ASB_ERR_INVALID, "null index");
    ASB_TRY(copy_out(ctx, indptr, ix->indptr, (size_t)((ix->r ? ix->r : ix->f) + 1) * sizeof(int64_t)));
    ASB_TRY(copy_out(ctx, indices, ix->indices, (size_t)ix->nnz * sizeof(int64_t)));
    return copy_out(ctx, data, ix->data, (size_t)ix->nnz * sizeof(double));
}

int asb_index_signals(asb_ctx *ctx, const asb_index *ix, int64_t *indptr, int64_t *indices, double *data) {
    ASB_TRY(set_device(ctx));
    if (!ix) ASB_FAIL(ctx, ASB_ERR_INVALID, "null index");
    if (!ix->sig_indptr) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_signals: the index was built without spectral signals");
    ASB_TRY(copy_out(ctx, indptr, ix->sig_indptr, (size_t)((ix->r ? ix->r : ix->f) + 1) * sizeof(int64_t)));
    ASB_TRY(copy_out(ctx, indices, ix->sig_indices, (size_t)ix->sig_nnz * sizeof(int64_t)));
    return copy_out(ctx, data, ix->sig_data, (size_t)ix->sig_nnz * sizeof(double));
}

int asb_index_search(asb_ctx *ctx, const asb_index *ix, const double *queries, int64_t nq, int64_t k, double alpha,
                     int64_t *idx, double *score, int64_t *count, double *lambda_q_out) {
    ASB_TRY(set_device(ctx));
    if (!ix || !queries || !idx || !score) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_search: null pointer");
    if (nq <= 0 || k < 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_search: bad sizes");
    if (ix->r)   // EigenMaps::search hands the PROJECTED query to lambda_similarity against raw items: src/core.rs:157-161
        ASB_FAIL(ctx, ASB_ERR_DIM, "items should be of the same length");
    const int64_t f = ix->f;
    DevIn<double> q;
    DevTmp<double> lq;
    DevTmp<int> flags;
    DevTmp<int64_t> cnt_tmp;
    DevOut<int64_t> oi, oc;
    DevOut<double> os;
    ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
    ASB_TRY(lq.init(ctx, (size_t)nq));
    ASB_TRY(flags.init(ctx, 2));
    ASB_TRY(cnt_tmp.init(ctx, (size_t)nq));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, 2 * sizeof(int), ctx->stream));
    // prepare_query_item (src/core.rs:533-549): tau from the query values, lambda on gl.matrix
    StageTimer tq(ctx, "query_lambda");
    ASB_TRY(asb_dev_taumode(ctx, q.ptr, nq, f, ix->plan, ix->tau_mode, ix->tau_value, lq.ptr, nullptr, nullptr, flags.ptr));
    tq.stop();
    int h[2] = {0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(h, flags.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[0])
        ASB_FAIL(ctx, ASB_ERR_NONFINITE_QUERY,
                 "Query item contains invalid values (NaN or infinity). All values must be finite.");
    if (lambda_q_out) ASB_TRY(copy_out(ctx, lambda_q_out, lq.ptr, (size_t)nq * sizeof(double)));
    if (k == 0) {
        if (count) {
            if (asb_is_device_ptr(count)) ASB_CUDA(ctx, cudaMemsetAsync(count, 0, nq * sizeof(int64_t), ctx->stream));
            else memset(count, 0, nq * sizeof(int64_t));
        }
        return asb_sync(ctx);
    }
    ASB_TRY(oi.init(ctx, idx, (size_t)nq * k));
    ASB_TRY(os.init(ctx, score, (size_t)nq * k));
    ASB_TRY(oc.init(ctx, count, count ? (size_t)nq : 0));
    StageTimer ts(ctx, "search");
    int rc = asb_dev_search(ctx, ix->items, ix->lambdas, ix->norms2, ix->n, f, q.ptr, lq.ptr, nq, k, alpha, 0, oi.ptr,
                            os.ptr, count ? oc.ptr : cnt_tmp.ptr, flags.ptr + 1);
    ts.stop();
    ASB_TRY(rc);
    ASB_CUDA(ctx, cudaMemcpyAsync(h + 1, flags.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASB_TRY(search_status_to_rc(ctx, h[1]));
    ASB_TRY(oi.finish(ctx));
    ASB_TRY(os.finish(ctx));
    ASB_TRY(oc.finish(ctx));
    return asb_sync(ctx);
}

int asb_index_search_lambda_aware(asb_ctx *ctx, const asb_index *ix, const double *queries, const double *lambda_q,
                                  int64_t nq, int64_t k, double alpha, int64_t *idx, double *score, int64_t *count) {
    if (!ix) return ASB_ERR_INVALID;
    return asb_search_lambda_aware_batch(ctx, ix->items, ix->lambdas, ix->norms2, ix->n, ix->f, queries, lambda_q, nq,
                                         k, alpha, 0, idx, score, count);
}


// ---- row-sharded entry points ("multi-GPU variants taking an ncclComm_t + shard offset") -------------------

/* EigenMaps::search over a row-sharded index: local top-k with global indices, all-gather of the per-shard lists
 * (Q x k x 16 B per rank), k-way merge by (score desc, global index asc) = the order of the reference's stable sort
 * (src/core.rs:785).  Every rank returns the merged result. */
int asb_index_search_sharded(asb_ctx *ctx, asb_comm *comm, const asb_index *ix, const double *queries, int64_t nq, int64_t k,
                             double alpha, int64_t *idx, double *score, int64_t *count, double *lambda_q_out) {
    ASB_TRY(set_device(ctx));
    if (!comm || !ix || !queries || !idx || !score) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_search_sharded: null pointer");
    if (nq <= 0 || k < 1) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_search_sharded: bad sizes");
    const int64_t f = ix->f;
    const int R = comm->nranks;
    DevIn<double> q;
    DevTmp<double> lq, ls, gs;
    DevTmp<int> flags;
    DevTmp<int64_t> li, lc, gi, cnt_tmp;
    DevOut<int64_t> oi, oc;
    DevOut<double> os;
    ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
    ASB_TRY(lq.init(ctx, (size_t)nq));
    ASB_TRY(flags.init(ctx, 2));
    ASB_TRY(ls.init(ctx, (size_t)nq * k));
    ASB_TRY(li.init(ctx, (size_t)nq * k));
    ASB_TRY(lc.init(ctx, (size_t)nq));
    ASB_TRY(gs.init(ctx, (size_t)R * nq * k));
    ASB_TRY(gi.init(ctx, (size_t)R * nq * k));
    ASB_TRY(cnt_tmp.init(ctx, (size_t)nq));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, 2 * sizeof(int), ctx->stream));
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    ASB_TRY(asb_dev_taumode(ctx, q.ptr, nq, f, ix->plan, ix->tau_mode, ix->tau_value, lq.ptr, nullptr, nullptr, flags.ptr));
    // every collective is reached by every rank: an error on one rank is agreed on first
    DevTmp<long long> agree;
    ASB_TRY(agree.init(ctx, 1));
    int h[2] = {0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(h, flags.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    long long bad = h[0] ? ASB_ERR_NONFINITE_QUERY : 0;
    int rc = ASB_OK;
    if (!bad) {
        rc = asb_dev_search(ctx, ix->items, ix->lambdas, ix->norms2, ix->n, f, q.ptr, lq.ptr, nq, k, alpha, ix->shard_offset,
                            li.ptr, ls.ptr, lc.ptr, flags.ptr + 1);
        if (rc == ASB_OK) {
            ASB_CUDA(ctx, cudaMemcpyAsync(h + 1, flags.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            rc = search_status_to_rc(ctx, h[1]);
        }
        bad = rc;
    }
    ctx->kernel_ms["sharded_search_local_ms"] = ms_since(t_begin);
    const auto t_x = std::chrono::steady_clock::now();
    ASB_CUDA(ctx, cudaMemcpyAsync(agree.ptr, &bad, 8, cudaMemcpyHostToDevice, ctx->stream));
    ASB_TRY(asb_comm_allreduce_i64(ctx, comm, agree.ptr, 1, ASB_RED_MAX));
    long long worst = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&worst, agree.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (worst != 0) {
        if (bad == 0) ctx->last_error = "another rank failed in index_search_sharded";
        else if (bad == ASB_ERR_NONFINITE_QUERY)
            ctx->last_error = "Query item contains invalid values (NaN or infinity). All values must be finite.";
        return (int)(bad ? bad : worst);
    }
    if (lambda_q_out) ASB_TRY(copy_out(ctx, lambda_q_out, lq.ptr, (size_t)nq * sizeof(double)));
    // unused tail slots of a short shard: idx = -1 (the merge skips them)
    ASB_TRY(asb_dev_mark_tail(ctx, li.ptr, lc.ptr, nq, k));
    ASB_TRY(asb_comm_allgather_bytes(ctx, comm, ls.ptr, gs.ptr, (size_t)nq * k * sizeof(double)));
    ASB_TRY(asb_comm_allgather_bytes(ctx, comm, li.ptr, gi.ptr, (size_t)nq * k * sizeof(int64_t)));
    ASB_TRY(oi.init(ctx, idx, (size_t)nq * k));
    ASB_TRY(os.init(ctx, score, (size_t)nq * k));
    ASB_TRY(oc.init(ctx, count, count ? (size_t)nq : 0));
    ASB_TRY(asb_dev_topk_merge(ctx, gs.ptr, gi.ptr, R, nq, k, os.ptr, oi.ptr, count ? oc.ptr : cnt_tmp.ptr));
    ASB_TRY(oi.finish(ctx));
    ASB_TRY(os.finish(ctx));
    ASB_TRY(oc.finish(ctx));
    ASB_TRY(asb_sync(ctx));
    ctx->kernel_ms["sharded_search_exchange_ms"] = ms_since(t_x);   // includes waiting for the slowest rank
    return ASB_OK;
}

/* The Two-NN scan over a row-sharded dataset (src/clustering.rs:118-145): sample_idx are GLOBAL row
 * indices (the same list on every rank).  The sample rows are assembled on every rank (all-reduce of a gather buffer:
 * each entry is non-zero on exactly one rank), every rank scans its own shard, the per-shard two nearest distances
 * are all-gathered and merged.  d1 / d2 (f64[s]) are returned on every rank. */
int asb_twonn_distances_sharded(asb_ctx *ctx, asb_comm *comm, const double *rows_local, int64_t n_local, int64_t f,
                                int64_t shard_offset, const int64_t *sample_idx, int64_t s, double *d1, double *d2) {
    ASB_TRY(set_device(ctx));
    if (!comm || !rows_local || !sample_idx || !d1 || !d2) ASB_FAIL(ctx, ASB_ERR_INVALID, "twonn_sharded: null pointer");
    if (n_local < 1 || f <= 0 || s <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "twonn_sharded: bad sizes");
    const int R = comm->nranks;
    DevIn<double> rows;
    DevIn<int64_t> samp;
    DevTmp<double> q, part, all;
    DevTmp<int64_t> self;
    DevOut<double> o1, o2;
    ASB_TRY(rows.init(ctx, rows_local, (size_t)n_local * f));
    ASB_TRY(samp.init(ctx, sample_idx, (size_t)s));
    ASB_TRY(q.init(ctx, (size_t)s * f));
    ASB_TRY(self.init(ctx, (size_t)s));
    ASB_TRY(part.init(ctx, (size_t)2 * s));
    ASB_TRY(all.init(ctx, (size_t)R * 2 * s));
    ASB_TRY(o1.init(ctx, d1, (size_t)s));
    ASB_TRY(o2.init(ctx, d2, (size_t)s));
    StageTimer t(ctx, "twonn");
    ASB_TRY(asb_dev_twonn_gather(ctx, rows.ptr, n_local, f, shard_offset, samp.ptr, s, q.ptr, self.ptr));
    ASB_TRY(asb_comm_allreduce_f64(ctx, comm, q.ptr, (size_t)s * f, ASB_RED_SUM));
    ASB_TRY(asb_dev_twonn_queries(ctx, q.ptr, self.ptr, s, rows.ptr, n_local, f, part.ptr, part.ptr + s));
    ASB_TRY(asb_comm_allgather_bytes(ctx, comm, part.ptr, all.ptr, (size_t)2 * s * sizeof(double)));
    ASB_TRY(asb_dev_twonn_merge(ctx, all.ptr, R, s, o1.ptr, o2.ptr));
    t.stop();
    ASB_TRY(o1.finish(ctx));
    ASB_TRY(o2.finish(ctx));
    return asb_sync(ctx);
}

int64_t asb_index_shard_offset(const asb_index *ix) { return ix ? ix->shard_offset : 0; }

}  // extern "C"

// queries -> (projected queries when the index carries a projection) and their lambdas (src/core.rs:533-549)
static int index_query_prep(asb_ctx *ctx, const asb_index *ix, const double *q_d, int64_t nq, DevTmp<double> &qp,
                            const double **q_used, double *lq_d, int *flag_d) {
    const int64_t f = ix->f;
    // the finiteness assert looks at the RAW query (:534-537): a throw-away lambda pass over it sets the flag
    *q_used = q_d;
    int64_t fq = f;
    if (ix->r) {
        DevTmp<double> scratch;
        ASB_TRY(scratch.init(ctx, (size_t)nq));
        ASB_TRY(asb_dev_nonfinite_rows(ctx, q_d, nq, f, flag_d));
        ASB_TRY(qp.init(ctx, (size_t)nq * ix->r));
        ASB_TRY(asb_dev_project(ctx, q_d, nq, f, ix->proj, ix->r, qp.ptr));
        *q_used = qp.ptr;
        fq = ix->r;
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return asb_dev_taumode(ctx, *q_used, nq, fq, ix->plan, ix->tau_mode, ix->tau_value, lq_d, nullptr, nullptr, flag_d);
}

extern "C" {

int asb_index_prepare_query(asb_ctx *ctx, const asb_index *ix, const double *queries, int64_t nq, double *lambda_q) {
    ASB_TRY(set_device(ctx));
    if (!ix || !queries || !lambda_q || nq <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_prepare_query: bad argument");
    DevIn<double> q;
    DevTmp<double> qp;
    DevOut<double> lq;
    DevTmp<int> flag;
    ASB_TRY(q.init(ctx, queries, (size_t)nq * ix->f));
    ASB_TRY(lq.init(ctx, lambda_q, (size_t)nq));
    ASB_TRY(flag.init(ctx, 1));
    ASB_CUDA(ctx, cudaMemsetAsync(flag.ptr, 0, sizeof(int), ctx->stream));
    const double *qu = nullptr;
    ASB_TRY(index_query_prep(ctx, ix, q.ptr, nq, qp, &qu, lq.ptr, flag.ptr));
    int h = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&h, flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h)
        ASB_FAIL(ctx, ASB_ERR_NONFINITE_QUERY,
                 "Query item contains invalid values (NaN or infinity). All values must be finite.");
    ASB_TRY(lq.finish(ctx));
    return asb_sync(ctx);
}

int asb_index_search_energy(asb_ctx *ctx, asb_index *ix, const double *queries, int64_t nq, int64_t k, double w_lambda,
                            double w_dirichlet, int64_t *idx, double *score, int64_t *count) {
    ASB_TRY(set_device(ctx));
    if (!ix || !queries || !idx || !score || nq <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "index_search_energy: bad argument");
    if (k < 1 || k > 56) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "index_search_energy: k=%lld outside 1..56", (long long)k);
    const int64_t f = ix->f, n = ix->n, d = ix->r ? ix->r : f;   // d: dimension the differences live in
    DevIn<double> q;
    DevTmp<double> qp, lq, qs;
    DevTmp<int> flags;
    DevTmp<int64_t> cnt_tmp;
    DevOut<int64_t> oi, oc;
    DevOut<double> os;
    ASB_TRY(q.init(ctx, queries, (size_t)nq * f));
    ASB_TRY(lq.init(ctx, (size_t)nq));
    ASB_TRY(flags.init(ctx, 2));
    ASB_TRY(cnt_tmp.init(ctx, (size_t)nq));
    ASB_CUDA(ctx, cudaMemsetAsync(flags.ptr, 0, 2 * sizeof(int), ctx->stream));
    const double *qu = nullptr;
    ASB_TRY(index_query_prep(ctx, ix, q.ptr, nq, qp, &qu, lq.ptr, flags.ptr));   // lambda_q (:885) and project_vec(query)
    int h[2] = {0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(h, flags.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[0])
        ASB_FAIL(ctx, ASB_ERR_NONFINITE_QUERY,
                 "Query item contains invalid values (NaN or infinity). All values must be finite.");
    // project_vec(item) for every item, once per index (the reference redoes it per query and item, :889-891)
    const double *xu = ix->items;
    if (ix->r) {
        if (!ix->items_proj) {
            ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->items_proj, (size_t)n * d * sizeof(double), ctx->stream));
            ASB_TRY(asb_dev_project(ctx, ix->items, n, f, ix->proj, d, ix->items_proj));
        }
        xu = ix->items_proj;
    }
    // projected_dirichlet (:866-882): through the signals when they exist and their column count is the difference's
    // length (they are built on the feature graph's nodes, so this holds whenever they exist)
    const bool use_sig = ix->sig_indptr != nullptr;
    const double *xr = xu, *qr = qu;   // what the ranking pass measures distances between
    if (use_sig) {
        if (!ix->items_sig) {
            ASB_CUDA(ctx, cudaMallocAsync((void **)&ix->items_sig, (size_t)n * d * sizeof(double), ctx->stream));
            ASB_TRY(asb_dev_csr_apply_rows(ctx, ix->sig_indptr, ix->sig_indices, ix->sig_data, d, xu, n, ix->items_sig));
        }
        ASB_TRY(qs.init(ctx, (size_t)nq * d));
        ASB_TRY(asb_dev_csr_apply_rows(ctx, ix->sig_indptr, ix->sig_indices, ix->sig_data, d, qu, nq, qs.ptr));
        xr = ix->items_sig;   // |S q' - S x'| = |S (q' - x')|: the ranking pass works on the transformed rows
        qr = qs.ptr;
    }
    ASB_TRY(oi.init(ctx, idx, (size_t)nq * k));
    ASB_TRY(os.init(ctx, score, (size_t)nq * k));
    ASB_TRY(oc.init(ctx, count, count ? (size_t)nq : 0));
    StageTimer ts(ctx, "search_energy");
    int rc = asb_dev_search_energy_ex(ctx, xr, xu, ix->lambdas, n, d, qr, qu, lq.ptr, nq, k, w_lambda, w_dirichlet,
                                      use_sig ? ix->sig_indptr : nullptr, ix->sig_indices, ix->sig_data, oi.ptr, os.ptr,
                                      count ? oc.ptr : cnt_tmp.ptr, flags.ptr + 1);
    ts.stop();
    ASB_TRY(rc);
    ASB_CUDA(ctx, cudaMemcpyAsync(h + 1, flags.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASB_TRY(search_status_to_rc(ctx, h[1]));
    ASB_TRY(oi.finish(ctx));
    ASB_TRY(os.finish(ctx));
    ASB_TRY(oc.finish(ctx));
    return asb_sync(ctx);
}

}  // extern "C"
