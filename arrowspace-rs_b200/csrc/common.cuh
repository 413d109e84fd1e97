// common.cuh -- context, error handling and host<->device staging shared by all stages.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/arrowspace_b200.h"

struct KTimerEvents {
    cudaEvent_t a = nullptr, b = nullptr;
};

// Rows still travelling host -> device on a second stream while the build already works on the head of the matrix
// (asb_index_build from host memory): events recorded after every copied chunk, in row order.
struct RowsInFlight {
    const double *base = nullptr;   // device address of row 0
    int64_t f = 0, rows_per_event = 0;
    std::vector<cudaEvent_t> events;
    int waited = -1;                // the compute stream already waits for events[0 .. waited]
};

struct asb_ctx {
    RowsInFlight *rows_in_flight = nullptr;
    std::map<std::string, KTimerEvents> ktimers;  // per-kernel device timers (see KernelTimer)
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string last_error;
    int64_t launches = 0;
    int sm_count = 148;
    std::map<std::string, double> kernel_ms;
    std::map<std::string, double> options;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

#define ASB_FAIL(ctx, code, ...)                          \
    do {                                                  \
        char _b[512];                                     \
        snprintf(_b, sizeof(_b), __VA_ARGS__);            \
        (ctx)->last_error = _b;                           \
        return (code);                                    \
    } while (0)

#define ASB_CUDA(ctx, expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            char _b[512];                                                                \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                        \
            (ctx)->last_error = _b;                                                      \
            return ASB_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define ASB_TRY(expr)                \
    do {                             \
        int _rc = (expr);            \
        if (_rc != ASB_OK) return _rc; \
    } while (0)

// Is `p` a device pointer usable by kernels on the current device?
static inline bool asb_is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Input staging: device pointers are borrowed, host pointers are copied to a
// stream-ordered temporary.
template <typename T>
struct DevIn {
    const T *ptr = nullptr;
    T *owned = nullptr;
    cudaStream_t stream = nullptr;
    int init(asb_ctx *ctx, const T *src, size_t count) {
        stream = ctx->stream;
        if (count == 0 || src == nullptr) {
            ptr = nullptr;
            return ASB_OK;
        }
        if (asb_is_device_ptr(src)) {
            ptr = src;
            return ASB_OK;
        }
        ASB_CUDA(ctx, cudaMallocAsync((void **)&owned, count * sizeof(T), stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(owned, src, count * sizeof(T), cudaMemcpyHostToDevice, stream));
        ptr = owned;
        return ASB_OK;
    }
    ~DevIn() {
        if (owned) cudaFreeAsync(owned, stream);
    }
};

// Output staging: device pointers are written in place, host pointers get a device
// temporary that finish() copies back (and synchronises).
template <typename T>
struct DevOut {
    T *ptr = nullptr;
    T *owned = nullptr;
    T *host = nullptr;
    size_t count = 0;
    cudaStream_t stream = nullptr;
    int init(asb_ctx *ctx, T *dst, size_t n) {
        stream = ctx->stream;
        count = n;
        if (n == 0 || dst == nullptr) {
            ptr = nullptr;
            return ASB_OK;
        }
        if (asb_is_device_ptr(dst)) {
            ptr = dst;
            return ASB_OK;
        }
        host = dst;
        ASB_CUDA(ctx, cudaMallocAsync((void **)&owned, n * sizeof(T), stream));
        ptr = owned;
        return ASB_OK;
    }
    int finish(asb_ctx *ctx, size_t n_valid = (size_t)-1) {
        if (owned && host) {
            size_t n = n_valid == (size_t)-1 ? count : n_valid;
            if (n > 0)
                ASB_CUDA(ctx, cudaMemcpyAsync(host, owned, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        }
        return ASB_OK;
    }
    ~DevOut() {
        if (owned) cudaFreeAsync(owned, stream);
    }
};

// Stream-ordered scratch buffer.
template <typename T>
struct DevTmp {
    T *ptr = nullptr;
    cudaStream_t stream = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp &) = delete;   // owns its buffer
    DevTmp &operator=(const DevTmp &) = delete;
    int init(asb_ctx *ctx, size_t count) {
        if (ptr) cudaFreeAsync(ptr, stream);  // re-init: release the previous buffer (stream-ordered)
        ptr = nullptr;
        stream = ctx->stream;
        if (count == 0) count = 1;
        ASB_CUDA(ctx, cudaMallocAsync((void **)&ptr, count * sizeof(T), stream));
        return ASB_OK;
    }
    ~DevTmp() {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
};

// Make the context's stream wait until the rows before `end` (a device address inside the matrix being uploaded) have
// landed.  No-op unless an upload is in flight.
static inline void asb_wait_rows(asb_ctx *ctx, const double *end) {
    RowsInFlight *r = ctx->rows_in_flight;
    if (!r || r->events.empty() || end <= r->base) return;
    const int64_t rows = (int64_t)((end - r->base) + r->f - 1) / r->f;
    int idx = (int)((rows + r->rows_per_event - 1) / r->rows_per_event) - 1;
    if (idx >= (int)r->events.size()) idx = (int)r->events.size() - 1;
    if (idx > r->waited) {
        cudaStreamWaitEvent(ctx->stream, r->events[idx], 0);
        r->waited = idx;
    }
}

static inline int asb_sync(asb_ctx *ctx) {
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}

static inline int asb_check_launch(asb_ctx *ctx, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char b[512];
        snprintf(b, sizeof(b), "launch of %s failed: %s (%s)", what, cudaGetErrorName(e),
                 cudaGetErrorString(e));
        ctx->last_error = b;
        return ASB_ERR_CUDA;
    }
    ctx->launches++;
    return ASB_OK;
}

// RAII device timer on the context's stream; stores ms in ctx->kernel_ms[name] (accumulating
// within one top-level call is the caller's business).
struct StageTimer {
    asb_ctx *ctx;
    const char *name;
    StageTimer(asb_ctx *c, const char *n) : ctx(c), name(n) { cudaEventRecord(ctx->ev0, ctx->stream); }
    double stop() {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->kernel_ms[name] = ms;
        return ms;
    }
};

// Brackets ONE kernel launch with CUDA events on the launching stream; the elapsed time is
// resolved lazily by asb_last_kernel_ms(name) (the roofline numbers of bench.py come from here).
struct KernelTimer {
    asb_ctx *ctx;
    KTimerEvents *ev;
    KernelTimer(asb_ctx *c, const char *name) : ctx(c) {
        ev = &ctx->ktimers[name];
        if (!ev->a) {
            cudaEventCreate(&ev->a);
            cudaEventCreate(&ev->b);
        }
        cudaEventRecord(ev->a, ctx->stream);
    }
    ~KernelTimer() { cudaEventRecord(ev->b, ctx->stream); }
};

// ---- internal stage entry points on DEVICE pointers (used by the C ABI and asb_index) ----

struct DeviceCsr {  // device-resident CSR of the feature graph (int64 as at the ABI)
    const int64_t *indptr = nullptr;
    const int64_t *indices = nullptr;
    const double *data = nullptr;
    int64_t f = 0, nnz = 0;
};

// Entry record of the on-device graph plan consumed by the taumode kernel.
struct __align__(16) GraphEntry {
    double v;   // L_ij
    int32_t j;  // column
    int32_t row;
};

struct GraphPlan {  // built from a CSR (host side, tiny), lives in device memory
    GraphEntry *entries = nullptr;  // all stored entries, row-major order
    int32_t *row_ptr = nullptr;     // f+1
    int64_t f = 0, nnz = 0;
    // symmetric form (taumode_sym.cuh): undirected edges i<j with w = -L_ij, per-row residuals
    void *sym_edges = nullptr;      // SymEdge[nedges]
    double *resid = nullptr;        // f
    int64_t nedges = 0;
    bool is_sym = false, all_pos = false;
    cudaStream_t stream = nullptr;
    void release() {
        if (entries) cudaFreeAsync(entries, stream);
        if (row_ptr) cudaFreeAsync(row_ptr, stream);
        if (sym_edges) cudaFreeAsync(sym_edges, stream);
        if (resid) cudaFreeAsync(resid, stream);
        entries = nullptr;
        row_ptr = nullptr;
        sym_edges = nullptr;
        resid = nullptr;
    }
};

int asb_graph_plan_from_host(asb_ctx *ctx, const int64_t *indptr, const int64_t *indices,
                             const double *data, int64_t f, GraphPlan *plan);
int asb_dev_taumode(asb_ctx *ctx, const double *items_d, int64_t n, int64_t f, const GraphPlan &plan,
                    int tau_mode, double tau_value, double *lambdas_d, double *norms2_d,
                    double *stats_d /*3 doubles: min,max,sum or null*/, int *nonfinite_flag_d);
int asb_dev_project(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, const double *proj_d, int64_t r,
                    double *out_d);
int asb_dev_search_energy(asb_ctx *ctx, const double *items_d, const double *lambdas_d, const double *norms2_d,
                          int64_t n, int64_t f, const double *queries_d, const double *lambda_q_d, int64_t nq,
                          int64_t k, double w_lambda, double w_dirichlet, int64_t index_offset, int64_t *idx_d,
                          double *score_d, int64_t *count_d, int *status_d);
int asb_dev_search(asb_ctx *ctx, const double *items_d, const double *lambdas_d, const double *norms2_d,
                   int64_t n, int64_t f, const double *queries_d, const double *lambda_q_d, int64_t nq,
                   int64_t k, double alpha, int64_t index_offset, int64_t *idx_d, double *score_d,
                   int64_t *count_d, int *status_d);
int asb_dev_twonn(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, const int64_t *sample_d,
                  int64_t s, double *d1_d, double *d2_d);
int asb_dev_twonn_queries(asb_ctx *ctx, const double *q_rows_d, const int64_t *self_d, int64_t s, const double *rows_d,
                          int64_t n, int64_t f, double *d1_d, double *d2_d);
int asb_dev_norms2(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, double *norms2_d);
int asb_dev_cluster(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int64_t max_clusters,
                    double radius, double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d,
                    int64_t *x_out_host, int64_t init_k = 0);
// the walk itself (cluster.cu); asb_dev_cluster (cluster_replay.cu) = this, or -- option "cluster_replay" -- a
// sequential prefix followed by certified parallel replay of row chunks with this kernel as the per-chunk fallback
int asb_dev_cluster_seq(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int64_t max_clusters,
                        double radius, double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d,
                        int64_t *x_out_host, int64_t init_k = 0);
// nearest and second nearest of `k_items` rows (centroids) for each of m query rows: ids and Euclidean distances
// (GEMM form on the FP64 tensor pipe, the Two-NN kernel); idx_d int64[m*2], dist_d f64[m*2] ascending
int asb_dev_top2_l2(asb_ctx *ctx, const double *q_d, int64_t m, int64_t f, const double *items_d, int64_t k_items,
                    const double *qn2_d, const double *xn2_d, const int64_t *minus1_d, int64_t *idx_d, double *dist_d,
                    int64_t *cnt_d, int *status_d);
// the same question answered by the tcgen05 tile with certified bounds instead of exact distances (search_umma.cuh,
// PF_NEAR): near_idx_d int64[m], near_b_d f64[m x 3] = {dlo, dhi, slo}; *done = false -> use asb_dev_top2_l2
int asb_dev_near_tf32(asb_ctx *ctx, const double *q_d, int64_t m, int64_t f, const double *items_d, int64_t k_items,
                      const double *qn2_d, const double *xn2_d, int64_t *near_idx_d, double *near_b_d, bool *done);
int asb_dev_laplacian(asb_ctx *ctx, const double *centroids_d, int64_t x, int64_t f,
                      const asb_graph_params &gp, int64_t *indptr_d, int64_t *indices_d, double *data_d,
                      int64_t capacity, int64_t *nnz_host);
int asb_dev_range_search(asb_ctx *ctx, const double *lambdas_d, int64_t n, double lambda_q, double eps,
                         int64_t index_offset, int64_t *idx_d, double *dist_d, int64_t capacity,
                         int64_t *count_host);
int asb_dev_topk_merge(asb_ctx *ctx, const double *in_score_d, const int64_t *in_idx_d, int64_t parts,
                       int64_t nq, int64_t k, double *out_score_d, int64_t *out_idx_d, int64_t *out_count_d);
int asb_dev_mark_tail(asb_ctx *ctx, int64_t *idx_d, const int64_t *cnt_d, int64_t nq, int64_t k);
int asb_dev_twonn_gather(asb_ctx *ctx, const double *rows_d, int64_t n_local, int64_t f, int64_t offset,
                         const int64_t *sample_d, int64_t s, double *q_d, int64_t *self_d);
int asb_dev_twonn_merge(asb_ctx *ctx, const double *all_d, int parts, int64_t s, double *d1_d, double *d2_d);
int asb_dev_nonfinite_rows(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int *flag_d);
int asb_dev_csr_apply_rows(asb_ctx *ctx, const int64_t *indptr_d, const int64_t *indices_d, const double *data_d, int64_t d,
                           const double *rows_d, int64_t n, double *out_d);
int asb_dev_search_energy_ex(asb_ctx *ctx, const double *rank_items_d, const double *items_d, const double *lambdas_d,
                             int64_t n, int64_t d, const double *rank_queries_d, const double *queries_d,
                             const double *lambda_q_d, int64_t nq, int64_t k, double w_lambda, double w_dirichlet,
                             const int64_t *sig_indptr_d, const int64_t *sig_indices_d, const double *sig_data_d,
                             int64_t *idx_d, double *score_d, int64_t *count_d, int *status_d);
