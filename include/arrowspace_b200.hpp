// arrowspace_b200.hpp -- header-only C++17 host mirror of the reference's API for the lambda-tau
// build + lambda-aware search path, on top of the C ABI (arrowspace_b200.h).
//
// Names, argument meaning and error behaviour follow the Rust originals (cited per member; paths
// relative to the arrowspace-rs repository).  Where the reference panics this mirror throws
// arrowspace::Panic carrying the ABI status and the reference's message.  No arithmetic happens
// on the host: every method marshals buffers into libarrowspace_b200.so.
#pragma once

#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "arrowspace_b200.h"

namespace arrowspace {

struct Panic : std::runtime_error {
    int status;
    Panic(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

class Context {
  public:
    explicit Context(int device = 0, void *stream = nullptr) {
        int rc = asb_ctx_create(device, stream, &ctx_);
        if (rc != ASB_OK) throw Panic(rc, "asb_ctx_create failed: no sm_100 device (there is no CPU fallback)");
    }
    ~Context() { asb_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    asb_ctx *get() const { return ctx_; }
    void check(int rc) const {
        if (rc != ASB_OK) throw Panic(rc, asb_last_error(ctx_));
    }
    /// switches listed in arrowspace_b200.h ("search_prefilter", "cluster_replay", ...); results never depend on them
    void set_option(const char *key, double value) { check(asb_ctx_set_option(ctx_, key, value)); }
    /// device time / diagnostics of the most recent call ("search_pf_kernel", "search_pf_used", "cluster_kernel", ...)
    double kernel_ms(const char *which) const { return asb_last_kernel_ms(ctx_, which); }

  private:
    asb_ctx *ctx_ = nullptr;
};

/// TauMode (src/taumode.rs:75-82)
struct TauMode {
    int mode = ASB_TAU_MEDIAN;
    double value = 0.0;
    static TauMode Fixed(double t) { return {ASB_TAU_FIXED, t}; }
    static TauMode Median() { return {ASB_TAU_MEDIAN, 0.0}; }
    static TauMode Mean() { return {ASB_TAU_MEAN, 0.0}; }
    static TauMode Percentile(double p) { return {ASB_TAU_PERCENTILE, p}; }
};

/// GraphLaplacian (src/graph.rs:127-135): `matrix` is the F x F CSR, `nnodes` = N.
struct GraphLaplacian {
    std::vector<int64_t> indptr, indices;
    std::vector<double> data;
    int64_t nnodes = 0;
    asb_graph_params graph_params{};
    std::vector<double> init_data;  // X x F centroids, row-major
    int64_t nnz() const { return indptr.empty() ? 0 : indptr.back(); }
    std::pair<int64_t, int64_t> shape() const {
        int64_t f = (int64_t)indptr.size() - 1;
        return {f, f};
    }
};

/// ArrowItem (src/core.rs:84-87)
struct ArrowItem {
    std::vector<double> item;
    double lambda = 0.0;
};

/// ArrowSpace (src/core.rs:366-385) backed by a device-resident index.
class ArrowSpace {
  public:
    int64_t nitems = 0, nfeatures = 0, n_clusters = 0;
    std::vector<double> lambdas;
    std::vector<int64_t> signals_indptr, signals_indices;  // aspace.signals (with_spectral), CSR F x F
    std::vector<double> signals_data;
    std::vector<int64_t> cluster_assignments;  // -1 = None
    std::vector<uint64_t> cluster_sizes;
    double cluster_radius = 0.0;
    TauMode taumode;

    ~ArrowSpace() { asb_index_destroy(index_); }
    ArrowSpace() = default;
    ArrowSpace(ArrowSpace &&o) noexcept { *this = std::move(o); }
    ArrowSpace &operator=(ArrowSpace &&o) noexcept {
        std::swap(index_, o.index_);
        std::swap(ctx_, o.ctx_);
        nitems = o.nitems; nfeatures = o.nfeatures; n_clusters = o.n_clusters;
        lambdas = std::move(o.lambdas); cluster_assignments = std::move(o.cluster_assignments);
        cluster_sizes = std::move(o.cluster_sizes); cluster_radius = o.cluster_radius; taumode = o.taumode;
        return *this;
    }

    /// ArrowSpace::prepare_query_item (src/core.rs:533-549)
    double prepare_query_item(const std::vector<double> &item, const GraphLaplacian &gl) const {
        if ((int64_t)item.size() != nfeatures)
            throw Panic(ASB_ERR_DIM, "Query dimension doesn't match index original dimension");
        double lq = 0.0;
        ctx_->check(asb_prepare_query_lambdas(ctx_->get(), item.data(), 1, nfeatures, gl.indptr.data(),
                                              gl.indices.data(), gl.data.data(), taumode.mode, taumode.value, &lq));
        return lq;
    }
    /// ArrowSpace::search_lambda_aware (src/core.rs:760-798)
    std::vector<std::pair<size_t, double>> search_lambda_aware(const ArrowItem &q, size_t k, double alpha) const {
        std::vector<int64_t> idx(k ? k : 1);
        std::vector<double> score(k ? k : 1);
        int64_t count = 0;
        ctx_->check(asb_index_search_lambda_aware(ctx_->get(), index_, q.item.data(), &q.lambda, 1, (int64_t)k, alpha,
                                                  idx.data(), score.data(), &count));
        std::vector<std::pair<size_t, double>> out;
        for (int64_t r = 0; r < count; ++r) out.emplace_back((size_t)idx[r], score[r]);
        return out;
    }
    /// EigenMaps::search (src/eigenmaps.rs:410-455)
    std::vector<std::pair<size_t, double>> search(const std::vector<double> &item, const GraphLaplacian &gl, size_t k,
                                                  double alpha) const {
        return search_lambda_aware(ArrowItem{item, prepare_query_item(item, gl)}, k, alpha);
    }
    /// Batched EigenMaps::search: queries nq x F row-major; returns (idx, score, count).
    void search_batch(const double *queries, int64_t nq, int64_t k, double alpha, std::vector<int64_t> &idx,
                      std::vector<double> &score, std::vector<int64_t> &count) const {
        idx.assign((size_t)(nq * k), -1);
        score.assign((size_t)(nq * k), 0.0);
        count.assign((size_t)nq, 0);
        ctx_->check(asb_index_search(ctx_->get(), index_, queries, nq, k, alpha, idx.data(), score.data(),
                                     count.data(), nullptr));
    }
    asb_index_info info() const {
        asb_index_info i{};
        asb_index_info_get(index_, &i);
        return i;
    }

  private:
    friend class ArrowSpaceBuilder;
    asb_index *index_ = nullptr;
    const Context *ctx_ = nullptr;
};

/// ArrowSpaceBuilder (src/builder.rs:20-57; defaults :59-91)
class ArrowSpaceBuilder {
  public:
    explicit ArrowSpaceBuilder(const Context &ctx) : ctx_(&ctx) {}
    static ArrowSpaceBuilder new_(const Context &ctx) { return ArrowSpaceBuilder(ctx); }

    ArrowSpaceBuilder &with_lambda_graph(double eps, size_t k, size_t topk, double p,
                                         std::optional<double> sigma) {  // :109-137
        lambda_eps = eps; lambda_k = (int64_t)k; lambda_topk = (int64_t)topk; lambda_p = p; lambda_sigma = sigma;
        return *this;
    }
    ArrowSpaceBuilder &with_synthesis(TauMode t) { synthesis = t; return *this; }           // :142-146
    ArrowSpaceBuilder &with_normalisation(bool v) { normalise = v; return *this; }          // :148-152
    ArrowSpaceBuilder &with_spectral(bool v) { prebuilt_spectral = v; return *this; }       // :157-162
    ArrowSpaceBuilder &with_sparsity_check(bool v) { sparsity_check = v; return *this; }    // :164-168
    ArrowSpaceBuilder &with_inline_sampling_none() { sampling = false; return *this; }      // :170-179 (None only)
    ArrowSpaceBuilder &with_dims_reduction(bool enable) { use_dims_reduction = enable; return *this; }  // :181-185
    ArrowSpaceBuilder &with_seed(uint64_t seed) { clustering_seed = seed; deterministic_clustering = true; return *this; }  // :190-195
    /// (k_opt, radius) from the host heuristic compute_optimal_k (src/clustering.rs:36-72)
    ArrowSpaceBuilder &with_cluster_params(size_t max_clusters, double radius) {
        cluster_max_clusters = (int64_t)max_clusters; cluster_radius = radius;
        return *this;
    }

    /// ArrowSpaceBuilder::build (src/builder.rs:249-455); rows: n x f row-major (host or device).
    std::pair<ArrowSpace, GraphLaplacian> build(const double *rows, int64_t n, int64_t f) {
        if (n == 0) throw Panic(ASB_ERR_EMPTY, "items cannot be empty");
        if (sampling) throw Panic(ASB_ERR_UNSUPPORTED, "inline sampling is OS-seeded in the reference; use None");
        if (use_dims_reduction) throw Panic(ASB_ERR_UNSUPPORTED, "JL projection is a 'next' row");
        if (cluster_max_clusters <= 0) throw Panic(ASB_ERR_INVALID, "call with_cluster_params (host heuristic output)");
        asb_build_params bp{};
        bp.graph = {lambda_eps, lambda_k, lambda_topk, lambda_p, lambda_sigma ? 1 : 0, lambda_sigma.value_or(0.0),
                    normalise ? 1 : 0, sparsity_check ? 1 : 0, 0, 0};
        bp.tau_mode = synthesis.mode; bp.tau_value = synthesis.value;
        bp.max_clusters = cluster_max_clusters; bp.radius = cluster_radius;
        bp.apply_define_result_k = 1;  // define_result_k, :225-233
        bp.spectral = prebuilt_spectral ? 1 : 0;
        ArrowSpace a;
        a.ctx_ = ctx_;
        ctx_->check(asb_index_build(ctx_->get(), rows, n, f, &bp, &a.index_));
        asb_index_info info = a.info();
        a.nitems = n; a.nfeatures = f; a.n_clusters = info.n_clusters; a.cluster_radius = cluster_radius;
        a.taumode = synthesis;
        a.lambdas.resize((size_t)n); a.cluster_assignments.resize((size_t)n); a.cluster_sizes.resize((size_t)info.n_clusters);
        GraphLaplacian gl;
        gl.indptr.resize((size_t)f + 1); gl.indices.resize((size_t)info.nnz); gl.data.resize((size_t)info.nnz);
        gl.init_data.resize((size_t)(info.n_clusters * f)); gl.nnodes = n; gl.graph_params = bp.graph;
        ctx_->check(asb_index_lambdas(ctx_->get(), a.index_, a.lambdas.data()));
        ctx_->check(asb_index_assignments(ctx_->get(), a.index_, a.cluster_assignments.data()));
        ctx_->check(asb_index_cluster_sizes(ctx_->get(), a.index_, a.cluster_sizes.data()));
        ctx_->check(asb_index_centroids(ctx_->get(), a.index_, gl.init_data.data()));
        ctx_->check(asb_index_laplacian(ctx_->get(), a.index_, gl.indptr.data(), gl.indices.data(), gl.data.data()));
        if (prebuilt_spectral) {  // aspace.signals, src/core.rs:370
            a.signals_indptr.resize((size_t)f + 1); a.signals_indices.resize((size_t)info.nnz_signals);
            a.signals_data.resize((size_t)info.nnz_signals);
            ctx_->check(asb_index_signals(ctx_->get(), a.index_, a.signals_indptr.data(), a.signals_indices.data(),
                                          a.signals_data.data()));
        }
        return {std::move(a), std::move(gl)};
    }

    double lambda_eps = 1e-3; int64_t lambda_k = 6, lambda_topk = 3; double lambda_p = 2.0;
    std::optional<double> lambda_sigma; bool normalise = false, sparsity_check = false, sampling = true;
    TauMode synthesis; int64_t cluster_max_clusters = 0; double cluster_radius = 1.0;
    std::optional<uint64_t> clustering_seed; bool deterministic_clustering = false, use_dims_reduction = false;
    bool prebuilt_spectral = false;

  private:
    const Context *ctx_;
};

}  // namespace arrowspace
