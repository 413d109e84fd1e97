// cpp_example.cpp -- examples/01_compare_cosine.rs through the C++ mirror: 64 x 24 table given on
// stdin-free form (generated blobs here), query = row 3 x 1.02, alpha = 1 -> the item itself first.
//   g++ -std=c++17 -Iinclude tools/cpp_example.cpp -Larrowspace-rs_b200 -larrowspace_b200 -o cpp_example
#include <cstdio>
#include <random>
#include "arrowspace_b200.hpp"

int main() {
    using namespace arrowspace;
    try {
        Context ctx(0);
        const int64_t n = 4096, f = 64;
        std::mt19937_64 rng(42);
        std::uniform_real_distribution<double> u(0.0, 1.0);
        std::normal_distribution<double> g(0.0, 0.05);
        std::vector<double> centres(16 * f), rows(n * f);
        for (auto &c : centres) c = u(rng);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < f; ++j) rows[i * f + j] = std::max(0.0, centres[(i % 16) * f + j] + g(rng));
        auto built = ArrowSpaceBuilder::new_(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25)
                         .with_synthesis(TauMode::Median()).with_seed(42).with_inline_sampling_none()
                         .with_cluster_params(32, 1.5 * f * 0.0025 * 2).build(rows.data(), n, f);
        ArrowSpace &aspace = built.first;
        GraphLaplacian &gl = built.second;
        std::vector<double> q(rows.begin() + 3 * f, rows.begin() + 4 * f);
        for (auto &v : q) v *= 1.02;
        auto res = aspace.search(q, gl, 3, 1.0);
        std::printf("clusters=%lld nnz=%lld top: %zu (%.6f) %zu %zu\n", (long long)aspace.n_clusters,
                    (long long)gl.nnz(), res[0].first, res[0].second, res[1].first, res[2].first);
        return res[0].first == 3 ? 0 : 1;
    } catch (const Panic &p) {
        std::fprintf(stderr, "panic (status %d): %s\n", p.status, p.what());
        return 2;
    }
}
