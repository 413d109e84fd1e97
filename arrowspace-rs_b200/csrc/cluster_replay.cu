// cluster_replay.cu -- K2, option "cluster_replay" (default 1; validated on B200 in round 2: bit-identical to the oracle's
// walk on every test shape and at 1M x 384, profiles/r02_replay_first_run.json; CPU prototype tests/replay_proto.py).
//
// The walk of run_incremental_clustering_with_sampling (src/clustering.rs:547-928) is sequential because row r sees
// the centroids all earlier rows left behind.  Once the centroids have settled, almost every row's decision can be
// PROVEN without walking:
//   1. guess: b(r) = nearest centroid of the chunk-start snapshot S0 and the distance to the runner-up, for all rows
//      of a chunk at once (dense contraction with a fused 2-min: the Two-NN kernel, asb_dev_top2_l2);
//   2. chains: given the guess, every centroid only sees its own rows, in row order -- K independent sequential
//      chains (one warp each).  A chain evaluates the row's distance to the centroid's CURRENT state, classifies it
//      (d^2 <= radius: running-mean update c += (x - c) / k, :747-751, the reference's element-wise IEEE operations;
//      <= 1.5 radius: count only, :767-781; else dropped) and tracks the centroid's net displacement |c - S0|;
//   3. certification: row r is proven when  d(x_r, c_b at r)  <  d(x_r, runner-up in S0) - max displacement of any
//      centroid in the chunk - rounding slack: no other centroid can be nearer at row r's time, so the walk picks b.
//      (A row farther than sqrt(1.5 radius) from every centroid is dropped whoever is nearest: proven as well.)
//      If every row of the chunk is proven, then by induction over the rows (row r depends only on rows < r) the
//      guess IS the walk's assignment and the chain results are the walk's centroids, bit for bit.
// A chunk with a single unproven row -- or a row that would open a new centroid, or a d^2 within 1e-9 radius of a
// threshold (the chains sum in a different order than the reference) -- is thrown away and walked by the sequential
// kernel from the same start state (asb_dev_cluster_seq with init_k: the resume entry the multi-GPU hand-off uses).
// Every further failure in a row doubles the stretch walked sequentially before the next attempt; every proven chunk
// doubles the next attempt (1024 rows after a 2048-row sequential prefix by default, up to 262144 rows).  On the C3
// bench data everything after row 2048 is proven that way, and one snapshot taken after 16k rows proves a single
// 184k-row chunk (tests/replay_proto.py: min margin 0.14 against a net displacement of 0.066).
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) replay_keys_kernel(const int64_t *__restrict__ top2_idx, int m, int K,
                                                          int *__restrict__ keys, int *__restrict__ vals) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) {
        const int64_t b = top2_idx[2 * (size_t)r];
        keys[r] = (b >= 0 && b < K) ? (int)b : K;   // a row without a nearest centroid (NaN row) goes to bucket K: no chain
        vals[r] = r;                                 // touches it, the certification fails on its missing candidates
    }
}

// seg_off[c] = first position of centroid c in the sorted key list (seg_off[K] = m)
__global__ void __launch_bounds__(256) replay_offsets_kernel(const int *__restrict__ keys_sorted, int m, int K,
                                                             int *__restrict__ seg_off) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > K) return;
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys_sorted[mid] < c) lo = mid + 1; else hi = mid;
    }
    seg_off[c] = lo;
}

__global__ void __launch_bounds__(256) replay_max_kernel(const double *__restrict__ v, int n, unsigned long long *__restrict__ out_bits) {
    double mx = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mx = fmax(mx, v[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx == mx) atomicMax(out_bits, (unsigned long long)__double_as_longlong(mx));
}

// One warp per centroid: its rows of the chunk, in row order.
__global__ void __launch_bounds__(128) replay_chain_kernel(const double *__restrict__ rows, int f,
                                                           const int *__restrict__ seg_off, const int *__restrict__ seg_rows,
                                                           int K, int saturated, double radius, double *cent,
                                                           const double *__restrict__ cent0, unsigned long long *sizes,
                                                           long long *__restrict__ assign, double *__restrict__ dcur,
                                                           unsigned long long *maxdisp_bits, int *fail) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= K) return;
    const int beg = seg_off[c], end = seg_off[c + 1];
    if (beg == end) return;
    double *cc = cent + (size_t)c * f;
    const double *c0 = cent0 + (size_t)c * f;
    unsigned long long cnt = sizes[c];
    const double guard = 1e-9 * radius;
    double dmax2 = 0.0;
    bool bad = false;
    const int lines = (f * 8 + 127) / 128;   // 128-byte lines per row
    for (int i = beg; i < end; ++i) {
        const int r = seg_rows[i];
        const double *x = rows + (size_t)r * f;
        if (i + 2 < end) {   // the chain is latency bound: pull the row after next towards the SM while this one is applied
            const char *nx = reinterpret_cast<const char *>(rows + (size_t)seg_rows[i + 2] * f);
            for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)l * 128));
        }
        double acc = 0.0;
        for (int j = lane; j < f; j += 32) {
            const double d = x[j] - cc[j];
            acc = fma(d, d, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double d2 = acc;   // accurate to ~1e-15 relative, NOT the reference's summation order: guard band below
        if (!(d2 == d2) || fabs(d2 - radius) <= guard || fabs(d2 - 1.5 * radius) <= guard ||
            (!saturated && fabs(d2 - 0.5 * radius) <= guard))
            bad = true;
        int cls;
        if (!saturated && d2 > 0.5 * radius) {   // the walk would open a new centroid here (:672): not replayable
            cls = 3;
            bad = true;
        } else if (d2 <= radius) {
            cls = 0;
        } else if (d2 <= 1.5 * radius) {
            cls = 1;
        } else {
            cls = 2;
        }
        if (lane == 0) {
            dcur[r] = sqrt(d2);
            assign[r] = cls == 2 ? -1ll : (long long)c;
        }
        if (cls == 0) {
            cnt += 1;
            const double k = (double)cnt;
            double dsp = 0.0;
            for (int j = lane; j < f; j += 32) {
                const double cj = cc[j];
                const double nj = __dadd_rn(cj, __ddiv_rn(__dsub_rn(x[j], cj), k));   // clustering.rs:747-751
                cc[j] = nj;
                const double e = nj - c0[j];
                dsp = fma(e, e, dsp);
            }
            for (int o = 16; o > 0; o >>= 1) dsp += __shfl_xor_sync(0xffffffffu, dsp, o);
            dmax2 = fmax(dmax2, dsp);
        } else if (cls == 1) {
            cnt += 1;
        }
    }
    if (lane == 0) {
        sizes[c] = cnt;
        const double dm = sqrt(dmax2) * (1.0 + 1e-12);
        if (dm == dm) atomicMax(maxdisp_bits, (unsigned long long)__double_as_longlong(dm));
        else bad = true;
        if (bad) atomicOr(fail, 1);
    }
}

// RN(a / b) from y = RN(1 / b) with two FMA correction steps (Markstein; the same routine the sequential kernel uses,
// cluster_f32p.cuh): correctly rounded for an integer divisor below 2^52 as long as nothing under- or overflows --
// operands outside [1e-280, 1e280] (and non-finite ones) take the library division.
__device__ __forceinline__ double replay_div_by_count(double a, double b, double y) {
    const double aa = fabs(a);
    if (!(aa >= 1e-280 && aa <= 1e280)) return a == 0.0 ? a / b : __ddiv_rn(a, b);
    const double q0 = __dmul_rn(a, y);
    const double q1 = __fma_rn(__fma_rn(-q0, b, a), y, q0);
    return __fma_rn(__fma_rn(-q1, b, a), y, q1);
}

// The same chain with the centroid, its snapshot and the NEXT row held in registers (f <= 32 NPL): the loads of row
// i + 1 are in flight while row i is applied, so a chain step costs its arithmetic, not a memory round trip.
template <int NPL>
__global__ void __launch_bounds__(128) replay_chain_reg_kernel(const double *__restrict__ rows, int f,
                                                               const int *__restrict__ seg_off,
                                                               const int *__restrict__ seg_rows, int K, int saturated,
                                                               double radius, double *__restrict__ cent,
                                                               const double *__restrict__ cent0, unsigned long long *sizes,
                                                               long long *__restrict__ assign, double *__restrict__ dcur,
                                                               unsigned long long *maxdisp_bits, int *fail) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= K) return;
    const int beg = seg_off[c], end = seg_off[c + 1];
    if (beg == end) return;
    double cr[NPL], s0[NPL], xn[NPL];
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int j = lane + 32 * u;
        cr[u] = s0[u] = j < f ? cent0[(size_t)c * f + j] : 0.0;   // padding lanes hold zeros everywhere: no effect
    }
    int r_next = seg_rows[beg];
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int j = lane + 32 * u;
        xn[u] = j < f ? rows[(size_t)r_next * f + j] : 0.0;
    }
    unsigned long long cnt = sizes[c];
    const double guard = 1e-9 * radius;
    const int lines = (f * 8 + 127) / 128;
    double dmax2 = 0.0;
    bool bad = false;
    for (int i = beg; i < end; ++i) {
        const int r = r_next;
        double x[NPL];
#pragma unroll
        for (int u = 0; u < NPL; ++u) x[u] = xn[u];
        if (i + 1 < end) {
            r_next = seg_rows[i + 1];
#pragma unroll
            for (int u = 0; u < NPL; ++u) {
                const int j = lane + 32 * u;
                xn[u] = j < f ? rows[(size_t)r_next * f + j] : 0.0;
            }
        }
        if (i + 4 < end) {
            const char *nx = reinterpret_cast<const char *>(rows + (size_t)seg_rows[i + 4] * f);
            for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)l * 128));
        }
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            const double d = x[u] - cr[u];
            acc = fma(d, d, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double d2 = acc;
        if (!(d2 == d2) || fabs(d2 - radius) <= guard || fabs(d2 - 1.5 * radius) <= guard ||
            (!saturated && fabs(d2 - 0.5 * radius) <= guard))
            bad = true;
        int cls;
        if (!saturated && d2 > 0.5 * radius) {
            cls = 3;
            bad = true;
        } else if (d2 <= radius) {
            cls = 0;
        } else if (d2 <= 1.5 * radius) {
            cls = 1;
        } else {
            cls = 2;
        }
        if (lane == 0) {
            dcur[r] = sqrt(d2);
            assign[r] = cls == 2 ? -1ll : (long long)c;
        }
        if (cls == 0) {
            cnt += 1;
            const double k = (double)cnt;
            const double y = __drcp_rn(k);
            double dsp = 0.0;
#pragma unroll
            for (int u = 0; u < NPL; ++u) {
                cr[u] = __dadd_rn(cr[u], replay_div_by_count(__dsub_rn(x[u], cr[u]), k, y));   // clustering.rs:747-751
                const double e = cr[u] - s0[u];
                dsp = fma(e, e, dsp);
            }
            for (int o = 16; o > 0; o >>= 1) dsp += __shfl_xor_sync(0xffffffffu, dsp, o);
            dmax2 = fmax(dmax2, dsp);
        } else if (cls == 1) {
            cnt += 1;
        }
    }
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int j = lane + 32 * u;
        if (j < f) cent[(size_t)c * f + j] = cr[u];
    }
    if (lane == 0) {
        sizes[c] = cnt;
        const double dm = sqrt(dmax2) * (1.0 + 1e-12);
        if (dm == dm) atomicMax(maxdisp_bits, (unsigned long long)__double_as_longlong(dm));
        else bad = true;
        if (bad) atomicOr(fail, 1);
    }
}

__global__ void __launch_bounds__(256) replay_certify_kernel(const double *__restrict__ dcur, const double *__restrict__ top2_dist,
                                                             const long long *__restrict__ top2_cnt,
                                                             const double *__restrict__ qn2, const unsigned long long *cn2max_bits,
                                                             const unsigned long long *maxdisp_bits,
                                                             const long long *__restrict__ assign, double radius, int f,
                                                             int m, int *fail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const double cn2max = __longlong_as_double((long long)*cn2max_bits);
    const double maxdisp = __longlong_as_double((long long)*maxdisp_bits);
    // GEMM-form squared distance |q|^2 + |c|^2 - 2 q.c from the FP64 tensor pipe: absolute error below e2
    const double e2 = 1e-15 * (double)(f + 32) * (qn2[r] + cn2max);
    const double s = top2_dist[2 * (size_t)r + 1];
    const double lower = sqrt(fmax(s * s - e2, 0.0)) - maxdisp;
    // A row farther than sqrt(1.5 radius) from EVERY centroid at its own time is dropped whichever centroid is the
    // nearest (no update, no count, assignment None: clustering.rs:759-815) -- e.g. a blob no centroid was opened for.
    const double b = top2_dist[2 * (size_t)r];
    const double lb = fmax(sqrt(fmax(b * b - e2, 0.0)) - maxdisp, 0.0);
    const bool dropped_anyway = lb * lb > 1.5 * radius * (1.0 + 1e-9) && assign[r] == -1;
    if (top2_cnt[r] < 2 || !(dcur[r] < lower || dropped_anyway)) atomicOr(fail, 2);
}

double opt_or(asb_ctx *ctx, const char *key, double dflt) {
    auto it = ctx->options.find(key);
    return it == ctx->options.end() ? dflt : it->second;
}

// elapsed time of the most recent launch bracketed by KernelTimer(name); the stream must be idle
double ktimer_ms(asb_ctx *ctx, const char *name) {
    auto it = ctx->ktimers.find(name);
    if (it == ctx->ktimers.end() || !it->second.a || !it->second.b) return 0.0;
    float ms = 0.f;
    if (cudaEventSynchronize(it->second.b) != cudaSuccess || cudaEventElapsedTime(&ms, it->second.a, it->second.b) != cudaSuccess) {
        cudaGetLastError();
        return 0.0;
    }
    return ms;
}

struct ReplayWs {
    DevTmp<double> qn2, xn2, dist, dcur, cent_tmp;
    DevTmp<int64_t> idx, cnt, minus1;
    DevTmp<int> keys, vals, keys_s, vals_s, seg_off, flags;
    DevTmp<unsigned long long> sizes_tmp, scal;   // scal[0] = max |c|^2 bits, scal[1] = max displacement bits
    DevTmp<unsigned char> cub_tmp;
    size_t cub_bytes = 0;
    int cap_m = 0, cap_k = 0;
};

int replay_ws_init(asb_ctx *ctx, ReplayWs &w, int m, int K, int f) {
    w.cap_m = m;
    w.cap_k = K;
    ASB_TRY(w.qn2.init(ctx, (size_t)m));
    ASB_TRY(w.xn2.init(ctx, (size_t)K));
    ASB_TRY(w.dist.init(ctx, (size_t)m * 2));
    ASB_TRY(w.dcur.init(ctx, (size_t)m));
    ASB_TRY(w.cent_tmp.init(ctx, (size_t)K * f));
    ASB_TRY(w.idx.init(ctx, (size_t)m * 2));
    ASB_TRY(w.cnt.init(ctx, (size_t)m));
    ASB_TRY(w.minus1.init(ctx, (size_t)m));
    ASB_TRY(w.keys.init(ctx, (size_t)m));
    ASB_TRY(w.vals.init(ctx, (size_t)m));
    ASB_TRY(w.keys_s.init(ctx, (size_t)m));
    ASB_TRY(w.vals_s.init(ctx, (size_t)m));
    ASB_TRY(w.seg_off.init(ctx, (size_t)K + 1));
    ASB_TRY(w.flags.init(ctx, 2));
    ASB_TRY(w.sizes_tmp.init(ctx, (size_t)K));
    ASB_TRY(w.scal.init(ctx, 2));
    ASB_CUDA(ctx, cudaMemsetAsync(w.minus1.ptr, 0xff, (size_t)m * sizeof(int64_t), ctx->stream));
    w.cub_bytes = 0;
    ASB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, w.keys.ptr, w.keys_s.ptr, w.vals.ptr, w.vals_s.ptr, m,
                                                  0, 32, ctx->stream));
    ASB_TRY(w.cub_tmp.init(ctx, w.cub_bytes));
    return ASB_OK;
}

// One chunk.  *ok = 1: centroids / sizes / assign hold the walk's state after the chunk; 0: nothing was touched
// except assign (the caller's sequential walk rewrites it).
int replay_chunk(asb_ctx *ctx, ReplayWs &w, const double *rows_d, int m, int f, int K, int saturated, double radius,
                 double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d, int *ok) {
    *ok = 0;
    ASB_CUDA(ctx, cudaMemsetAsync(w.flags.ptr, 0, 2 * sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(w.scal.ptr, 0, 2 * sizeof(unsigned long long), ctx->stream));
    ASB_TRY(asb_dev_norms2(ctx, rows_d, m, f, w.qn2.ptr));
    ASB_TRY(asb_dev_norms2(ctx, centroids_d, K, f, w.xn2.ptr));
    replay_max_kernel<<<1, 256, 0, ctx->stream>>>(w.xn2.ptr, K, w.scal.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_max_kernel"));
    ASB_TRY(asb_dev_top2_l2(ctx, rows_d, m, f, centroids_d, K, w.qn2.ptr, w.xn2.ptr, w.minus1.ptr, w.idx.ptr, w.dist.ptr,
                            w.cnt.ptr, w.flags.ptr + 1));
    replay_keys_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(w.idx.ptr, m, K, w.keys.ptr, w.vals.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_keys_kernel"));
    int bits = 1;
    while ((1ll << bits) < (long long)K + 1) ++bits;
    ASB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(w.cub_tmp.ptr, w.cub_bytes, w.keys.ptr, w.keys_s.ptr, w.vals.ptr,
                                                  w.vals_s.ptr, m, 0, bits, ctx->stream));   // stable: rows stay ascending
    ctx->launches++;
    replay_offsets_kernel<<<(K + 1 + 255) / 256, 256, 0, ctx->stream>>>(w.keys_s.ptr, m, K, w.seg_off.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_offsets_kernel"));
    ASB_CUDA(ctx, cudaMemcpyAsync(w.cent_tmp.ptr, centroids_d, (size_t)K * f * sizeof(double), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(w.sizes_tmp.ptr, sizes_d, (size_t)K * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    {
        KernelTimer kt(ctx, "cluster_chain_kernel");
        const bool generic = f > 512 || opt_or(ctx, "cluster_replay_generic_chain", 0.0) != 0.0;
#define ASB_CHAIN_ARGS                                                                                                  \
    rows_d, f, w.seg_off.ptr, w.vals_s.ptr, K, saturated, radius, w.cent_tmp.ptr, centroids_d, w.sizes_tmp.ptr,          \
        (long long *)assign_d, w.dcur.ptr, w.scal.ptr + 1, w.flags.ptr
        const unsigned grid = (unsigned)((K + 3) / 4);
        if (generic) replay_chain_kernel<<<grid, 128, 0, ctx->stream>>>(ASB_CHAIN_ARGS);
        else if (f <= 128) replay_chain_reg_kernel<4><<<grid, 128, 0, ctx->stream>>>(ASB_CHAIN_ARGS);
        else if (f <= 256) replay_chain_reg_kernel<8><<<grid, 128, 0, ctx->stream>>>(ASB_CHAIN_ARGS);
        else if (f <= 384) replay_chain_reg_kernel<12><<<grid, 128, 0, ctx->stream>>>(ASB_CHAIN_ARGS);
        else replay_chain_reg_kernel<16><<<grid, 128, 0, ctx->stream>>>(ASB_CHAIN_ARGS);
#undef ASB_CHAIN_ARGS
    }
    ASB_TRY(asb_check_launch(ctx, "replay_chain_kernel"));
    replay_certify_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(w.dcur.ptr, w.dist.ptr, (const long long *)w.cnt.ptr,
                                                                    w.qn2.ptr, w.scal.ptr, w.scal.ptr + 1,
                                                                    (const long long *)assign_d, radius, f, m, w.flags.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_certify_kernel"));
    int hflags[2] = {0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(hflags, w.flags.ptr, sizeof(hflags), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hflags[0] != 0 || hflags[1] != 0) return ASB_OK;
    ASB_CUDA(ctx, cudaMemcpyAsync(centroids_d, w.cent_tmp.ptr, (size_t)K * f * sizeof(double), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(sizes_d, w.sizes_tmp.ptr, (size_t)K * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    *ok = 1;
    return ASB_OK;
}

}  // namespace

int asb_dev_cluster(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int64_t max_clusters, double radius,
                    double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d, int64_t *x_out_host,
                    int64_t init_k) {
    const int64_t prefix = (int64_t)opt_or(ctx, "cluster_replay_prefix", 2048.0);
    const int64_t chunk = (int64_t)opt_or(ctx, "cluster_replay_chunk", 1024.0);
    int64_t chunk_max = (int64_t)opt_or(ctx, "cluster_replay_chunk_max", 262144.0);
    if (chunk_max < chunk) chunk_max = chunk;
    const bool replay = opt_or(ctx, "cluster_replay", 1.0) != 0.0 && prefix >= 1 && chunk >= 256 && chunk_max <= (1 << 24) &&
                        n >= prefix + chunk / 4 && f <= 16384 && max_clusters * f <= (1ll << 27);
    ctx->kernel_ms["cluster_replay_chunks"] = 0.0;
    ctx->kernel_ms["cluster_replay_chunks_ok"] = 0.0;
    ctx->kernel_ms["cluster_replay_rows"] = 0.0;
    if (!replay)
        return asb_dev_cluster_seq(ctx, rows_d, n, f, max_clusters, radius, centroids_d, assign_d, sizes_d, x_out_host, init_k);

    int64_t x = init_k;
    double ms_seq = 0.0, ms_top2 = 0.0, ms_chain = 0.0;   // device time per part, summed over the chunks
    ASB_TRY(asb_dev_cluster_seq(ctx, rows_d, prefix, f, max_clusters, radius, centroids_d, assign_d, sizes_d, &x, init_k));
    ms_seq += ktimer_ms(ctx, "cluster_kernel");
    ReplayWs w;
    bool ws_ready = false;
    int fails = 0, tried = 0, proven = 0;
    int64_t rows_replayed = 0;
    int64_t lo = prefix;
    int64_t cur = chunk;   // rows per attempt: doubles after every proven chunk (one snapshot holds for longer and
                           // longer stretches as the centroids settle), back to `chunk` after a failure
    while (lo < n) {
        int64_t hi = lo + cur < n ? lo + cur : n;
        int ok = 0;
        if (x >= 2) {
            if (!ws_ready || w.cap_k < x || w.cap_m < hi - lo) {
                ASB_TRY(replay_ws_init(ctx, w, (int)cur, (int)(x > max_clusters ? x : max_clusters), (int)f));
                ws_ready = true;
            }
            ++tried;
            ASB_TRY(replay_chunk(ctx, w, rows_d + lo * f, (int)(hi - lo), (int)f, (int)x, x >= max_clusters ? 1 : 0, radius,
                                 centroids_d, assign_d + lo, sizes_d, &ok));
            ms_top2 += ktimer_ms(ctx, "cluster_top2_kernel");
            ms_chain += ktimer_ms(ctx, "cluster_chain_kernel");
        }
        if (ok) {
            fails = 0;
            ++proven;
            rows_replayed += hi - lo;
            cur = cur * 2 < chunk_max ? cur * 2 : chunk_max;
        } else {
            // not provable (yet): walk sequentially, and twice as far after every further failure in a row -- data that
            // never settles costs O(log(n / chunk)) wasted attempts, data that settles late is picked up when it does
            ++fails;
            const int64_t span = chunk << (fails - 1 < 12 ? fails - 1 : 12);
            hi = lo + span < n ? lo + span : n;
            cur = chunk;
            const int64_t x_before = x;
            ASB_TRY(asb_dev_cluster_seq(ctx, rows_d + lo * f, hi - lo, f, max_clusters, radius, centroids_d, assign_d + lo,
                                        sizes_d, &x, x_before));
            ms_seq += ktimer_ms(ctx, "cluster_kernel");
        }
        lo = hi;
    }
    *x_out_host = x;
    ctx->kernel_ms["cluster_replay_chunks"] = (double)tried;
    ctx->kernel_ms["cluster_replay_chunks_ok"] = (double)proven;
    ctx->kernel_ms["cluster_replay_rows"] = (double)rows_replayed;
    ctx->kernel_ms["cluster_replay_seq_ms"] = ms_seq;
    ctx->kernel_ms["cluster_replay_top2_ms"] = ms_top2;
    ctx->kernel_ms["cluster_replay_chain_ms"] = ms_chain;
    return ASB_OK;
}
