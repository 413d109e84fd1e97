// taumode.cu -- K5+K6+K7: per-item tau selection fused with the Rayleigh / dispersion sums.
//
// Replaces TauMode::select_tau (src/taumode.rs:87-127) and compute_synthetic_lambda_csr
// (src/taumode.rs:552-660) as driven by compute_taumode_lambdas_parallel (:174-312) and
// ArrowSpace::prepare_query_item (src/core.rs:533-549).
//
// Layout: a CTA owns a tile of TI items.  The tile is read once from HBM (coalesced along
// features) and stored TRANSPOSED in shared memory, X[feature][item] with an odd pitch, so
// that "lane = item" makes every later access conflict free: for a graph entry (i, j, L_ij)
// the warp reads X[i][*] and X[j][*] as two contiguous 256 B rows while (i, j, L_ij) itself
// is warp-uniform (one 16 B read-only load).  This is the batched SpMM  L * X^T  fused with
// the row dots; the F x F graph never leaves L1/L2 and HBM traffic is the algorithmic
// 8*F + 16 bytes per item.
//
// The 8 warps of the CTA split the graph rows (row i -> partition i mod P); partial sums are
// combined through a small shared scratch.  tau (median / percentile) is an exact order
// statistic found by counting passes over the same transposed tile (no sort): each pass
// counts the finite values below a pivot and tracks the largest value below / smallest value
// at-or-above it with multiplicities, which resolves every rank inside those two groups.
// Pivots are interpolated inside the current bracket; after 8 passes the bracket is bisected
// in ordered-key space, so termination is guaranteed for any input.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr double kTauFloor = 1e-10;  // src/taumode.rs:84

__device__ __forceinline__ unsigned long long ord_key(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

template <int TI>
__global__ void __launch_bounds__(kThreads)
taumode_kernel(const double *__restrict__ items, long long n, int f,
               const GraphEntry *__restrict__ entries, const int *__restrict__ row_ptr, int tau_mode,
               double tau_value, double *__restrict__ lambdas, double *__restrict__ norms2,
               int *__restrict__ nonfinite_flag) {
    constexpr int PITCH = TI + 1;
    constexpr int G = 32 / TI;      // row partitions per warp
    constexpr int P = kWarps * G;   // row partitions per CTA
    extern __shared__ double smem[];
    double *X = smem;                          // f * PITCH
    double *red = smem + (size_t)f * PITCH;    // 4 * P * TI

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int item = lane % TI;
    const int part = warp * G + lane / TI;
    auto R = [&](int slot, int p, int it) -> double & { return red[(slot * P + p) * TI + it]; };

    const long long ntiles = (n + TI - 1) / TI;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TI;
        __syncthreads();  // previous tile fully consumed
        // ---- load: warp w brings items w, w+8, ... ; lanes run along features (coalesced)
        for (int it = warp; it < TI; it += kWarps) {
            const long long row = base + it;
            const bool valid = row < n;
            const double *src = items + row * (long long)f;
#pragma unroll 4
            for (int j = lane; j < f; j += 32) {
                double v = valid ? __ldg(src + j) : 0.0;
                X[j * PITCH + it] = v;
            }
        }
        __syncthreads();

        // ---- pass 0: Rayleigh / dispersion sums + value statistics over this partition's rows
        double den = 0.0, num = 0.0, s1 = 0.0, s2 = 0.0, vsum = 0.0;
        double vmin = INFINITY, vmax = -INFINITY;
        int cnt = 0;
        for (int i = part; i < f; i += P) {
            const double xi = X[i * PITCH + item];
            den = fma(xi, xi, den);
            if (fabs(xi) < INFINITY) {  // is_finite
                cnt++;
                vsum += xi;
                vmin = fmin(vmin, xi);
                vmax = fmax(vmax, xi);
            }
            const int e0 = row_ptr[i], e1 = row_ptr[i + 1];
            for (int e = e0; e < e1; ++e) {
                const GraphEntry ge = entries[e];
                const double xj = X[ge.j * PITCH + item];
                num = fma(xi * ge.v, xj, num);  // x_i * L_ij * x_j   (taumode.rs:575)
                if (ge.j != i && ge.v < 0.0) {  // w = max(-L_ij, 0) > 0  (:577-583)
                    const double d = xi - xj;
                    const double c = (-ge.v * d) * d;
                    s1 += c;
                    s2 = fma(c, c, s2);
                }
            }
        }
        // reduce in two rounds of 4 slots
        R(0, part, item) = den;
        R(1, part, item) = num;
        R(2, part, item) = s1;
        R(3, part, item) = s2;
        __syncthreads();
        den = num = s1 = s2 = 0.0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            den += R(0, p, item);
            num += R(1, p, item);
            s1 += R(2, p, item);
            s2 += R(3, p, item);
        }
        __syncthreads();
        R(0, part, item) = vsum;
        R(1, part, item) = vmin;
        R(2, part, item) = vmax;
        R(3, part, item) = (double)cnt;
        __syncthreads();
        vsum = 0.0;
        vmin = INFINITY;
        vmax = -INFINITY;
        double cntd = 0.0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            vsum += R(0, p, item);
            vmin = fmin(vmin, R(1, p, item));
            vmax = fmax(vmax, R(2, p, item));
            cntd += R(3, p, item);
        }
        const int nf = (int)cntd;
        __syncthreads();

        // ---- tau (src/taumode.rs:87-127)
        double tau;
        bool done = true;
        int rb = 0;
        bool need_pair = false;
        if (tau_mode == ASB_TAU_FIXED) {
            tau = (fabs(tau_value) < INFINITY && tau_value > 0.0) ? tau_value : kTauFloor;
        } else if (tau_mode == ASB_TAU_MEAN) {
            const double m = nf > 0 ? vsum / (double)nf : 0.0;
            tau = fmax(m, kTauFloor);
        } else {
            tau = kTauFloor;  // empty -> floor
            if (nf > 0) {
                done = false;
                if (tau_mode == ASB_TAU_PERCENTILE) {
                    double pp = tau_value;
                    pp = pp < 0.0 ? 0.0 : (pp > 1.0 ? 1.0 : pp);
                    const double fi = round((double)(nf - 1) * pp);  // half away from zero
                    rb = (fi != fi || fi < 0.0) ? 0 : (int)fi;
                    if (rb > nf - 1) rb = nf - 1;
                } else {
                    rb = nf / 2;
                    need_pair = (nf % 2 == 0);
                }
            }
        }
        double lo = -INFINITY, hi = INFINITY;
        int n_le_lo = 0, n_lt_hi = nf;
        int target = rb;
        int phase = 0;
        double v_rb = 0.0, v_ra = 0.0;
        for (int iter = 0;; ++iter) {
            if (!__syncthreads_or(!done)) break;
            if (iter > 200 && !done) {  // cannot happen (bisection bounds the passes); never hang the GPU
                done = true;
                tau = __longlong_as_double(0x7ff8000000000000ll);
            }
            double pv = 0.0;
            if (!done) {
                const double lo_f = (lo == -INFINITY) ? vmin : lo;
                const double hi_f = (hi == INFINITY) ? vmax : hi;
                const int width = n_lt_hi - n_le_lo;
                const double frac = ((double)(target - n_le_lo) + 0.5) / (double)(width > 0 ? width : 1);
                pv = lo_f + (hi_f - lo_f) * frac;
                if (iter >= 8 || !(fabs(pv) < INFINITY)) {  // bisection in ordered-key space
                    const unsigned long long kl = ord_key(lo_f), kh = ord_key(hi_f);
                    pv = ord_val(kl + ((kh - kl) >> 1) + ((kh - kl) & 1ull));
                }
                if (lo != -INFINITY && !(pv > lo)) pv = ord_val(ord_key(lo) + 1ull);
                if (pv > hi_f) pv = hi_f;
            }
            int c_lt = 0, mb = 0, ma = 0;
            double below = -INFINITY, above = INFINITY;
            if (!done) {
                for (int i = part; i < f; i += P) {
                    const double x = X[i * PITCH + item];
                    if (fabs(x) < INFINITY) {
                        if (x < pv) {
                            c_lt++;
                            if (x > below) {
                                below = x;
                                mb = 1;
                            } else if (x == below) {
                                mb++;
                            }
                        } else {
                            if (x < above) {
                                above = x;
                                ma = 1;
                            } else if (x == above) {
                                ma++;
                            }
                        }
                    }
                }
            }
            R(0, part, item) = below;
            R(1, part, item) = above;
            R(2, part, item) =
                __longlong_as_double((long long)c_lt | ((long long)mb << 20) | ((long long)ma << 40));
            __syncthreads();
            if (!done) {
                double gb = -INFINITY, ga = INFINITY;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    gb = fmax(gb, R(0, p, item));
                    ga = fmin(ga, R(1, p, item));
                }
                c_lt = 0;
                mb = 0;
                ma = 0;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const long long pk = __double_as_longlong(R(2, p, item));
                    c_lt += (int)(pk & 0xfffff);
                    if (R(0, p, item) == gb) mb += (int)((pk >> 20) & 0xfffff);
                    if (R(1, p, item) == ga) ma += (int)((pk >> 40) & 0xfffff);
                }
                if (gb == -INFINITY) mb = 0;
                if (ga == INFINITY) ma = 0;
                bool found = false;
                bool in_below = false;
                double fv = 0.0;
                if (target < c_lt - mb) {
                    hi = gb;
                    n_lt_hi = c_lt - mb;
                } else if (target < c_lt) {
                    found = true;
                    in_below = true;
                    fv = gb;
                } else if (target < c_lt + ma) {
                    found = true;
                    fv = ga;
                } else {
                    lo = ga;
                    n_le_lo = c_lt + ma;
                }
                if (found) {
                    if (phase == 1) {
                        v_ra = fv;
                        done = true;
                    } else {
                        v_rb = fv;
                        if (!need_pair) {
                            done = true;
                        } else {
                            const int ra = rb - 1;
                            if (!in_below) {
                                v_ra = (ra >= c_lt) ? ga : gb;
                                done = true;
                            } else if (ra >= c_lt - mb) {
                                v_ra = gb;
                                done = true;
                            } else if (n_le_lo == rb) {
                                v_ra = lo;
                                done = true;
                            } else {
                                phase = 1;
                                target = ra;
                                hi = gb;
                                n_lt_hi = c_lt - mb;
                            }
                        }
                    }
                    if (done) {
                        const double m = need_pair ? 0.5 * (v_ra + v_rb) : v_rb;  // :119-124
                        tau = fmax(m, kTauFloor);
                    }
                }
            }
            __syncthreads();
        }

        // ---- lambda (src/taumode.rs:596-647)
        if (part == 0 && base + item < n) {
            const double e_raw = den > 1e-12 ? num / den : 0.0;
            double g = 0.0;
            if (s1 > 0.0) {
                g = s2 / (s1 * s1);  // sum (c/edge)^2
                g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
            }
            const double e_bounded = e_raw / (e_raw + tau);
            lambdas[base + item] = tau * e_bounded + (1.0 - tau) * g;
            if (norms2) norms2[base + item] = den;
            if (nonfinite_flag && nf < f) atomicOr(nonfinite_flag, 1);
        }
    }
}

// min / max / sum of lambdas (src/eigenmaps.rs:372-382): single CTA, deterministic order.
__global__ void __launch_bounds__(1024) lambda_stats_kernel(const double *__restrict__ lam, long long n,
                                                            double *__restrict__ out) {
    __shared__ double smin[32], smax[32], ssum[32];
    double mn = INFINITY, mx = -INFINITY, sm = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = lam[i];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        sm += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        smin[w] = mn;
        smax[w] = mx;
        ssum[w] = sm;
    }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        mn = l < nw ? smin[l] : INFINITY;
        mx = l < nw ? smax[l] : -INFINITY;
        sm = l < nw ? ssum[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
        }
        if (l == 0) {
            out[0] = mn;
            out[1] = mx;
            out[2] = sm;
        }
    }
}

// sum(x^2) per row: one warp per row, coalesced.
__global__ void __launch_bounds__(256) norms2_kernel(const double *__restrict__ rows, long long n, int f,
                                                     double *__restrict__ out) {
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const double *r = rows + w * (long long)f;
    double s = 0.0;
    for (int j = lane; j < f; j += 32) {
        const double v = __ldg(r + j);
        s = fma(v, v, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[w] = s;
}

}  // namespace

#include "taumode_sym.cuh"

namespace {

template <int TI>
int launch_taumode(asb_ctx *ctx, const double *items_d, int64_t n, int f, const GraphPlan &plan, int tau_mode,
                   double tau_value, double *lambdas_d, double *norms2_d, int *flag_d) {
    // flag_d != nullptr <=> query preparation (K7); timed under its own name
    const char *timer_name = flag_d ? "query_taumode_kernel" : "taumode_kernel";
    constexpr int P = kWarps * (32 / TI);
    const size_t smem = ((size_t)f * (TI + 1) + (size_t)4 * P * TI) * sizeof(double);
    const void *kern = (const void *)taumode_kernel<TI>;
    ASB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    ASB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "taumode tile does not fit (f=%d)", f);
    const long long ntiles = (n + TI - 1) / TI;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    {
        KernelTimer kt(ctx, timer_name);
        taumode_kernel<TI><<<(unsigned)grid, kThreads, smem, ctx->stream>>>(
            items_d, (long long)n, f, plan.entries, plan.row_ptr, tau_mode, tau_value, lambdas_d, norms2_d, flag_d);
    }
    return asb_check_launch(ctx, "taumode_kernel");
}

}  // namespace

int asb_graph_plan_from_host(asb_ctx *ctx, const int64_t *indptr, const int64_t *indices, const double *data,
                             int64_t f, GraphPlan *plan) {
    if (f <= 0 || f > (1 << 19)) ASB_FAIL(ctx, ASB_ERR_INVALID, "graph plan: bad f=%lld", (long long)f);
    const int64_t nnz = indptr[f];
    if (indptr[0] != 0 || nnz < 0 || nnz > (int64_t)1 << 30)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "graph plan: bad indptr (nnz=%lld)", (long long)nnz);
    std::vector<GraphEntry> ent((size_t)nnz);
    std::vector<int32_t> rp((size_t)f + 1);
    for (int64_t i = 0; i < f; ++i) {
        if (indptr[i + 1] < indptr[i]) ASB_FAIL(ctx, ASB_ERR_INVALID, "graph plan: indptr not monotone");
        rp[i] = (int32_t)indptr[i];
        for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
            if (indices[e] < 0 || indices[e] >= f)
                ASB_FAIL(ctx, ASB_ERR_INVALID, "graph plan: column %lld out of range", (long long)indices[e]);
            ent[e].v = data[e];
            ent[e].j = (int32_t)indices[e];
            ent[e].row = (int32_t)i;
        }
    }
    rp[f] = (int32_t)nnz;
    plan->f = f;
    plan->nnz = nnz;
    plan->stream = ctx->stream;
    ASB_CUDA(ctx, cudaMallocAsync((void **)&plan->entries, (nnz > 0 ? nnz : 1) * sizeof(GraphEntry), ctx->stream));
    ASB_CUDA(ctx, cudaMallocAsync((void **)&plan->row_ptr, (f + 1) * sizeof(int32_t), ctx->stream));
    if (nnz > 0)
        ASB_CUDA(ctx, cudaMemcpyAsync(plan->entries, ent.data(), nnz * sizeof(GraphEntry), cudaMemcpyHostToDevice,
                                      ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(plan->row_ptr, rp.data(), (f + 1) * sizeof(int32_t), cudaMemcpyHostToDevice,
                                  ctx->stream));
    // ---- symmetric form: is L_ij == L_ji bit for bit for every stored off-diagonal?
    std::vector<SymEdge> sym;
    std::vector<double> resid((size_t)f, 0.0);
    bool is_sym = true, all_pos = true;
    auto find = [&](int64_t r, int64_t c, double *out) -> bool {  // rows are short; linear scan
        for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e)
            if (indices[e] == c) {
                *out = data[e];
                return true;
            }
        return false;
    };
    // error-free accumulation (TwoSum cascade): the residual r_i = L_ii + sum_{j != i} L_ij of a STORED Laplacian row is
    // not zero -- the stored diagonal is the rounded degree -- and x^T L x of the stored matrix contains sum r_i x_i^2.
    // Computing r_i in plain floating point would return the rounding noise of the subtraction instead of the value.
    auto two_sum = [](double a, double b, double &err) {
        const double s = a + b;
        const double bb = s - a;
        err = (a - (s - bb)) + (b - bb);
        return s;
    };
    for (int64_t i = 0; i < f && is_sym; ++i) {
        double diag = 0.0, wsum = 0.0;
        double rs = 0.0, rc = 0.0;   // residual: running sum and its accumulated rounding error
        int64_t last_col = -1;
        for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
            const int64_t j = indices[e];
            if (j <= last_col) is_sym = false;  // duplicates / unsorted: use the generic kernel
            last_col = j;
            if (j == i) {
                diag = data[e];
                continue;
            }
            double back;
            if (!find(j, i, &back) || memcmp(&back, &data[e], sizeof(double)) != 0) {
                is_sym = false;
                break;
            }
            const double w = -data[e];
            wsum += w;  // ascending j, like the degree sum of src/laplacian.rs:369
            {
                double e1;
                rs = two_sum(rs, data[e], e1);
                rc += e1;
            }
            if (!(w > 0.0)) all_pos = false;
            if (j > i) {
                SymEdge se;
                se.w = w;
                se.i = (int)i;
                se.j = (int)j;
                sym.push_back(se);
            }
        }
        {
            double e1;
            rs = two_sum(rs, diag, e1);
            rc += e1;
        }
        (void)wsum;
        resid[i] = rs + rc;   // exact to ~1e-32 of the row's magnitude
    }
    plan->is_sym = is_sym;
    plan->all_pos = is_sym && all_pos;
    std::vector<SymEdge> sched;
    if (is_sym) asb_schedule_edges(sym, sched);  // conflict-free steps of 32 edges (padded)
    plan->nedges = is_sym ? (int64_t)sched.size() : 0;
    if (is_sym) {
        ASB_CUDA(ctx, cudaMallocAsync(&plan->sym_edges, (sched.size() > 0 ? sched.size() : 1) * sizeof(SymEdge), ctx->stream));
        ASB_CUDA(ctx, cudaMallocAsync((void **)&plan->resid, f * sizeof(double), ctx->stream));
        if (!sched.empty())
            ASB_CUDA(ctx, cudaMemcpyAsync(plan->sym_edges, sched.data(), sched.size() * sizeof(SymEdge),
                                          cudaMemcpyHostToDevice, ctx->stream));
        ASB_CUDA(ctx, cudaMemcpyAsync(plan->resid, resid.data(), f * sizeof(double), cudaMemcpyHostToDevice,
                                      ctx->stream));
    }
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    return ASB_OK;
}

// v4 (taumode_sym.cuh): item in registers + shared memory, IPP items per warp pass; f <= 32 NPL <= 1024
template <bool ALLPOS, int NPL, int IPP>
static int launch_taumode_reg(asb_ctx *ctx, const char *tname, const double *items_d, int64_t n, int f,
                              const GraphPlan &plan, int tau_mode, double tau_value, double *lambdas_d, double *norms2_d,
                              int *nonfinite_flag_d) {
    auto kern = taumode_reg_kernel<ALLPOS, NPL, IPP>;
    const size_t smem = (size_t)kTauWarps * (IPP * f + kSelCap) * sizeof(double);   // items + the selection scratch
    ASB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    ASB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTauWarps * 32, smem));
    if (per_sm < 1) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "taumode: item does not fit shared memory (f=%d)", f);
    long long grid = (long long)ctx->sm_count * per_sm;
    const long long need = ((n + IPP - 1) / IPP + kTauWarps - 1) / kTauWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    {
        KernelTimer kt(ctx, tname);
        kern<<<(unsigned)grid, kTauWarps * 32, smem, ctx->stream>>>(items_d, (long long)n, f, (const SymEdge *)plan.sym_edges,
                                                                    (int)(plan.nedges / 32), plan.resid, tau_mode, tau_value,
                                                                    lambdas_d, norms2_d, nonfinite_flag_d, (int)plan.f);
    }
    return asb_check_launch(ctx, "taumode_reg_kernel");
}

int asb_dev_taumode(asb_ctx *ctx, const double *items_d, int64_t n, int64_t f, const GraphPlan &plan,
                    int tau_mode, double tau_value, double *lambdas_d, double *norms2_d, double *stats_d,
                    int *nonfinite_flag_d) {
    if (n <= 0 || f <= 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "taumode: n=%lld f=%lld", (long long)n, (long long)f);
    if (plan.f > f)
        ASB_FAIL(ctx, ASB_ERR_DIM, "taumode: graph is %lldx%lld but items have %lld features", (long long)plan.f,
                 (long long)plan.f, (long long)f);
    if (tau_mode < ASB_TAU_FIXED || tau_mode > ASB_TAU_PERCENTILE)
        ASB_FAIL(ctx, ASB_ERR_INVALID, "taumode: unknown tau mode %d", tau_mode);
    bool generic = !plan.is_sym;
    {
        auto it = ctx->options.find("taumode_generic");
        if (it != ctx->options.end() && it->second != 0.0) generic = true;
    }
    if (plan.f < f && (generic || (size_t)f * sizeof(double) > 200 * 1024))
        ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "taumode: a graph smaller than the items (JL-projected build) needs the symmetric kernel");
    bool regs = !generic && f <= 1024;
    {
        auto it = ctx->options.find("taumode_regs");
        if (it != ctx->options.end() && it->second == 0.0) regs = false;
    }
    if (regs) {
        int rc = ASB_OK;
        const char *tname = nonfinite_flag_d ? "query_taumode_kernel" : "taumode_kernel";
#define ASB_TAU_REG(NPL_, IPP_)                                                                                       \
    rc = plan.all_pos ? launch_taumode_reg<true, NPL_, IPP_>(ctx, tname, items_d, n, (int)f, plan, tau_mode, tau_value,  \
                                                             lambdas_d, norms2_d, nonfinite_flag_d)                  \
                      : launch_taumode_reg<false, NPL_, IPP_>(ctx, tname, items_d, n, (int)f, plan, tau_mode, tau_value, \
                                                              lambdas_d, norms2_d, nonfinite_flag_d)
        int ipp = 1;  // items per warp pass: 1 (default: 64 registers, 32 warps per SM -- 3.1 ms per 1M x 384 against 3.6 ms
                      // with two items per pass at 128 registers), 2 on request (f <= 512)
        {
            auto it = ctx->options.find("taumode_ipp");
            if (it != ctx->options.end() && it->second == 2.0) ipp = 2;
        }
        if (f <= 128) { if (ipp == 2) ASB_TAU_REG(4, 2); else ASB_TAU_REG(4, 1); }
        else if (f <= 256) { if (ipp == 2) ASB_TAU_REG(8, 2); else ASB_TAU_REG(8, 1); }
        else if (f <= 384) { if (ipp == 2) ASB_TAU_REG(12, 2); else ASB_TAU_REG(12, 1); }
        else if (f <= 512) { if (ipp == 2) ASB_TAU_REG(16, 2); else ASB_TAU_REG(16, 1); }
        else if (f <= 768) ASB_TAU_REG(24, 1);
        else ASB_TAU_REG(32, 1);
#undef ASB_TAU_REG
        ASB_TRY(rc);
        if (stats_d) {
            lambda_stats_kernel<<<1, 1024, 0, ctx->stream>>>(lambdas_d, (long long)n, stats_d);
            ASB_TRY(asb_check_launch(ctx, "lambda_stats_kernel"));
        }
        return ASB_OK;
    }
    if (!generic && (size_t)f * sizeof(double) <= 200 * 1024) {
        // one warp per item; as many warps per CTA as the private shared-memory copies allow
        int wpc = kTauWarps;
        while (wpc > 1 && (size_t)wpc * f * sizeof(double) > 96 * 1024) wpc >>= 1;
        const size_t smem = (size_t)wpc * f * sizeof(double);
        const void *kern = plan.all_pos ? (const void *)taumode_warp_kernel<true> : (const void *)taumode_warp_kernel<false>;
        ASB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        ASB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTauWarps * 32, smem));
        if (per_sm < 1) ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "taumode: item does not fit shared memory (f=%lld)", (long long)f);
        long long grid = (long long)ctx->sm_count * per_sm;
        const long long need = (n + wpc - 1) / wpc;
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        const int nsteps = (int)(plan.nedges / 32);
        {
            KernelTimer kt(ctx, nonfinite_flag_d ? "query_taumode_kernel" : "taumode_kernel");
            if (plan.all_pos)
                taumode_warp_kernel<true><<<(unsigned)grid, kTauWarps * 32, smem, ctx->stream>>>(
                    items_d, (long long)n, (int)f, (const SymEdge *)plan.sym_edges, nsteps, plan.resid, tau_mode,
                    tau_value, lambdas_d, norms2_d, nonfinite_flag_d, wpc, (int)plan.f);
            else
                taumode_warp_kernel<false><<<(unsigned)grid, kTauWarps * 32, smem, ctx->stream>>>(
                    items_d, (long long)n, (int)f, (const SymEdge *)plan.sym_edges, nsteps, plan.resid, tau_mode,
                    tau_value, lambdas_d, norms2_d, nonfinite_flag_d, wpc, (int)plan.f);
        }
        ASB_TRY(asb_check_launch(ctx, "taumode_warp_kernel"));
        if (stats_d) {
            lambda_stats_kernel<<<1, 1024, 0, ctx->stream>>>(lambdas_d, (long long)n, stats_d);
            ASB_TRY(asb_check_launch(ctx, "lambda_stats_kernel"));
        }
        return ASB_OK;
    }
    const size_t budget = 227 * 1024 - 1024;
    auto fits = [&](int ti) { return ((size_t)f * (ti + 1) + 4 * 256) * sizeof(double) <= budget; };
    int rc;
    const int fi = (int)f;
    if (fits(32))
        rc = launch_taumode<32>(ctx, items_d, n, fi, plan, tau_mode, tau_value, lambdas_d, norms2_d, nonfinite_flag_d);
    else if (fits(16))
        rc = launch_taumode<16>(ctx, items_d, n, fi, plan, tau_mode, tau_value, lambdas_d, norms2_d, nonfinite_flag_d);
    else if (fits(8))
        rc = launch_taumode<8>(ctx, items_d, n, fi, plan, tau_mode, tau_value, lambdas_d, norms2_d, nonfinite_flag_d);
    else if (fits(4))
        rc = launch_taumode<4>(ctx, items_d, n, fi, plan, tau_mode, tau_value, lambdas_d, norms2_d, nonfinite_flag_d);
    else
        ASB_FAIL(ctx, ASB_ERR_UNSUPPORTED, "taumode: f=%lld exceeds the on-chip tile (max ~5500)", (long long)f);
    ASB_TRY(rc);
    if (stats_d) {
        lambda_stats_kernel<<<1, 1024, 0, ctx->stream>>>(lambdas_d, (long long)n, stats_d);
        ASB_TRY(asb_check_launch(ctx, "lambda_stats_kernel"));
    }
    return ASB_OK;
}

int asb_dev_norms2(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, double *norms2_d) {
    const int wpb = 8;
    const long long grid = (n + wpb - 1) / wpb;
    norms2_kernel<<<(unsigned)grid, wpb * 32, 0, ctx->stream>>>(rows_d, (long long)n, (int)f, norms2_d);
    return asb_check_launch(ctx, "norms2_kernel");
}
