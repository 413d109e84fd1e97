// taumode_sym.cuh -- K5+K6 for SYMMETRIC feature graphs (every graph the Laplacian stage emits):
// one WARP per item, no block-level synchronisation.
//
// Replaces TauMode::select_tau (src/taumode.rs:87-127) + compute_synthetic_lambda_csr
// (src/taumode.rs:552-660) per item, as driven by compute_taumode_lambdas_parallel (:174-312).
//
// History (profiles/): v1 (CTA tile of 32 items, CSR walk) was instruction-issue bound at 6 k
// warp-instructions per item; v2 (edge-once arithmetic, counting selection, same tile) 3.7 k and still
// dozens of CTA barriers per tile.  v3 gives every item to one warp:
//   * the warp streams its item from HBM with coalesced 256 B loads into a private shared-memory copy
//     (lane l owns features l, l+32, ...), so the value statistics and the tau selection run on the
//     lane's own elements with warp-wide integer reductions (__reduce_add_sync) -- no barriers at all;
//   * each undirected edge {i,j} is visited once, 32 edges per step (one per lane).  For a symmetric L
//       x^T L x = sum_i (L_ii - sum_j w_ij) x_i^2 + sum_{i<j} w_ij (x_i - x_j)^2          (exact algebra;
//     the residual of the STORED row is accumulated error-free on the host: ~1e-16, not 0, for a Laplacian), and the
//     dispersion sums of :577-583,:621-631 count every edge twice: edge = 2 S1, G = S2 / (2 S1^2);
//   * the random x[i], x[j] reads would bank-conflict ~3-way, so the host packs the edge list into steps
//     whose 16-lane halves touch 16 distinct bank pairs (greedy colouring; the graph is fixed for millions
//     of items) -- the schedule is padded with zero-weight edges;
//   * tau (median / percentile): counting-only pivot passes narrow a value bracket until <= 8 finite
//     values remain, which are gathered with ballots and sorted in registers.
// HBM traffic is the algorithmic 8 F + 16 bytes per item; the schedule (<= 30 kB) lives in L1/L2.
#pragma once

namespace {

struct __align__(16) SymEdge {
    double w;  // -L_ij (> 0 for a Laplacian); 0 for padding
    int i, j;  // i < j
};

__device__ __forceinline__ SymEdge ld_edge(const SymEdge *p) {  // one 16-byte read-only load
    const int4 r = __ldg(reinterpret_cast<const int4 *>(p));
    SymEdge e;
    e.w = __hiloint2double(r.y, r.x);
    e.i = r.z;
    e.j = r.w;
    return e;
}

constexpr int kSelCap = 8;
constexpr int kTauWarps = 8;

template <bool ALLPOS>
__global__ void __launch_bounds__(kTauWarps * 32)
taumode_warp_kernel(const double *__restrict__ items, long long n, int f, const SymEdge *__restrict__ sched,
                    int nsteps, const double *__restrict__ resid, int tau_mode, double tau_value,
                    double *__restrict__ lambdas, double *__restrict__ norms2, int *__restrict__ nonfinite_flag,
                    int warps_per_cta, int rg) {
    // rg = nodes of the graph, <= f: after a JL-projected build the graph is r x r while the items keep their F
    // values -- the reference then reads item[0 .. r) in the Rayleigh sums and ALL F values for tau and the
    // denominator (src/taumode.rs:233-245,596 with src/eigenmaps.rs:248-269)
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_cta) return;  // no block-level barrier anywhere below
    double *xs = smem + (size_t)warp * f;
    const long long wstride = (long long)gridDim.x * warps_per_cta;
    for (long long item = (long long)blockIdx.x * warps_per_cta + warp; item < n; item += wstride) {
        const double *src = items + item * (long long)f;
        // ---- load + per-lane statistics
        double den = 0.0, num = 0.0, vsum = 0.0, vmin = INFINITY, vmax = -INFINITY;
        int cnt = 0, nneg = 0;
        __syncwarp();  // the previous item's readers are done with xs
#pragma unroll 4
        for (int j = lane; j < f; j += 32) {
            const double x = __ldg(src + j);
            xs[j] = x;
            const double x2 = x * x;
            den += x2;
            if (j < rg) num = fma(__ldg(resid + j), x2, num);
            if (fabs(x) < INFINITY) {
                cnt++;
                vsum += x;
                vmin = fmin(vmin, x);
                vmax = fmax(vmax, x);
            }
            nneg += (x == -INFINITY) ? 1 : 0;
        }
        __syncwarp();
        // ---- edges: one conflict-free step = 32 edges
        double s1 = 0.0, s2 = 0.0;
        {
            const SymEdge *sp = sched + lane;
            int s = 0;
            for (; s + 2 <= nsteps; s += 2) {
                const SymEdge e0 = ld_edge(sp + (size_t)s * 32), e1 = ld_edge(sp + (size_t)(s + 1) * 32);
                const double a0 = xs[e0.i], b0 = xs[e0.j], a1 = xs[e1.i], b1 = xs[e1.j];
                const double d0 = a0 - b0, d1 = a1 - b1;
                const double c0 = e0.w * (d0 * d0), c1 = e1.w * (d1 * d1);
                if (ALLPOS) {
                    s1 += c0;
                    s2 = fma(c0, c0, s2);
                    s1 += c1;
                    s2 = fma(c1, c1, s2);
                } else {
                    num += c0 + c1;
                    if (e0.w > 0.0) {
                        s1 += c0;
                        s2 = fma(c0, c0, s2);
                    }
                    if (e1.w > 0.0) {
                        s1 += c1;
                        s2 = fma(c1, c1, s2);
                    }
                }
            }
            for (; s < nsteps; ++s) {
                const SymEdge e0 = ld_edge(sp + (size_t)s * 32);
                const double d0 = xs[e0.i] - xs[e0.j];
                const double c0 = e0.w * (d0 * d0);
                if (ALLPOS) {
                    s1 += c0;
                    s2 = fma(c0, c0, s2);
                } else {
                    num += c0;
                    if (e0.w > 0.0) {
                        s1 += c0;
                        s2 = fma(c0, c0, s2);
                    }
                }
            }
        }
        // ---- warp reductions
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            den += __shfl_xor_sync(0xffffffffu, den, o);
            num += __shfl_xor_sync(0xffffffffu, num, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (ALLPOS) num += s1;  // x^T L x = sum r_i x_i^2 + sum_edges w d^2
        const int nf = __reduce_add_sync(0xffffffffu, cnt);
        const int nneg_all = __reduce_add_sync(0xffffffffu, nneg);

        // ---- tau (src/taumode.rs:87-127)
        double tau;
        if (tau_mode == ASB_TAU_FIXED) {
            tau = (fabs(tau_value) < INFINITY && tau_value > 0.0) ? tau_value : kTauFloor;
        } else if (tau_mode == ASB_TAU_MEAN) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
            tau = fmax(nf > 0 ? vsum / (double)nf : 0.0, kTauFloor);
        } else if (nf == 0) {
            tau = kTauFloor;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
                vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            }
            int rb;
            bool need_pair = false;
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
            // bracket [lo, hi): c_lo = #(x < lo) <= rb < c_hi = #(x < hi); hi = +inf means "all finite values"
            double lo = vmin, hi = INFINITY;
            int c_lo = 0, c_hi = nf;
            for (int iter = 0; iter < 200; ++iter) {  // every quantity below is warp-uniform
                const bool dup = (lo == vmax) || (hi < INFINITY && ord_key(hi) == ord_key(lo) + 1ull);
                if (c_hi - c_lo <= kSelCap || dup) break;
                const double hi_f = (hi == INFINITY) ? vmax : hi;
                const double frac = ((double)(rb - c_lo) + 0.5) / (double)(c_hi - c_lo);
                double pv = lo + (hi_f - lo) * frac;
                if (iter >= 6 || !(fabs(pv) < INFINITY)) {  // bisection in ordered-key space: <= 64 more passes
                    const unsigned long long kl = ord_key(lo), kh = ord_key(hi_f);
                    pv = ord_val(kl + ((kh - kl) >> 1) + ((kh - kl) & 1ull));
                }
                if (hi < INFINITY) {
                    if (!(pv < hi)) pv = ord_val(ord_key(hi) - 1ull);
                } else if (pv > vmax) {
                    pv = vmax;
                }
                if (!(pv > lo)) pv = ord_val(ord_key(lo) + 1ull);
                int c = 0;
#pragma unroll 4
                for (int j = lane; j < f; j += 32) c += (xs[j] < pv) ? 1 : 0;  // NaN / +inf compare false
                const int c_p = __reduce_add_sync(0xffffffffu, c) - nneg_all;    // -inf entries are not finite
                if (c_p <= rb) {
                    lo = pv;
                    c_lo = c_p;
                } else {
                    hi = pv;
                    c_hi = c_p;
                }
            }
            // ---- extraction: the <= 8 finite values inside [lo, hi) are gathered one per lane (ballot order),
            //      ranked with 8 shuffles; also the largest finite value below lo
            const bool small = (c_hi - c_lo <= kSelCap);
            double mine = INFINITY;  // lane q < nv holds the q-th gathered value
            int nv = 0;
            double below = -INFINITY;
            for (int j0 = 0; j0 < f; j0 += 32) {
                const int j = j0 + lane;
                const double x = j < f ? xs[j] : __longlong_as_double(0x7ff8000000000000ll);
                const bool fin = fabs(x) < INFINITY;
                if (fin && x < lo) below = fmax(below, x);
                const unsigned m = __ballot_sync(0xffffffffu, small && fin && x >= lo && x < hi);
                if (m) {  // warp-uniform
                    const int want = lane - nv;  // lanes nv .. nv+popc(m)-1 fetch the (want+1)-th set bit's value
                    const int cntm = __popc(m);
                    const int srcl = (want >= 0 && want < cntm) ? (int)__fns(m, 0, want + 1) : lane;
                    const double xv = __shfl_sync(0xffffffffu, x, srcl);
                    if (want >= 0 && want < cntm) mine = xv;
                    nv += cntm;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) below = fmax(below, __shfl_xor_sync(0xffffffffu, below, o));
            double v_rb, v_ra;
            if (small) {
                // rank of my value among the gathered ones (ties broken by lane) -> fetch ranks k and k-1
                int rank = 0;
#pragma unroll
                for (int q = 0; q < kSelCap; ++q) {
                    const double o = __shfl_sync(0xffffffffu, mine, q);
                    rank += (q < nv && (o < mine || (o == mine && q < lane))) ? 1 : 0;
                }
                const int k = rb - c_lo;
                const unsigned has_k = __ballot_sync(0xffffffffu, lane < nv && rank == k);
                const unsigned has_km1 = __ballot_sync(0xffffffffu, lane < nv && rank == k - 1);
                v_rb = __shfl_sync(0xffffffffu, mine, has_k ? __ffs(has_k) - 1 : 0);
                v_ra = has_km1 ? __shfl_sync(0xffffffffu, mine, __ffs(has_km1) - 1) : below;
            } else {  // every finite value inside the bracket equals lo
                v_rb = lo;
                v_ra = (rb - 1 >= c_lo) ? lo : below;
            }
            tau = fmax(need_pair ? 0.5 * (v_ra + v_rb) : v_rb, kTauFloor);  // :119-124
        }

        // ---- lambda (src/taumode.rs:596-647); edge_energy = 2 S1, G = sum (c / edge)^2 = S2 / (2 S1^2)
        if (lane == 0) {
            const double e_raw = den > 1e-12 ? num / den : 0.0;
            double g = 0.0;
            if (s1 > 0.0) {
                g = s2 / (2.0 * s1 * s1);
                g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
            }
            const double e_bounded = e_raw / (e_raw + tau);
            lambdas[item] = tau * e_bounded + (1.0 - tau) * g;
            if (norms2) norms2[item] = den;
            if (nonfinite_flag && nf < f) atomicOr(nonfinite_flag, 1);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// v4: the item lives in REGISTERS as well (lane l holds features l, l+32, ...: NPL of them, compile time), two
// items per warp pass.  ncu of v3 (profiles/r02_taumode_v3_*): the kernel is bound by the L1 / shared-memory data path
// (570 wavefront cycles per item and SM: 188 for the two x reads per edge, 188 for the 16-byte schedule entry per
// edge, 144 for the selection's re-reads of the item, 48 for the copy-in), not by HBM or the FP64 pipe.  v4 keeps only
// the random x[i], x[j] reads in shared memory: the statistics, the counting passes of the tau selection and the
// extraction run on the lane's registers, and one schedule entry serves IPP items (330 cycles per item at IPP = 2).
// Lanes' missing elements (f < 32 NPL) are NaN: every comparison is false, every statistic is predicated on j < f.
template <int NPL>
struct TauItem {
    double xv[NPL];
    double den, num, vsum, vmin, vmax;
    int cnt, nneg;
};

template <int NPL>
__device__ __forceinline__ double tau_select_regs(const double (&xv)[NPL], int lane, int tau_mode, double tau_value, int nf,
                                                  int nneg_all, double vmin, double vmax, double *scratch) {
    // scratch: kSelCap doubles of this warp's shared memory (the gathered bracket)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    int rb;
    bool need_pair = false;
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
    double lo = vmin, hi = INFINITY;
    int c_lo = 0, c_hi = nf;
    for (int iter = 0; iter < 200; ++iter) {  // every quantity below is warp-uniform
        const bool dup = (lo == vmax) || (hi < INFINITY && ord_key(hi) == ord_key(lo) + 1ull);
        if (c_hi - c_lo <= kSelCap || dup) break;
        const double hi_f = (hi == INFINITY) ? vmax : hi;
        // (any value inside (lo, hi) is a valid pivot -- only the exact counts decide -- so the interpolation weight
        // may be a fast FP32 quotient)
        const float fracf = __fdividef((float)(rb - c_lo) + 0.5f, (float)(c_hi - c_lo));
        double pv = fma(hi_f - lo, (double)fracf, lo);
        if (iter >= 6 || !(fabs(pv) < INFINITY)) {
            const unsigned long long kl = ord_key(lo), kh = ord_key(hi_f);
            pv = ord_val(kl + ((kh - kl) >> 1) + ((kh - kl) & 1ull));
        }
        if (hi < INFINITY) {
            if (!(pv < hi)) pv = ord_val(ord_key(hi) - 1ull);
        } else if (pv > vmax) {
            pv = vmax;
        }
        if (!(pv > lo)) pv = ord_val(ord_key(lo) + 1ull);
        int c = 0;
#pragma unroll
        for (int t = 0; t < NPL; ++t) c += (xv[t] < pv) ? 1 : 0;  // NaN / +inf compare false
        const int c_p = __reduce_add_sync(0xffffffffu, c) - nneg_all;
        if (c_p <= rb) {
            lo = pv;
            c_lo = c_p;
        } else {
            hi = pv;
            c_hi = c_p;
        }
    }
    const bool small = (c_hi - c_lo <= kSelCap);
    // ---- extraction: the <= kSelCap finite values inside [lo, hi) go to the warp's scratch (lane-major order: any
    //      order will do, they are ranked below); also the largest finite value below lo
    int mycnt = 0;
    double below = -INFINITY;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const double x = xv[t];
        mycnt += (x >= lo && x < hi) ? 1 : 0;                      // NaN compares false; +-inf fall outside
        if (x < lo && x > -INFINITY) below = x > below ? x : below;
    }
    int pos = mycnt;   // inclusive scan over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pos, o);
        if (lane >= o) pos += v;
    }
    const int nv = small ? __shfl_sync(0xffffffffu, pos, 31) : 0;
    pos -= mycnt;
    __syncwarp();
    if (small && mycnt > 0) {
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const double x = xv[t];
            if (x >= lo && x < hi) scratch[pos++] = x;
        }
    }
    __syncwarp();
    const double mine = (small && lane < nv) ? scratch[lane] : INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, below, o);
        below = ob > below ? ob : below;
    }
    double v_rb, v_ra;
    if (small) {
        int rank = 0;
#pragma unroll
        for (int q = 0; q < kSelCap; ++q) {
            const double o = __shfl_sync(0xffffffffu, mine, q);
            rank += (q < nv && (o < mine || (o == mine && q < lane))) ? 1 : 0;
        }
        const int k = rb - c_lo;
        const unsigned has_k = __ballot_sync(0xffffffffu, lane < nv && rank == k);
        const unsigned has_km1 = __ballot_sync(0xffffffffu, lane < nv && rank == k - 1);
        v_rb = __shfl_sync(0xffffffffu, mine, has_k ? __ffs(has_k) - 1 : 0);
        v_ra = has_km1 ? __shfl_sync(0xffffffffu, mine, __ffs(has_km1) - 1) : below;
    } else {
        v_rb = lo;
        v_ra = (rb - 1 >= c_lo) ? lo : below;
    }
    return fmax(need_pair ? 0.5 * (v_ra + v_rb) : v_rb, kTauFloor);
}

template <bool ALLPOS, int NPL, int IPP>
__global__ void __launch_bounds__(kTauWarps * 32)
taumode_reg_kernel(const double *__restrict__ items, long long n, int f, const SymEdge *__restrict__ sched, int nsteps,
                   const double *__restrict__ resid, int tau_mode, double tau_value, double *__restrict__ lambdas,
                   double *__restrict__ norms2, int *__restrict__ nonfinite_flag, int rg) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = smem + (size_t)warp * IPP * f;
    double *scratch = smem + (size_t)kTauWarps * IPP * f + warp * kSelCap;
    const long long npass = (n + IPP - 1) / IPP;
    const long long wstride = (long long)gridDim.x * kTauWarps;
    for (long long pass = (long long)blockIdx.x * kTauWarps + warp; pass < npass; pass += wstride) {
        TauItem<NPL> it[IPP];
        long long idx[IPP];
        __syncwarp();  // the previous pass's readers are done with xs
        // ---- all loads of the pass first (IPP * NPL independent 256-byte requests per warp)
#pragma unroll
        for (int p = 0; p < IPP; ++p) {
            const long long item = pass * IPP + p;
            idx[p] = item < n ? item : -1;
            const double *src = items + (item < n ? item : pass * IPP) * (long long)f;  // a missing item repeats item 0
#pragma unroll
            for (int t = 0; t < NPL; ++t) {
                const int j = 32 * t + lane;
                it[p].xv[t] = j < f ? __ldg(src + j) : __longlong_as_double(0x7ff8000000000000ll);
            }
        }
#pragma unroll
        for (int p = 0; p < IPP; ++p) {
            it[p].den = 0.0, it[p].num = 0.0, it[p].vsum = 0.0, it[p].vmin = INFINITY, it[p].vmax = -INFINITY;
            it[p].cnt = 0, it[p].nneg = 0;
        }
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int j = 32 * t + lane;
            if (j < f) {
                const double r = j < rg ? __ldg(resid + j) : 0.0;
#pragma unroll
                for (int p = 0; p < IPP; ++p) {
                    const double x = it[p].xv[t];
                    xs[p * f + j] = x;
                    const double x2 = x * x;
                    it[p].den += x2;
                    if (j < rg) it[p].num = fma(r, x2, it[p].num);
                    if (fabs(x) < INFINITY) {
                        it[p].cnt++;
                        it[p].vsum += x;
                        it[p].vmin = x < it[p].vmin ? x : it[p].vmin;   // (x is finite: no NaN semantics needed)
                        it[p].vmax = x > it[p].vmax ? x : it[p].vmax;
                    }
                    it[p].nneg += (x == -INFINITY) ? 1 : 0;
                }
            }
        }
        __syncwarp();
        // ---- edges: one schedule entry per lane and step, applied to every item of the pass
        double s1[IPP], s2[IPP];
#pragma unroll
        for (int p = 0; p < IPP; ++p) s1[p] = 0.0, s2[p] = 0.0;
        {
            const SymEdge *sp = sched + lane;
#pragma unroll 4
            for (int s = 0; s < nsteps; ++s) {
                const SymEdge e = ld_edge(sp + (size_t)s * 32);
#pragma unroll
                for (int p = 0; p < IPP; ++p) {
                    const double d = xs[p * f + e.i] - xs[p * f + e.j];
                    const double c = e.w * (d * d);
                    if (ALLPOS) {
                        s1[p] += c;
                        s2[p] = fma(c, c, s2[p]);
                    } else {
                        it[p].num += c;
                        if (e.w > 0.0) {
                            s1[p] += c;
                            s2[p] = fma(c, c, s2[p]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < IPP; ++p) {
            double den = it[p].den, num = it[p].num, a1 = s1[p], a2 = s2[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                den += __shfl_xor_sync(0xffffffffu, den, o);
                num += __shfl_xor_sync(0xffffffffu, num, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (ALLPOS) num += a1;
            const int nf = __reduce_add_sync(0xffffffffu, it[p].cnt);
            const int nneg_all = __reduce_add_sync(0xffffffffu, it[p].nneg);
            double tau;
            if (tau_mode == ASB_TAU_FIXED) {
                tau = (fabs(tau_value) < INFINITY && tau_value > 0.0) ? tau_value : kTauFloor;
            } else if (tau_mode == ASB_TAU_MEAN) {
                double vsum = it[p].vsum;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
                tau = fmax(nf > 0 ? vsum / (double)nf : 0.0, kTauFloor);
            } else if (nf == 0) {
                tau = kTauFloor;
            } else {
                tau = tau_select_regs<NPL>(it[p].xv, lane, tau_mode, tau_value, nf, nneg_all, it[p].vmin, it[p].vmax, scratch);
            }
            if (lane == 0 && idx[p] >= 0) {
                const double e_raw = den > 1e-12 ? num / den : 0.0;
                double g = 0.0;
                if (a1 > 0.0) {
                    g = a2 / (2.0 * a1 * a1);
                    g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
                }
                const double e_bounded = e_raw / (e_raw + tau);
                lambdas[idx[p]] = tau * e_bounded + (1.0 - tau) * g;
                if (norms2) norms2[idx[p]] = den;
                if (nonfinite_flag && nf < f) atomicOr(nonfinite_flag, 1);
            }
        }
    }
}

}  // namespace

// Host side: pack the undirected edges into steps of 32 such that, inside each half-warp (16 lanes), the
// 8-byte shared-memory words x[i] hit 16 distinct bank pairs (i mod 16) and so do the x[j]; equal indices
// are allowed (broadcast).  Greedy first-fit; holes are padded with zero-weight edges.
static void asb_schedule_edges(const std::vector<SymEdge> &edges, std::vector<SymEdge> &sched) {
    sched.clear();
    std::vector<char> used(edges.size(), 0);
    size_t remaining = edges.size(), first_free = 0;
    const SymEdge pad = {0.0, 0, 0};
    while (remaining > 0) {
        SymEdge step[32];
        for (int half = 0; half < 2; ++half) {
            int bank_i[16], bank_j[16];  // index stored in each bank pair, -1 = free
            for (int b = 0; b < 16; ++b) bank_i[b] = bank_j[b] = -1;
            int filled = 0;
            for (size_t e = first_free; e < edges.size() && filled < 16; ++e) {
                if (used[e]) continue;
                const int bi = edges[e].i & 15, bj = edges[e].j & 15;
                const bool ok_i = bank_i[bi] < 0 || bank_i[bi] == edges[e].i;
                const bool ok_j = bank_j[bj] < 0 || bank_j[bj] == edges[e].j;
                if (!(ok_i && ok_j)) continue;
                bank_i[bi] = edges[e].i;
                bank_j[bj] = edges[e].j;
                step[half * 16 + filled] = edges[e];
                used[e] = 1;
                filled++;
                remaining--;
            }
            for (; filled < 16; ++filled) step[half * 16 + filled] = pad;
            while (first_free < edges.size() && used[first_free]) first_free++;
        }
        sched.insert(sched.end(), step, step + 32);
    }
}
