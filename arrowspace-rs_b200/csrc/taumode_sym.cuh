// taumode_sym.cuh -- K5+K6 for SYMMETRIC feature graphs (every graph the Laplacian stage emits).
//
// Same contract and tile layout as taumode_kernel (taumode.cu) but with the arithmetic
// restructured around the profile of the first version (profiles/r01_v1_taumode_summary.csv:
// instruction-issue bound, 6 k warp-instructions per item):
//   * each undirected edge {i,j} is visited once.  For a symmetric L with off-diagonals -w_ij
//       x^T L x = sum_i (L_ii - sum_j w_ij) x_i^2 + sum_{i<j} w_ij (x_i - x_j)^2
//     (exact algebra; the residual r_i = L_ii - sum_j w_ij is computed on the host in the
//     reference's own summation order and is exactly 0 for a Laplacian), and the dispersion sums
//     of src/taumode.rs:577-583,621-631 count every edge twice: edge = 2 S1, G = S2 / (2 S1^2)
//     with S1 = sum_{w>0} w d^2, S2 = sum_{w>0} (w d^2)^2.  No cancellation, so the result agrees
//     with the reference to its own rounding error (<< 1e-9 relative);
//   * the edge list is split EVENLY over the row partitions and walked 4 edges at a time with
//     all loads issued first (the first version was latency bound on one dependent load chain);
//   * tau: counting-only pivot passes (one compare + one add per element) narrow a value bracket
//     until at most 8 finite values remain inside, which one extraction pass collects and sorts.
#pragma once

namespace {

struct __align__(16) SymEdge {
    double w;  // -L_ij (> 0 for a Laplacian)
    int i, j;  // i < j
};

constexpr int kSelCap = 8;

template <int TI, bool ALLPOS>
__global__ void __launch_bounds__(kThreads)
taumode_sym_kernel(const double *__restrict__ items, long long n, int f, const SymEdge *__restrict__ edges,
                   int nedges, const double *__restrict__ resid, int tau_mode, double tau_value,
                   double *__restrict__ lambdas, double *__restrict__ norms2, int *__restrict__ nonfinite_flag) {
    constexpr int PITCH = TI + 1;
    constexpr int G = 32 / TI;
    constexpr int P = kWarps * G;
    extern __shared__ double smem[];
    double *X = smem;                        // f * PITCH
    double *red = smem + (size_t)f * PITCH;  // 4 * P * TI
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = lane % TI;
    const int part = warp * G + lane / TI;
    auto R = [&](int slot, int p, int it) -> double & { return red[(slot * P + p) * TI + it]; };
    // selection scratch lives in slots 1..3 of `red` (slot 0 is used by the reductions of that phase)
    double *sel_list = red + (size_t)1 * P * TI;                         // TI * kSelCap doubles
    int *sel_cnt = reinterpret_cast<int *>(red + (size_t)3 * P * TI);    // TI ints

    const int e_begin = (int)(((long long)nedges * part) / P);
    const int e_end = (int)(((long long)nedges * (part + 1)) / P);

    const long long ntiles = (n + TI - 1) / TI;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TI;
        __syncthreads();
        for (int it = warp; it < TI; it += kWarps) {
            const long long row = base + it;
            const bool valid = row < n;
            const double *src = items + row * (long long)f;
#pragma unroll 12
            for (int j = lane; j < f; j += 32) X[j * PITCH + it] = valid ? __ldg(src + j) : 0.0;
        }
        __syncthreads();

        // ---- pass 0: norms, value statistics, residual diagonal term
        double den = 0.0, num = 0.0, s1 = 0.0, s2 = 0.0, vsum = 0.0;
        double vmin = INFINITY, vmax = -INFINITY;
        int cnt = 0, nneg = 0;  // finite values / values equal to -inf
        for (int i = part; i < f; i += P) {
            const double xi = X[i * PITCH + item];
            const double x2 = xi * xi;
            den += x2;
            num = fma(__ldg(resid + i), x2, num);
            if (fabs(xi) < INFINITY) {
                cnt++;
                vsum += xi;
                vmin = fmin(vmin, xi);
                vmax = fmax(vmax, xi);
            }
            nneg += (xi == -INFINITY) ? 1 : 0;
        }
        // ---- edges of this partition, 4 at a time
        {
            const double *Xi = X + item;
            int e = e_begin;
            for (; e + 4 <= e_end; e += 4) {
                SymEdge ed[4];
                double a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) ed[u] = edges[e + u];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[u] = Xi[ed[u].i * PITCH];
                    b[u] = Xi[ed[u].j * PITCH];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double d = a[u] - b[u];
                    const double c = ed[u].w * (d * d);
                    if (ALLPOS) {
                        s1 += c;
                        s2 = fma(c, c, s2);
                    } else {
                        num += c;
                        if (ed[u].w > 0.0) {
                            s1 += c;
                            s2 = fma(c, c, s2);
                        }
                    }
                }
            }
            for (; e < e_end; ++e) {
                const SymEdge ed = edges[e];
                const double d = Xi[ed.i * PITCH] - Xi[ed.j * PITCH];
                const double c = ed.w * (d * d);
                if (ALLPOS) {
                    s1 += c;
                    s2 = fma(c, c, s2);
                } else {
                    num += c;
                    if (ed.w > 0.0) {
                        s1 += c;
                        s2 = fma(c, c, s2);
                    }
                }
            }
        }
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
        if (ALLPOS) num += s1;  // x^T L x = sum r_i x_i^2 + sum_edges w d^2
        __syncthreads();
        R(0, part, item) = vsum;
        R(1, part, item) = vmin;
        R(2, part, item) = vmax;
        R(3, part, item) = (double)(cnt + 65536 * nneg);  // both counts fit (f <= 5500)
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
        const int nf = ((int)cntd) & 65535;
        const int nneg_all = ((int)cntd) >> 16;
        __syncthreads();

        // ---- tau (src/taumode.rs:87-127)
        double tau;
        bool need_sel = false, need_pair = false;
        int rb = 0;
        if (tau_mode == ASB_TAU_FIXED) {
            tau = (fabs(tau_value) < INFINITY && tau_value > 0.0) ? tau_value : kTauFloor;
        } else if (tau_mode == ASB_TAU_MEAN) {
            tau = fmax(nf > 0 ? vsum / (double)nf : 0.0, kTauFloor);
        } else {
            tau = kTauFloor;
            if (nf > 0) {
                need_sel = true;
                if (tau_mode == ASB_TAU_PERCENTILE) {
                    double pp = tau_value;
                    pp = pp < 0.0 ? 0.0 : (pp > 1.0 ? 1.0 : pp);
                    const double fi = round((double)(nf - 1) * pp);
                    rb = (fi != fi || fi < 0.0) ? 0 : (int)fi;
                    if (rb > nf - 1) rb = nf - 1;
                } else {
                    rb = nf / 2;
                    need_pair = (nf % 2 == 0);
                }
            }
        }
        const bool block_sel = __syncthreads_or(need_sel);
        if (block_sel) {
            // bracket [lo, hi): c_lo = #(x < lo) <= rb < c_hi = #(x < hi); hi = +inf means "all finite"
            double lo = vmin, hi = INFINITY;
            int c_lo = 0, c_hi = nf;
            bool narrowing = need_sel;
            for (int iter = 0;; ++iter) {
                if (narrowing) {
                    const bool dup = (lo == vmax) || (hi < INFINITY && ord_key(hi) == ord_key(lo) + 1ull);
                    if (c_hi - c_lo <= kSelCap || dup || iter > 200) narrowing = false;
                }
                if (!__syncthreads_or(narrowing)) break;
                double pv = 0.0;
                if (narrowing) {
                    const double hi_f = (hi == INFINITY) ? vmax : hi;
                    const double frac = ((double)(rb - c_lo) + 0.5) / (double)(c_hi - c_lo);
                    pv = lo + (hi_f - lo) * frac;
                    if (iter >= 6 || !(fabs(pv) < INFINITY)) {
                        const unsigned long long kl = ord_key(lo), kh = ord_key(hi_f);
                        pv = ord_val(kl + ((kh - kl) >> 1) + ((kh - kl) & 1ull));
                    }
                    if (hi < INFINITY) {  // keep lo < pv < hi (room exists: adjacent keys end the search above)
                        if (!(pv < hi)) pv = ord_val(ord_key(hi) - 1ull);
                    } else if (pv > vmax) {
                        pv = vmax;
                    }
                    if (!(pv > lo)) pv = ord_val(ord_key(lo) + 1ull);
                }
                int c = 0;
                if (narrowing) {  // #(x < pv): NaN and +inf compare false, the -inf entries are subtracted below
                    const double *px = X + part * PITCH + item;
                    constexpr int STR = P * PITCH;
                    const int nrows = (f - part + P - 1) / P;
                    int i = 0;
                    for (; i + 4 <= nrows; i += 4) {
                        const double x0 = px[0], x1 = px[STR], x2 = px[2 * STR], x3 = px[3 * STR];
                        c += (x0 < pv ? 1 : 0) + (x1 < pv ? 1 : 0) + (x2 < pv ? 1 : 0) + (x3 < pv ? 1 : 0);
                        px += 4 * STR;
                    }
                    for (; i < nrows; ++i) {
                        c += (px[0] < pv ? 1 : 0);
                        px += STR;
                    }
                }
                R(0, part, item) = (double)c;
                __syncthreads();
                if (narrowing) {
                    double cs = 0.0;
#pragma unroll
                    for (int p = 0; p < P; ++p) cs += R(0, p, item);
                    const int c_p = (int)cs - nneg_all;
                    if (c_p <= rb) {
                        lo = pv;
                        c_lo = c_p;
                    } else {
                        hi = pv;
                        c_hi = c_p;
                    }
                }
                __syncthreads();
            }
            // ---- extraction: values in [lo, hi) (when few) and the largest value below lo
            const bool small = need_sel && (c_hi - c_lo <= kSelCap);
            if (part == 0) sel_cnt[item] = 0;
            __syncthreads();
            double below = -INFINITY;
            if (need_sel) {
                for (int i = part; i < f; i += P) {
                    const double x = X[i * PITCH + item];
                    if (fabs(x) < INFINITY) {
                        if (x < lo) {
                            below = fmax(below, x);
                        } else if (small && x < hi) {
                            const int slot = atomicAdd(&sel_cnt[item], 1);
                            if (slot < kSelCap) sel_list[item * kSelCap + slot] = x;
                        }
                    }
                }
            }
            R(0, part, item) = below;
            __syncthreads();
            if (need_sel && part == 0) {
                below = -INFINITY;
#pragma unroll
                for (int p = 0; p < P; ++p) below = fmax(below, R(0, p, item));
                double v_rb, v_ra;
                if (small) {
                    double v[kSelCap];
                    const int m = c_hi - c_lo;
#pragma unroll
                    for (int q = 0; q < kSelCap; ++q) v[q] = q < m ? sel_list[item * kSelCap + q] : INFINITY;
#pragma unroll
                    for (int a = 0; a < kSelCap; ++a)  // odd-even transposition sort, 8 elements
#pragma unroll
                        for (int q = (a & 1); q + 1 < kSelCap; q += 2) {
                            const double x0 = fmin(v[q], v[q + 1]), x1 = fmax(v[q], v[q + 1]);
                            v[q] = x0;
                            v[q + 1] = x1;
                        }
                    const int k = rb - c_lo;
                    v_rb = v[0];
                    v_ra = below;
#pragma unroll
                    for (int q = 0; q < kSelCap; ++q) {
                        if (q == k) v_rb = v[q];
                        if (q == k - 1) v_ra = v[q];
                    }
                } else {  // every finite value inside the bracket equals lo
                    v_rb = lo;
                    v_ra = (rb - 1 >= c_lo) ? lo : below;
                }
                const double m = need_pair ? 0.5 * (v_ra + v_rb) : v_rb;
                tau = fmax(m, kTauFloor);
            }
            __syncthreads();
        }

        // ---- lambda (src/taumode.rs:596-647); edge_energy = 2 S1, G = sum (c / edge)^2 = S2 / (2 S1^2)
        if (part == 0 && base + item < n) {
            const double e_raw = den > 1e-12 ? num / den : 0.0;
            double g = 0.0;
            if (s1 > 0.0) {
                g = s2 / (2.0 * s1 * s1);
                g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
            }
            const double e_bounded = e_raw / (e_raw + tau);
            lambdas[base + item] = tau * e_bounded + (1.0 - tau) * g;
            if (norms2) norms2[base + item] = den;
            if (nonfinite_flag && nf < f) atomicOr(nonfinite_flag, 1);
        }
    }
}

}  // namespace
