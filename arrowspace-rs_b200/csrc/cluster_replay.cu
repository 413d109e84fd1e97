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

#include <algorithm>
#include <chrono>
#include <vector>

#include "comm.cuh"

namespace {

__global__ void __launch_bounds__(256) replay_keys_kernel(const int64_t *__restrict__ top2_idx, int stride, int m, int K,
                                                          int *__restrict__ keys, int *__restrict__ vals) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) {
        const int64_t b = top2_idx[(size_t)stride * r];
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

// order[b] = the centroid block b of the chain kernel walks: longest chain first, so that the critical path of the chunk
// starts at once and the short chains fill the SMs behind it (K may exceed the CTAs that fit the device at a time)
__global__ void __launch_bounds__(1024) replay_order_kernel(const int *__restrict__ seg_off, int K, int *__restrict__ order) {
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        const int len = seg_off[c + 1] - seg_off[c];
        int rank = 0;
        for (int o = 0; o < K; ++o) {
            const int lo = seg_off[o + 1] - seg_off[o];
            rank += (lo > len || (lo == len && o < c)) ? 1 : 0;
        }
        order[rank] = c;
    }
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
                                                           const double *__restrict__ cent0,
                                                           const double *__restrict__ disp0, unsigned long long *sizes,
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
    double dmax2 = disp0 ? disp0[c] * disp0[c] * (1.0 + 1e-12) : 0.0;
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

// ---- the chain kernel proper -----------------------------------------------------------------------------------
// A chain is ONE dependent sequence: the longest chain of a chunk is the chunk's critical path (17.8k steps per 1M rows
// on the bench data), so the cost of a step is all that counts, and a lone warp pays ~4 cycles per instruction.  The
// first register-resident version spent 500 warp-instructions per step (ncu: profiles/r02_chain_v2_*), most of them on
// the distance |x - c|^2 and the displacement |c - S0|^2 with their cross-lane reductions; the second one (features
// split over four warps, no distance in the common step) 150, but stalled on its own global loads (index -> address
// -> row, and scoreboard slots shared between the load generations in flight: profiles/r02_chain_v3_*).  This version
//   * splits the FEATURES of a chain over 4 consumer warps (the update c += (x - c) / k is element-wise): each lane owns
//     ceil(f / 128) elements, centroid and snapshot in registers;
//   * streams the rows through a shared-memory ring filled by a fifth, producer warp with 1-D bulk async copies
//     (cp.async.bulk = the TMA engine, SASS UBLKCP) that complete on per-slot mbarriers: a slot = kRG rows + their
//     metadata, kRingSlots slots in flight, "full" / "empty" barriers both ways -- a consumer step reads shared memory
//     only, its address arithmetic is immediate offsets;
//   * does NOT compute the distance in the common step.  The nearest-centroid pass already delivered the distance d0
//     of the row to the SNAPSHOT of its centroid (certified to +-e: dlo <= true <= dhi), and the chain knows a bound B
//     on how far its centroid has moved since, so  |x - c_now|  lies in  [dlo - B, dhi + B]: when (dhi + B)^2 is safely
//     below the radius the row is an update, when (dlo - B)^2 is safely above 1.5 radius it is dropped -- no reduction,
//     no communication between the warps (all four evaluate the same scalars, bit for bit).  Anything else -- and any
//     row whose bound dhi + B would not certify against the runner-up's distance under the displacement the previous
//     chunk saw (`disp_hint`) -- takes the exact step: |x - c|^2 reduced across lanes and warps (shared memory, one
//     named barrier), classified with the guard band;
//   * keeps B rigorous and cheap: every update moves the centroid by |x - c| / k <= (dhi + B) / k, which is added to B;
//     every `interval` steps (1 while the count is small, k / 256 up to 64 later: the bound may grow by about 0.4 %
//     of the row distance between two exact values) B is reset to the exact |c - S0| (one reduction + one barrier);
//   * divides branch-free: reciprocal + two FMA corrections for every element (Markstein; div_by_count of the
//     sequential kernel, correctly rounded for an integer divisor), an exponent-field test on the integer pipe sends the
//     step to the careful variant (library division for the affected elements) when an element is zero, denormal-small,
//     huge or not finite; the reciprocal of the next count is computed while the current update is in flight.
// The certification kernel receives an UPPER BOUND of the row's distance to its centroid at the row's own time
// (dhi + B, or the exact value from the exact step).
constexpr int kChainWarpsMax = 16;   // consumer warps CW (template); warps CW .. CW + kChainProducers - 1 are producers
constexpr int kChainProducers = 4;
constexpr int kRingSlots = 8;   // at most

struct __align__(16) SegMeta {    // per sorted position (replay_bounds_kernel)
    double dlo, dhi;              // certified bounds of the row's distance to the snapshot of its nearest centroid
    double slo;                   // certified lower bound of its distance to the runner-up's snapshot
    int row, pad;
};

struct ChainShared {
    double part[2][kChainWarpsMax];
    int stop[2];
    int drops;                    // rows the consumers classified "dropped" so far (the producer's count prediction)
    unsigned long long full[kRingSlots], empty[kRingSlots];
};

// per ring slot, written by the PRODUCER warp (it has nothing else to do between two bulk copies): everything of a
// group's bookkeeping that follows from the metadata and the predicted count alone
template <int RG>
struct __align__(16) GroupDesc {
    double yk[RG + 2];            // 1 / (kd_pred + 1 + t), t = 0 .. RG  (yk[RG]: the next group's first)
    double hmax;                  // max dhi of the group
    double minslack;              // min (slo - dhi)
    double kd_pred;               // the count the reciprocals assume
    double pad;
};

// ring depth: the chain kernel is bound by the round trip of its ring (consumer frees a slot -> producer wakes, issues the
// bulk copies -> rows arrive from DRAM -> consumer wakes: ~4 us measured -- 16 rows in flight gave 241 ns per row, 24 gave
// 180, 32 gave 130), so the ring takes what one CTA per SM can have: up to 192 kB of rows, at most kRingSlots slots;
// row128 = the ring pitch of a row in units of 128 features (1 kB)
__host__ __device__ constexpr int chain_ring_slots(int row128, int rg) {
    return 192 / (row128 * rg) < 2 ? 2 : (192 / (row128 * rg) > kRingSlots ? kRingSlots : (192 / (row128 * rg)) & ~1);   // even
}
__device__ __forceinline__ int chain_interval(double kd) {
    // steps between two exact values of the displacement bound: 1 while the count is small, then count / 128 (the
    // bound may grow by about 0.8 % of the row distance in between), at most 128
    return kd < 256.0 ? 1 : (kd < 16384.0 ? (int)(kd * (1.0 / 128.0)) : 128);
}

// FP64 operations the compiler may not reorder among themselves: the chain kernel issues the element chains of a row
// LAYER BY LAYER (sub for every element, then mul for every element, ...) so that a warp always has NPW independent
// instructions between two dependent ones (8 cycles of FP64 latency: profiles/r02_dp_latency.json).  Left to its own
// scheduler ptxas emitted the rows of a group element after element -- one fully dependent chain of 7 operations after
// the other, ~150 cycles per row instead of ~60 (profiles/r02_chain6_hotspots.txt).
__device__ __forceinline__ double rp_vsub(double a, double b) {
    double r;
    asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double rp_vadd(double a, double b) {
    double r;
    asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double rp_vmul(double a, double b) {
    double r;
    asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double rp_vfma(double a, double b, double c) {
    double r;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
    return r;
}

__device__ __forceinline__ unsigned rp_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rp_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rp_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rp_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rp_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rp_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rp_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "RP_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RP_DONE;\n\t"
        "bra RP_WAIT_LOOP;\n\t"
        "RP_DONE:\n\t}" ::"r"(rp_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void rp_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rp_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(rp_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void rp_prefetch_l2(const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// shared memory: ring[kRingSlots][RG][fpad] doubles, then meta[kRingSlots][RG]; fpad = 128 NPW
// CW consumer warps, NPW elements per lane: element u of a lane = feature 32 CW u + 32 warp + lane.  ONE element per lane
// wherever the row allows it (f <= 512): the update of an element is a chain of 7 dependent FP64 operations (8 cycles
// each, profiles/r02_dp_latency.json), and ptxas emits the chains of a lane's elements one AFTER the other rather than
// interleaved (profiles/r02_chain6_hotspots.txt: ~150 cycles per row at 3 elements per lane) -- with one element per lane
// the hardware interleaves the warps instead and a row costs its own latency.
template <int CW, int NPW, int RG>
__global__ void __launch_bounds__((CW + kChainProducers) * 32, 1) replay_chain_tma_kernel(
    const double *__restrict__ rows, int f, const int *__restrict__ seg_off, const int *__restrict__ order,
    const SegMeta *__restrict__ seg_meta, int K,
    int saturated, double radius, double disp_hint, double *__restrict__ cent, const double *__restrict__ cent0,
    const double *__restrict__ disp0, unsigned long long *sizes, long long *__restrict__ assign, double *__restrict__ dub,
    unsigned long long *maxdisp_bits, int *fail, unsigned long long *probe) {
    // probe (or null): per block {centroid, rows, start ns, end ns} (option cluster_chain_probe)
    // cent: the state the chains start from (in) and leave behind (out); cent0: the snapshot the rows were ranked
    // against; disp0 (or null when the two coincide): |cent - cent0| per centroid on entry
    extern __shared__ __align__(128) unsigned char rp_smem[];
    __shared__ ChainShared sh;
    constexpr int FP = 32 * CW * NPW;
    constexpr int NS = chain_ring_slots((FP + 127) / 128, RG);
    constexpr int NP = NS % 4 == 0 ? 4 : 2;   // producer warps in use (<= kChainProducers)
    static_assert(NS % NP == 0 && NP <= kChainProducers, "slot ownership");
    static_assert(NS <= kRingSlots && RG <= 8 && RG >= 2, "ring geometry");
    double *ring = reinterpret_cast<double *>(rp_smem);
    SegMeta *metas = reinterpret_cast<SegMeta *>(rp_smem + (size_t)NS * RG * FP * sizeof(double));
    GroupDesc<RG> *descs = reinterpret_cast<GroupDesc<RG> *>(metas + NS * RG);
    const int c = order[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int beg = seg_off[c], end = seg_off[c + 1];
    if (beg == end) return;
    const int ngroups = (end - beg + RG - 1) / RG;
    unsigned long long t_start = 0;
    if (probe && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            rp_mbar_init(&sh.full[s], 1);
            rp_mbar_init(&sh.empty[s], CW);
        }
        sh.drops = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= CW) {
        // ---- producer: one slot = RG rows + their metadata + the group descriptor.  The count a group starts from is
        // predictable: every row of the chain adds one unless it is dropped (rare; the consumers publish their tally,
        // a descriptor built from a stale tally is recognised by its kd_pred and ignored)
        const double kd0 = (double)sizes[c];
        // the metadata of 32 rows (32 / RG groups) per coalesced load, one batch ahead of the groups being issued: the
        // producer's own global-memory latency stays off the ring's critical path
        constexpr int GPB = 32 / RG;   // groups per batch
        // Several producer warps: one warp cannot issue a group's descriptor and bulk copies (a loop over the lanes: the
        // copy instruction takes uniform operands) in the time the consumers need for the group -- with one producer the
        // longest chain paid 138 ns per row whatever the ring depth or the consumers' instruction count, with two 103.
        // Producer p owns the groups g = p (mod NP), hence the slots s = p (mod NP) (NS is a multiple of NP): every slot
        // is refilled by the same warp each time round, so no waiter ever lags its mbarrier by more than one phase.
        const int pw = warp - CW;
        if (pw >= NP) return;
        const int my_groups = ngroups > pw ? (ngroups - pw + NP - 1) / NP : 0;   // groups pw, pw + NP, ...
        const int nbatches = (my_groups + GPB - 1) / GPB;
        auto batch_row = [&](int bt) {   // position of this lane's row in batch bt of this producer
            const int g = pw + (bt * GPB + lane / RG) * NP;
            return beg + g * RG + lane % RG;
        };
        SegMeta cur, nxt;
        cur.dlo = cur.dhi = cur.slo = 0.0;
        cur.row = cur.pad = 0;
        nxt = cur;
        if (nbatches > 0 && batch_row(0) < end) nxt = seg_meta[batch_row(0)];
        for (int bt = 0; bt < nbatches; ++bt) {
            cur = nxt;
            if (bt + 1 < nbatches) {
                const int np = batch_row(bt + 1);
                if (np < end) {
                    nxt = seg_meta[np];
                    rp_prefetch_l2(rows + (size_t)nxt.row * f, (unsigned)(f * 8));   // the ring's copies then hit L2
                }
            }
#pragma unroll 1
            for (int gi = 0; gi < GPB; ++gi) {
                const int g = pw + (bt * GPB + gi) * NP;
                if (g >= ngroups) break;
                const int slot = g % NS;
                if (g >= NS) rp_mbar_wait(&sh.empty[slot], ((g / NS) - 1) & 1);
                const int pos = beg + g * RG;
                const int nr = end - pos < RG ? end - pos : RG;
                const int drops = atomicAdd(&sh.drops, 0);   // (an atomic read: the tally may be stale, never torn)
                const double kdp = kd0 + (double)(g * RG - drops);
                const int src = gi * RG + (lane < RG ? lane : 0);
                const int r = __shfl_sync(0xffffffffu, cur.row, src);
                double dh = __shfl_sync(0xffffffffu, cur.dhi, src);
                double sl = __shfl_sync(0xffffffffu, cur.slo, src) - dh;
                if (lane >= nr) {
                    dh = -INFINITY;
                    sl = INFINITY;
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {   // RG <= 8
                    dh = fmax(dh, __shfl_xor_sync(0xffffffffu, dh, o));
                    sl = fmin(sl, __shfl_xor_sync(0xffffffffu, sl, o));
                }
                if (lane <= RG) descs[slot].yk[lane] = __drcp_rn(kdp + (double)(lane + 1));
                if (lane == 0) {
                    descs[slot].hmax = dh;
                    descs[slot].minslack = sl;
                    descs[slot].kd_pred = kdp;
                }
                __syncwarp();
                if (lane == 0) rp_mbar_expect_tx(&sh.full[slot], (unsigned)(nr * (f * 8 + (int)sizeof(SegMeta))));   // (release)
                __syncwarp();
                if (lane < nr)
                    rp_bulk_g2s(ring + ((size_t)slot * RG + lane) * FP, rows + (size_t)r * f, (unsigned)(f * 8), &sh.full[slot]);
                if (lane == 0) rp_bulk_g2s(metas + slot * RG, seg_meta + pos, (unsigned)(nr * sizeof(SegMeta)), &sh.full[slot]);
            }
        }
        return;
    }

    // ---- consumers.  Element u of this lane = feature 32 CW u + 32 warp + lane
    constexpr int EP = 32 * CW;   // features per element index
    const int j0 = 32 * warp + lane;
    const bool last_valid = EP * (NPW - 1) + j0 < f;   // NPW = ceil(f / EP): only the last element can be padding
    double cr[NPW], s0[NPW];
#pragma unroll
    for (int u = 0; u < NPW; ++u) {
        const bool v = u < NPW - 1 || last_valid;
        s0[u] = v ? cent0[(size_t)c * f + EP * u + j0] : 0.0;   // padding holds zeros everywhere: no effect
        cr[u] = v ? cent[(size_t)c * f + EP * u + j0] : 0.0;
    }
    double kd = (double)sizes[c];
    double y_next = __drcp_rn(kd + 1.0);
    double B = disp0 ? disp0[c] : 0.0, dmax = B;
    int since = 0, par = 0;
    bool bad = false, go = true;
    if (disp_hint < 0.0) {   // (CTAs that start later may see a larger value: more exact steps, same bits)
        const double md0 = __longlong_as_double((long long)*(volatile unsigned long long *)maxdisp_bits);
        disp_hint = (md0 == md0) ? 2.0 * md0 : INFINITY;
    }
    const double guard = 1e-9 * radius;
    const double thr_upd = (saturated ? radius : 0.5 * radius) * (1.0 - 1e-9);   // (dhi + B)^2 below: an update for sure
    const double thr_drop = 1.5 * radius * (1.0 + 1e-9);                         // (dlo - B)^2 above: dropped for sure

    // all consumer warps reduce `v` to the same bits: butterfly inside the warp, fixed-order sum of the CW partials
    auto block_sum = [&](double v, bool want_stop) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh.part[par][warp] = v;
        if (want_stop && threadIdx.x == 0) sh.stop[par] = *(volatile int *)fail;
        asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < CW; i += 4)   // the same fixed order in every warp
            t += (sh.part[par][i] + sh.part[par][i + 1]) + (sh.part[par][i + 2] + sh.part[par][i + 3]);
        if (want_stop && sh.stop[par] != 0) go = false;   // the chunk is lost already: stop early
        par ^= 1;
        return t;
    };

    unsigned n_exact = 0;
    auto step = [&](const double(&X)[NPW], const SegMeta &mt) {
        const double hi = mt.dhi + B, lo = mt.dlo - B;
        const double hi2 = hi * hi;
        int cls;                  // 0 update, 1 count only, 2 dropped
        double d_up = hi;         // what the certification sees: an upper bound of |x - c_now|
        double a[NPW];
#pragma unroll
        for (int u = 0; u < NPW; ++u) a[u] = __dsub_rn(X[u], cr[u]);   // clustering.rs:749 (and the diff of :917)
        if (hi2 < thr_upd && hi + disp_hint < mt.slo) {
            cls = 0;
        } else if (saturated && lo > 0.0 && lo * lo > thr_drop) {
            cls = 2;   // (while centroids can still be opened a far row opens one, :672: the exact step below says so)
        } else {
            double p = 0.0;
#pragma unroll
            for (int u = 0; u < NPW; ++u) p = fma(a[u], a[u], p);
            const double d2 = block_sum(p, false);   // ~1e-15 relative, NOT the reference's summation order: guard band
            ++n_exact;
            if (!(d2 == d2) || fabs(d2 - radius) <= guard || fabs(d2 - 1.5 * radius) <= guard ||
                (!saturated && fabs(d2 - 0.5 * radius) <= guard))
                bad = true;
            if (!saturated && d2 > 0.5 * radius) {   // the walk would open a new centroid here (:672): not replayable
                bad = true;
                go = false;
                return;
            }
            cls = d2 <= radius ? 0 : (d2 <= 1.5 * radius ? 1 : 2);
            d_up = fmin(hi, sqrt(d2) * (1.0 + 1e-12));
        }
        if (cls == 0) {
            kd += 1.0;
            const double y = y_next;
            bool slow = false;
#pragma unroll
            for (int u = 0; u < NPW; ++u) {
                const unsigned h = (unsigned)__double2hiint(a[u]) & 0x7fffffffu;   // |a| in [~1e-280, ~1e280]?
                const bool out = h - 0x05d00000u > 0x74200000u;
                slow |= (u < NPW - 1) ? out : (out && last_valid);
            }
            if (!__any_sync(0xffffffffu, slow)) {
#pragma unroll
                for (int u = 0; u < NPW; ++u) {
                    const double q0 = __dmul_rn(a[u], y);
                    const double q1 = __fma_rn(__fma_rn(-q0, kd, a[u]), y, q0);
                    cr[u] = __dadd_rn(cr[u], __fma_rn(__fma_rn(-q1, kd, a[u]), y, q1));   // clustering.rs:747-751
                }
            } else {
#pragma unroll
                for (int u = 0; u < NPW; ++u) {
                    const unsigned h = (unsigned)__double2hiint(a[u]) & 0x7fffffffu;
                    double q;
                    if (h - 0x05d00000u > 0x74200000u) {
                        q = a[u] == 0.0 ? a[u] / kd : __ddiv_rn(a[u], kd);
                    } else {
                        const double q0 = __dmul_rn(a[u], y);
                        const double q1 = __fma_rn(__fma_rn(-q0, kd, a[u]), y, q0);
                        q = __fma_rn(__fma_rn(-q1, kd, a[u]), y, q1);
                    }
                    cr[u] = __dadd_rn(cr[u], q);
                }
            }
            B = fma(hi * y, 1.0 + 1e-9, B);   // the centroid moved by |x - c| / k <= hi / k
            dmax = fmax(dmax, B);
            y_next = __drcp_rn(kd + 1.0);
            ++since;
            if (since >= chain_interval(kd)) {   // B back to the exact displacement
                double p = 0.0;
#pragma unroll
                for (int u = 0; u < NPW; ++u) {
                    const double e = cr[u] - s0[u];
                    p = fma(e, e, p);
                }
                const double t = block_sum(p, true);
                const double nb = sqrt(t) * (1.0 + 1e-12);
                if (!(nb == nb) || !(nb <= 1e300)) bad = true;
                else B = fmin(B, nb);
                since = 0;
            }
        } else if (cls == 1) {
            kd += 1.0;
            y_next = __drcp_rn(kd + 1.0);
        }
        if (threadIdx.x == 0) {
            assign[mt.row] = cls == 2 ? -1ll : (long long)c;
            dub[mt.row] = d_up;
            if (cls == 2) atomicAdd(&sh.drops, 1);   // the producer's count prediction
        }
    };

    int done = 0;   // rows applied
    unsigned n_grouped = 0, n_byrow = 0, n_ckpt = 0;   // diagnostics (thread 0 publishes them)
    for (int g = 0; g < ngroups; ++g) {
        const int slot = g % NS;
        rp_mbar_wait(&sh.full[slot], (g / NS) & 1);
        const int nr = end - (beg + g * RG) < RG ? end - (beg + g * RG) : RG;
        if (go) {
            const double *xrow = ring + (size_t)slot * RG * FP + j0;
            const SegMeta *mrow = metas + slot * RG;
            // ---- the group path: all RG rows of the slot are certain updates (by far the common case once the counts
            // are large).  Everything scalar comes from the producer's descriptor (reciprocals of the next RG + 1 counts,
            // the largest dhi and the thinnest margin of the group); the displacement bound of the whole group follows
            // from one product:  with u = 1.01 (hmax + B) / (k + 1)  every step t of the group keeps B_t <= B + t u
            // (induction: the increment (dhi_t + B_{t-1}) / (k + t) (1 + 1e-9) is at most
            // (hmax + B + (t - 1) u) / (k + 1) (1 + 1e-9) <= u (1 + 7 * 1.01 / 1024) (1 + 1e-9) / 1.01 < u  for
            // k >= 1024, RG <= 8), so the element chains c += (x - c) / k run back
            // to back: a row costs about its own dependent latency (sub, mul, 4 FMA, add).  The exponent-range test of
            // the division is voted on once per group; a hit restores the centroid and replays the slot row by row.
            bool grouped = false;
            const GroupDesc<RG> &gd = descs[slot];
            const int interval = chain_interval(kd);
            if (nr == RG && kd >= 64.0 && gd.kd_pred == kd) {
                if (since + RG > interval) {   // B back to the exact displacement before the group, not in the middle
                    double p = 0.0;
#pragma unroll
                    for (int u = 0; u < NPW; ++u) {
                        const double e = cr[u] - s0[u];
                        p = fma(e, e, p);
                    }
                    const double tt = block_sum(p, true);
                    const double nb = sqrt(tt) * (1.0 + 1e-12);
                    if (!(nb == nb) || !(nb <= 1e300)) bad = true;
                    else B = fmin(B, nb);
                    since = 0;
                    ++n_ckpt;
                }
                // (k >= 1024: factor 1.01 as derived above; 64 <= k < 1024: 1.15 >= (1 + 7 * 1.15 / 65) (1 + 1e-9).  Small
                // counts get an exact B before every group -- `interval` < RG -- and the margin test below decides
                // whether 8 steps of growth are affordable)
                const double uinc = (gd.hmax + B) * gd.yk[0] * (kd >= 1024.0 ? 1.01 : 1.15);
                const double Bend = fma((double)RG, uinc, B);
                const double himax = gd.hmax + Bend;
                if (himax * himax < thr_upd && Bend + disp_hint < gd.minslack) {
                    double save[NPW];
                    bool slow = false;
#pragma unroll
                    for (int u = 0; u < NPW; ++u) save[u] = cr[u];
#pragma unroll
                    for (int t = 0; t < RG; ++t) {
                        const double y = gd.yk[t], nkk = -(kd + (double)(t + 1));
                        double a[NPW], q[NPW], r[NPW];
                        // (fma(q, -k, a) = fma(-q, k, a) bit for bit: the product is exact either way)
#pragma unroll
                        for (int u = 0; u < NPW; ++u) {
                            const double xv = (u < NPW - 1 || last_valid) ? xrow[(size_t)t * FP + EP * u] : 0.0;
                            a[u] = rp_vsub(xv, cr[u]);                      // clustering.rs:749
                        }
#pragma unroll
                        for (int u = 0; u < NPW; ++u) q[u] = rp_vmul(a[u], y);
#pragma unroll
                        for (int u = 0; u < NPW; ++u) r[u] = rp_vfma(q[u], nkk, a[u]);
#pragma unroll
                        for (int u = 0; u < NPW; ++u) q[u] = rp_vfma(r[u], y, q[u]);
#pragma unroll
                        for (int u = 0; u < NPW; ++u) r[u] = rp_vfma(q[u], nkk, a[u]);
#pragma unroll
                        for (int u = 0; u < NPW; ++u) q[u] = rp_vfma(r[u], y, q[u]);
#pragma unroll
                        for (int u = 0; u < NPW; ++u) cr[u] = rp_vadd(cr[u], q[u]);   // :747-751
#pragma unroll
                        for (int u = 0; u < NPW; ++u) {
                            const unsigned h = (unsigned)__double2hiint(a[u]) & 0x7fffffffu;
                            const bool out = h - 0x05d00000u > 0x74200000u;
                            slow |= (u < NPW - 1) ? out : (out && last_valid);
                        }
                    }
                    if (!__any_sync(0xffffffffu, slow)) {
                        grouped = true;
                        n_grouped += RG;
                        kd += (double)RG;
                        y_next = gd.yk[RG];
                        B = Bend;
                        dmax = fmax(dmax, B);
                        since += RG;
                        done += RG;
                        if (warp == 0 && lane < RG) {
                            const SegMeta m = mrow[lane];
                            assign[m.row] = (long long)c;
                            dub[m.row] = m.dhi + Bend;   // an upper bound of the row's distance at its own time
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < NPW; ++u) cr[u] = save[u];   // an element needs the careful division
                    }
                }
            }
            if (!grouped) {
#pragma unroll 1
                for (int t = 0; t < nr && go; ++t) {
                    double X[NPW];
#pragma unroll
                    for (int u = 0; u < NPW; ++u) X[u] = (u < NPW - 1 || last_valid) ? xrow[(size_t)t * FP + EP * u] : 0.0;
                    const SegMeta mt = mrow[t];
                    step(X, mt);
                    ++n_byrow;
                    if (go) ++done;
                }
            }
        }
        __syncwarp();
        if (lane == 0) rp_mbar_arrive(&sh.empty[slot]);   // (a lost chain keeps draining its ring: the producer must finish)
    }
#pragma unroll
    for (int u = 0; u < NPW; ++u)
        if (u < NPW - 1 || last_valid) cent[(size_t)c * f + EP * u + j0] = cr[u];
    if (threadIdx.x == 0) {
        sizes[c] = (unsigned long long)kd;
        const double dm = dmax * (1.0 + 1e-12);
        if (dm == dm) atomicMax(maxdisp_bits, (unsigned long long)__double_as_longlong(dm));
        else bad = true;
        if (bad || !go || done < end - beg) atomicOr(fail, 1);
        if (probe) {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            probe[4 * blockIdx.x] = (unsigned long long)c;
            probe[4 * blockIdx.x + 1] = (unsigned long long)(end - beg);
            probe[4 * blockIdx.x + 2] = t_start;
            probe[4 * blockIdx.x + 3] = t_end;
        }
        atomicAdd(maxdisp_bits + 1, (unsigned long long)n_grouped);
        atomicAdd(maxdisp_bits + 2, (unsigned long long)n_byrow);
        atomicAdd(maxdisp_bits + 3, (unsigned long long)n_exact);
        atomicAdd(maxdisp_bits + 4, (unsigned long long)n_ckpt);
    }
}

// per sorted position: the row, certified bounds [dlo, dhi] of its distance to the SNAPSHOT of its nearest centroid and
// a certified lower bound of its distance to the runner-up (GEMM-form squared distance |q|^2 + |c|^2 - 2 q.c from the
// FP64 tensor pipe: absolute error below e2, the bound the certification kernel uses)
__global__ void __launch_bounds__(256) replay_bounds_kernel(const int *__restrict__ rows_sorted, int m, int f,
                                                            const double *__restrict__ top2_dist,
                                                            const long long *__restrict__ top2_cnt,
                                                            const double *__restrict__ qn2,
                                                            const unsigned long long *cn2max_bits,
                                                            const double *__restrict__ near_b,
                                                            SegMeta *__restrict__ meta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int r = rows_sorted[i];
    SegMeta mt;
    if (near_b) {   // certified bounds straight from the tcgen05 ranking pass (search_umma.cuh, PF_NEAR)
        mt.dlo = near_b[3 * (size_t)r];
        mt.dhi = near_b[3 * (size_t)r + 1];
        mt.slo = near_b[3 * (size_t)r + 2];
    } else {
        const double cn2max = __longlong_as_double((long long)*cn2max_bits);
        const double e2 = 1e-15 * (double)(f + 32) * (qn2[r] + cn2max);
        const double d0 = top2_dist[2 * (size_t)r], d1 = top2_dist[2 * (size_t)r + 1];
        mt.dlo = sqrt(fmax(d0 * d0 - e2, 0.0)) * (1.0 - 1e-15);
        mt.dhi = sqrt(d0 * d0 + e2) * (1.0 + 1e-15);
        mt.slo = top2_cnt[r] >= 2 ? sqrt(fmax(d1 * d1 - e2, 0.0)) * (1.0 - 1e-15) : 0.0;
    }
    if (!(mt.dhi == mt.dhi) || !(mt.dlo == mt.dlo) || !(mt.slo == mt.slo)) {   // no usable distance: exact steps only
        mt.dlo = 0.0;
        mt.dhi = 1e300;
        mt.slo = 0.0;
    }
    mt.row = r;
    mt.pad = 0;
    meta[i] = mt;
}

__global__ void __launch_bounds__(256) replay_certify_kernel(const double *__restrict__ dcur, const double *__restrict__ top2_dist,
                                                             const long long *__restrict__ top2_cnt,
                                                             const double *__restrict__ qn2, const unsigned long long *cn2max_bits,
                                                             const unsigned long long *maxdisp_bits,
                                                             const long long *__restrict__ assign, double radius, int f,
                                                             int m, const double *__restrict__ near_b, int *fail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const double maxdisp = __longlong_as_double((long long)*maxdisp_bits);
    if (near_b) {   // bounds from the tcgen05 ranking pass: slo / dlo are already certified lower bounds
        const double lower_n = near_b[3 * (size_t)r + 2] - maxdisp;
        const double lb_n = fmax(near_b[3 * (size_t)r] - maxdisp, 0.0);
        const bool dropped_n = lb_n * lb_n > 1.5 * radius * (1.0 + 1e-9) && assign[r] == -1;
        if (!(dcur[r] < lower_n || dropped_n)) atomicOr(fail, 2);
        return;
    }
    const double cn2max = __longlong_as_double((long long)*cn2max_bits);
    // GEMM-form squared distance |q|^2 + |c|^2 - 2 q.c from the FP64 tensor pipe: absolute error below e2
    const double e2 = 1e-15 * (double)(f + 32) * (qn2[r] + cn2max);
    const double s = top2_dist[2 * (size_t)r + 1];
    const double lower = sqrt(fmax(s * s - e2, 0.0)) - maxdisp;
    // A row farther than sqrt(1.5 radius) from EVERY centroid at its own time is dropped whichever centroid is the
    // nearest (no update, no count, assignment None: clustering.rs:759-815) -- e.g. a blob no centroid was opened for.
    const double b = top2_dist[2 * (size_t)r];
    const double lb = fmax(sqrt(fmax(b * b - e2, 0.0)) - maxdisp, 0.0);
    const bool dropped_anyway = lb * lb > 1.5 * radius * (1.0 + 1e-9) && assign[r] == -1;
    // dcur: the row's distance to its centroid at the row's own time, or an upper bound of it
    if (top2_cnt[r] < 2 || !(dcur[r] < lower || dropped_anyway)) atomicOr(fail, 2);
}

// ---- the creator run at the start of a fresh walk --------------------------------------------------------------
// While every row opens a new centroid (len < max_clusters and d^2 > radius / 2 to everything opened so far,
// clustering.rs:672) the centroids ARE the rows, nothing is averaged, and the decisions of the first P rows follow
// from their pairwise distances alone: row r opens a centroid iff no earlier row lies within radius / 2 of it.  One
// parallel pass evaluates all pairs in the reference's arithmetic (sequential separately rounded (a - b)^2 sums,
// :917-921) and returns the first row that does NOT open one; the walk proper starts there with rows 0 .. G-1 as
// centroids of size 1.  On the bench data that is the whole growth phase (384 centroids in the first 384 rows), which
// costs the sequential kernel 7 ms because every new centroid ends one of its blocks.
constexpr int kGrowTile = 16, kGrowChunk = 32;
__global__ void __launch_bounds__(kGrowTile *kGrowTile) growth_pairs_kernel(const double *__restrict__ rows, int p, int f,
                                                                            double half_radius, int *first_noncreator) {
    if (blockIdx.x > blockIdx.y) return;   // pairs (r, r') with r' < r only: tile column <= tile row
    __shared__ double sa[kGrowTile][kGrowChunk + 1], sb[kGrowTile][kGrowChunk + 1];
    const int tx = threadIdx.x % kGrowTile, ty = threadIdx.x / kGrowTile;
    const int r = blockIdx.y * kGrowTile + ty, q = blockIdx.x * kGrowTile + tx;
    double d2 = 0.0;
    for (int j0 = 0; j0 < f; j0 += kGrowChunk) {
        for (int e = threadIdx.x; e < kGrowTile * kGrowChunk; e += kGrowTile * kGrowTile) {
            const int rr = e / kGrowChunk, jj = e % kGrowChunk;
            const int ra = blockIdx.y * kGrowTile + rr, rb = blockIdx.x * kGrowTile + rr;
            sa[rr][jj] = (ra < p && j0 + jj < f) ? rows[(size_t)ra * f + j0 + jj] : 0.0;
            sb[rr][jj] = (rb < p && j0 + jj < f) ? rows[(size_t)rb * f + j0 + jj] : 0.0;
        }
        __syncthreads();
        const int lim = f - j0 < kGrowChunk ? f - j0 : kGrowChunk;
        for (int jj = 0; jj < lim; ++jj) {
            const double diff = __dsub_rn(sa[ty][jj], sb[tx][jj]);
            d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
        }
        __syncthreads();
    }
    // d2 > radius * 0.5 opens a centroid; NaN compares false both ways: a NaN distance is never the nearest (:922) and
    // a row with only NaN distances sees d2 = +inf, so it opens one -- "not within half_radius" in every case
    if (r < p && q < r && d2 <= half_radius) atomicMin(first_noncreator, r);
}

__global__ void __launch_bounds__(256) growth_fill_kernel(int g, unsigned long long *__restrict__ sizes,
                                                          long long *__restrict__ assign) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < g) {
        sizes[c] = 1ull;
        assign[c] = c;
    }
}

// |start_c - snap_c| per centroid (rounded up), and its maximum: the displacement the chains of a chunk start with when
// the rows were ranked against an OLDER snapshot than the state they are applied to (row-sharded build)
__global__ void __launch_bounds__(128) replay_disp0_kernel(const double *__restrict__ start, const double *__restrict__ snap,
                                                           int K, int f, double *__restrict__ disp0,
                                                           unsigned long long *maxdisp_bits, int *fail) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= K) return;
    double acc = 0.0;
    for (int j = lane; j < f; j += 32) {
        const double e = start[(size_t)c * f + j] - snap[(size_t)c * f + j];
        acc = fma(e, e, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const double d = sqrt(acc) * (1.0 + 1e-12);
        disp0[c] = d;
        if (d == d && d <= 1e300) atomicMax(maxdisp_bits, (unsigned long long)__double_as_longlong(d));
        else atomicOr(fail, 1);
    }
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
    DevTmp<double> qn2, xn2, dist, dcur, cent_tmp, disp0;
    DevTmp<int64_t> idx, cnt, minus1;
    DevTmp<int> keys, vals, keys_s, vals_s, seg_off, order, flags;
    DevTmp<unsigned long long> sizes_tmp, scal;   // scal[0] = max |c|^2 bits, scal[1] = max displacement bits
    DevTmp<unsigned char> cub_tmp;
    DevTmp<SegMeta> meta;
    DevTmp<int64_t> near_idx;
    DevTmp<double> near_b;
    bool use_near = false;   // the current preparation ranked with the tcgen05 tile (certified bounds, no exact top-2)
    size_t cub_bytes = 0;
    int cap_m = 0, cap_k = 0;
    double last_disp = INFINITY;   // largest centroid displacement of the previous (proven) chunk
    int last_flags = 0;            // fail flags of the last run (bit 1: a row missed its certificate, bit 0: not replayable)
};

int replay_ws_init(asb_ctx *ctx, ReplayWs &w, int m, int K, int f) {
    w.cap_m = m;
    w.cap_k = K;
    ASB_TRY(w.qn2.init(ctx, (size_t)m));
    ASB_TRY(w.xn2.init(ctx, (size_t)K));
    ASB_TRY(w.dist.init(ctx, (size_t)m * 2));
    ASB_TRY(w.dcur.init(ctx, (size_t)m));
    ASB_TRY(w.cent_tmp.init(ctx, (size_t)K * f));
    ASB_TRY(w.disp0.init(ctx, (size_t)K));
    ASB_TRY(w.idx.init(ctx, (size_t)m * 2));
    ASB_TRY(w.cnt.init(ctx, (size_t)m));
    ASB_TRY(w.minus1.init(ctx, (size_t)m));
    ASB_TRY(w.keys.init(ctx, (size_t)m));
    ASB_TRY(w.vals.init(ctx, (size_t)m));
    ASB_TRY(w.keys_s.init(ctx, (size_t)m));
    ASB_TRY(w.vals_s.init(ctx, (size_t)m));
    ASB_TRY(w.seg_off.init(ctx, (size_t)K + 1));
    ASB_TRY(w.order.init(ctx, (size_t)K));
    ASB_TRY(w.meta.init(ctx, (size_t)m));
    ASB_TRY(w.near_idx.init(ctx, (size_t)m));
    ASB_TRY(w.near_b.init(ctx, (size_t)3 * m));
    ASB_TRY(w.flags.init(ctx, 2));
    ASB_TRY(w.sizes_tmp.init(ctx, (size_t)K));
    ASB_TRY(w.scal.init(ctx, 6));   // + [2..5]: chain diagnostics (rows by group / by row, exact steps, checkpoints)
    ASB_CUDA(ctx, cudaMemsetAsync(w.minus1.ptr, 0xff, (size_t)m * sizeof(int64_t), ctx->stream));
    w.cub_bytes = 0;
    ASB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, w.keys.ptr, w.keys_s.ptr, w.vals.ptr, w.vals_s.ptr, m,
                                                  0, 32, ctx->stream));
    ASB_TRY(w.cub_tmp.init(ctx, w.cub_bytes));
    return ASB_OK;
}

// Everything of a chunk that depends only on the SNAPSHOT the rows are ranked against: nearest / runner-up centroid of
// every row, the rows sorted by nearest centroid, the per-position metadata.
int replay_prepare(asb_ctx *ctx, ReplayWs &w, const double *rows_d, int m, int f, int K, const double *snap_d,
                   bool allow_near = true) {
    asb_wait_rows(ctx, rows_d + (size_t)m * f);
    ASB_CUDA(ctx, cudaMemsetAsync(w.flags.ptr, 0, 2 * sizeof(int), ctx->stream));
    ASB_CUDA(ctx, cudaMemsetAsync(w.scal.ptr, 0, 6 * sizeof(unsigned long long), ctx->stream));
    ASB_TRY(asb_dev_norms2(ctx, rows_d, m, f, w.qn2.ptr));
    ASB_TRY(asb_dev_norms2(ctx, snap_d, K, f, w.xn2.ptr));
    replay_max_kernel<<<1, 256, 0, ctx->stream>>>(w.xn2.ptr, K, w.scal.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_max_kernel"));
    // ranking: the tcgen05 tile with certified distance bounds (option "cluster_replay_near", default 1), else -- and
    // again whenever its bounds turn out too wide to certify a chunk -- the FP64 tensor kernel's exact top-2
    w.use_near = false;
    if (allow_near && opt_or(ctx, "cluster_replay_near", 1.0) != 0.0 && m >= 4096) {
        bool done = false;
        ASB_TRY(asb_dev_near_tf32(ctx, rows_d, m, f, snap_d, K, w.qn2.ptr, w.xn2.ptr, w.near_idx.ptr, w.near_b.ptr, &done));
        w.use_near = done;
    }
    if (!w.use_near)
        ASB_TRY(asb_dev_top2_l2(ctx, rows_d, m, f, snap_d, K, w.qn2.ptr, w.xn2.ptr, w.minus1.ptr, w.idx.ptr, w.dist.ptr,
                                w.cnt.ptr, w.flags.ptr + 1));
    replay_keys_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(w.use_near ? w.near_idx.ptr : w.idx.ptr, w.use_near ? 1 : 2, m,
                                                                 K, w.keys.ptr, w.vals.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_keys_kernel"));
    int bits = 1;
    while ((1ll << bits) < (long long)K + 1) ++bits;
    ASB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(w.cub_tmp.ptr, w.cub_bytes, w.keys.ptr, w.keys_s.ptr, w.vals.ptr,
                                                  w.vals_s.ptr, m, 0, bits, ctx->stream));   // stable: rows stay ascending
    ctx->launches++;
    replay_offsets_kernel<<<(K + 1 + 255) / 256, 256, 0, ctx->stream>>>(w.keys_s.ptr, m, K, w.seg_off.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_offsets_kernel"));
    replay_order_kernel<<<1, 1024, 0, ctx->stream>>>(w.seg_off.ptr, K, w.order.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_order_kernel"));
    replay_bounds_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(w.vals_s.ptr, m, f, w.dist.ptr, (const long long *)w.cnt.ptr,
                                                                   w.qn2.ptr, w.scal.ptr, w.use_near ? w.near_b.ptr : nullptr,
                                                                   w.meta.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_bounds_kernel"));
    return ASB_OK;
}

// The chains and the certification of a prepared chunk, from the state in centroids_d / sizes_d (snap_d == centroids_d
// on one GPU; an older snapshot in the row-sharded build: the chains then start with the displacement |start - snap|).
// *ok = 1: centroids / sizes / assign hold the walk's state after the chunk; 0: nothing was touched except assign (the
// caller's sequential walk rewrites it).
int replay_run(asb_ctx *ctx, ReplayWs &w, const double *rows_d, int m, int f, int K, int saturated, double radius,
               const double *snap_d, double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d, int *ok) {
    *ok = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(w.cent_tmp.ptr, centroids_d, (size_t)K * f * sizeof(double), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(w.sizes_tmp.ptr, sizes_d, (size_t)K * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    const double *disp0 = nullptr;
    if (snap_d != centroids_d) {
        replay_disp0_kernel<<<(unsigned)((K + 3) / 4), 128, 0, ctx->stream>>>(centroids_d, snap_d, K, f, w.disp0.ptr,
                                                                             w.scal.ptr + 1, w.flags.ptr);
        ASB_TRY(asb_check_launch(ctx, "replay_disp0_kernel"));
        disp0 = w.disp0.ptr;
    }
    unsigned long long *probe_d = nullptr;
    DevTmp<unsigned long long> probe;
    if (opt_or(ctx, "cluster_chain_probe", 0.0) != 0.0) {
        ASB_TRY(probe.init(ctx, (size_t)K * 4));
        ASB_CUDA(ctx, cudaMemsetAsync(probe.ptr, 0, (size_t)K * 4 * sizeof(unsigned long long), ctx->stream));
        probe_d = probe.ptr;
    }
    {
        KernelTimer kt(ctx, "cluster_chain_kernel");
        // the ring kernel needs 16-byte aligned rows of a multiple of 16 bytes (bulk copies) and f <= 1024
        const bool generic = f > 1024 || (f & 1) || (((uintptr_t)rows_d) & 15) ||
                             opt_or(ctx, "cluster_replay_generic_chain", 0.0) != 0.0;
        // the displacement the previous chunk saw, doubled (inf at first); negative: the kernel takes twice the largest
        // |start - snapshot| instead (row-sharded build: the drift since the common snapshot dominates)
        const double hint = w.last_disp < 0.0 ? -1.0 : 2.0 * w.last_disp;
#define ASB_CHAIN_ARGS                                                                                                  \
    K, saturated, radius, w.cent_tmp.ptr, snap_d, disp0, w.sizes_tmp.ptr, (long long *)assign_d, w.dcur.ptr,           \
        w.scal.ptr + 1, w.flags.ptr
#define ASB_CHAIN_TMA(CW, NPW, RG)                                                                                     \
    {                                                                                                                   \
        constexpr int fp_ = 32 * CW * NPW;                                                                              \
        const size_t smem = (size_t)chain_ring_slots((fp_ + 127) / 128, RG) *                                           \
                            (RG * (fp_ * sizeof(double) + sizeof(SegMeta)) + sizeof(GroupDesc<RG>));                    \
        ASB_CUDA(ctx, cudaFuncSetAttribute(replay_chain_tma_kernel<CW, NPW, RG>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        replay_chain_tma_kernel<CW, NPW, RG><<<(unsigned)K, (CW + kChainProducers) * 32, smem, ctx->stream>>>(         \
            rows_d, f, w.seg_off.ptr, w.order.ptr, w.meta.ptr, K, saturated, radius, hint, w.cent_tmp.ptr, snap_d, disp0, \
            w.sizes_tmp.ptr, (long long *)assign_d, w.dcur.ptr, w.scal.ptr + 1, w.flags.ptr, probe_d);                 \
    }
        if (generic)
            replay_chain_kernel<<<(unsigned)((K + 3) / 4), 128, 0, ctx->stream>>>(rows_d, f, w.seg_off.ptr, w.vals_s.ptr,
                                                                                  ASB_CHAIN_ARGS);
        else if (f <= 128) ASB_CHAIN_TMA(4, 1, 8)
        else if (f <= 256) ASB_CHAIN_TMA(8, 1, 8)
        else if (f <= 384) ASB_CHAIN_TMA(12, 1, 8)
        else if (f <= 512) ASB_CHAIN_TMA(16, 1, 8)
        else if (f <= 768) ASB_CHAIN_TMA(12, 2, 4)
        else ASB_CHAIN_TMA(16, 2, 4)
#undef ASB_CHAIN_TMA
#undef ASB_CHAIN_ARGS
    }
    ASB_TRY(asb_check_launch(ctx, "replay_chain_kernel"));
    replay_certify_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(w.dcur.ptr, w.dist.ptr, (const long long *)w.cnt.ptr,
                                                                    w.qn2.ptr, w.scal.ptr, w.scal.ptr + 1,
                                                                    (const long long *)assign_d, radius, f, m,
                                                                    w.use_near ? w.near_b.ptr : nullptr, w.flags.ptr);
    ASB_TRY(asb_check_launch(ctx, "replay_certify_kernel"));
    int hflags[2] = {0, 0};
    unsigned long long hsc[5] = {0, 0, 0, 0, 0};
    ASB_CUDA(ctx, cudaMemcpyAsync(hflags, w.flags.ptr, sizeof(hflags), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(hsc, w.scal.ptr + 1, sizeof(hsc), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (probe_d) {   // the largest chunk so far: who was on the critical path?
        std::vector<unsigned long long> hp((size_t)K * 4);
        ASB_CUDA(ctx, cudaMemcpy(hp.data(), probe_d, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull, t1 = 0, longest_rows = 0, longest_ns = 0, last_rows = 0;
        for (int b = 0; b < K; ++b) {
            if (hp[4 * b + 3] == 0) continue;
            if (hp[4 * b + 2] < t0) t0 = hp[4 * b + 2];
            if (hp[4 * b + 3] > t1) {
                t1 = hp[4 * b + 3];
                last_rows = hp[4 * b + 1];
            }
            if (hp[4 * b + 1] > longest_rows) {
                longest_rows = hp[4 * b + 1];
                longest_ns = hp[4 * b + 3] - hp[4 * b + 2];
            }
        }
        if (m >= ctx->kernel_ms["cluster_probe_rows"]) {
            ctx->kernel_ms["cluster_probe_rows"] = (double)m;
            ctx->kernel_ms["cluster_probe_span_us"] = (t1 > t0 ? (double)(t1 - t0) : 0.0) * 1e-3;
            ctx->kernel_ms["cluster_probe_longest_rows"] = (double)longest_rows;
            ctx->kernel_ms["cluster_probe_longest_us"] = (double)longest_ns * 1e-3;
            ctx->kernel_ms["cluster_probe_last_block_rows"] = (double)last_rows;
        }
    }
    const unsigned long long hdisp = hsc[0];
    ctx->kernel_ms["cluster_chain_rows_grouped"] += (double)hsc[1];   // (reset by asb_dev_cluster / the sharded entry)
    ctx->kernel_ms["cluster_chain_rows_by_row"] += (double)hsc[2];
    ctx->kernel_ms["cluster_chain_exact_steps"] += (double)hsc[3];
    ctx->kernel_ms["cluster_chain_checkpoints"] += (double)hsc[4];
    w.last_flags = hflags[0] | (hflags[1] << 8);
    if (hflags[0] != 0 || hflags[1] != 0) {
        w.last_disp = INFINITY;   // the next attempt follows a sequential stretch: no estimate
        return ASB_OK;
    }
    memcpy(&w.last_disp, &hdisp, sizeof(double));
    ASB_CUDA(ctx, cudaMemcpyAsync(centroids_d, w.cent_tmp.ptr, (size_t)K * f * sizeof(double), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(sizes_d, w.sizes_tmp.ptr, (size_t)K * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    *ok = 1;
    return ASB_OK;
}

}  // namespace

// rows 0 .. g-1 of a fresh walk that each open a centroid (see growth_pairs_kernel); *g_out = 0 when not applicable
static int growth_run(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int64_t max_clusters, double radius,
                      double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d, int64_t *g_out) {
    *g_out = 0;
    int64_t p = n < max_clusters ? n : max_clusters;
    if (p > 2048) p = 2048;
    if (p < 8 || !(radius == radius)) return ASB_OK;
    asb_wait_rows(ctx, rows_d + p * f);
    DevTmp<int> first;
    ASB_TRY(first.init(ctx, 1));
    const int big = (int)p;
    ASB_CUDA(ctx, cudaMemcpyAsync(first.ptr, &big, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned t = (unsigned)((p + kGrowTile - 1) / kGrowTile);
    {
        KernelTimer kt(ctx, "cluster_growth_kernel");
        growth_pairs_kernel<<<dim3(t, t), kGrowTile * kGrowTile, 0, ctx->stream>>>(rows_d, (int)p, (int)f, radius * 0.5,
                                                                                   first.ptr);
    }
    ASB_TRY(asb_check_launch(ctx, "growth_pairs_kernel"));
    int g = 0;
    ASB_CUDA(ctx, cudaMemcpyAsync(&g, first.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (g < 2) return ASB_OK;
    ASB_CUDA(ctx, cudaMemcpyAsync(centroids_d, rows_d, (size_t)g * f * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    growth_fill_kernel<<<(g + 255) / 256, 256, 0, ctx->stream>>>(g, sizes_d, (long long *)assign_d);
    ASB_TRY(asb_check_launch(ctx, "growth_fill_kernel"));
    *g_out = g;
    return ASB_OK;
}

int asb_dev_cluster(asb_ctx *ctx, const double *rows_d, int64_t n, int64_t f, int64_t max_clusters, double radius,
                    double *centroids_d, int64_t *assign_d, unsigned long long *sizes_d, int64_t *x_out_host,
                    int64_t init_k) {
    const int64_t prefix = (int64_t)opt_or(ctx, "cluster_replay_prefix", 2048.0);
    const int64_t chunk = (int64_t)opt_or(ctx, "cluster_replay_chunk", 1024.0);
    int64_t chunk_max = (int64_t)opt_or(ctx, "cluster_replay_chunk_max", 262144.0);
    if (chunk_max < chunk) chunk_max = chunk;
    int64_t grow = (int64_t)opt_or(ctx, "cluster_replay_growth", 2.0);   // chunk size factor after a proven chunk
    if (grow < 1) grow = 1;
    if (grow > 16) grow = 16;
    const bool replay = opt_or(ctx, "cluster_replay", 1.0) != 0.0 && prefix >= 1 && chunk >= 256 && chunk_max <= (1 << 24) &&
                        n >= prefix + chunk / 4 && f <= 16384 && max_clusters * f <= (1ll << 27);
    for (const char *key : {"cluster_chain_rows_grouped", "cluster_chain_rows_by_row", "cluster_chain_exact_steps",
                            "cluster_chain_checkpoints", "cluster_probe_rows"})
        ctx->kernel_ms[key] = 0.0;
    ctx->kernel_ms["cluster_replay_chunks"] = 0.0;
    ctx->kernel_ms["cluster_replay_chunks_ok"] = 0.0;
    ctx->kernel_ms["cluster_replay_rows"] = 0.0;
    ctx->kernel_ms["cluster_growth_rows"] = 0.0;
    if (n <= 0 || f <= 0 || max_clusters <= 0 || init_k < 0 || init_k > max_clusters || !replay)
        return asb_dev_cluster_seq(ctx, rows_d, n, f, max_clusters, radius, centroids_d, assign_d, sizes_d, x_out_host, init_k);

    int64_t x = init_k;
    double ms_seq = 0.0, ms_top2 = 0.0, ms_chain = 0.0;   // device time per part, summed over the chunks
    double wall_growth = 0.0, wall_prefix = 0.0, wall_prepare = 0.0, wall_run = 0.0, wall_seq = 0.0;   // host wall, ms
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    int64_t lo = 0;
    if (init_k == 0 && opt_or(ctx, "cluster_growth_run", 1.0) != 0.0) {
        const auto t = now();
        ASB_TRY(growth_run(ctx, rows_d, n, f, max_clusters, radius, centroids_d, assign_d, sizes_d, &lo));
        wall_growth = since(t);
        x = lo;
        ctx->kernel_ms["cluster_growth_rows"] = (double)lo;
    }
    if (lo < prefix) {
        const int64_t x_before = x;
        const auto t = now();
        ASB_TRY(asb_dev_cluster_seq(ctx, rows_d + lo * f, prefix - lo, f, max_clusters, radius, centroids_d, assign_d + lo,
                                    sizes_d, &x, x_before));
        wall_prefix = since(t);
        ms_seq += ktimer_ms(ctx, "cluster_kernel");
        lo = prefix;
    }
    ReplayWs w;
    bool ws_ready = false;
    int fails = 0, tried = 0, proven = 0, near_retries = 0;
    int64_t rows_replayed = 0;
    int64_t cur = chunk;   // rows per attempt: doubles after every proven chunk (one snapshot holds for longer and
                           // longer stretches as the centroids settle), back to `chunk` after a failure
    while (lo < n) {
        int64_t hi = lo + cur < n ? lo + cur : n;
        int ok = 0;
        if (x >= 2) {
            if (!ws_ready) {   // sized once for the largest chunk this call can attempt
                const int64_t cap = chunk_max < n - prefix ? chunk_max : n - prefix;
                ASB_TRY(replay_ws_init(ctx, w, (int)(cap > cur ? cap : cur), (int)max_clusters, (int)f));
                ws_ready = true;
            }
            ++tried;
            auto t = now();
            ASB_TRY(replay_prepare(ctx, w, rows_d + lo * f, (int)(hi - lo), (int)f, (int)x, centroids_d));
            wall_prepare += since(t);
            t = now();
            ASB_TRY(replay_run(ctx, w, rows_d + lo * f, (int)(hi - lo), (int)f, (int)x, x >= max_clusters ? 1 : 0, radius,
                               centroids_d, centroids_d, assign_d + lo, sizes_d, &ok));
            wall_run += since(t);
            if (!ok && w.use_near && w.last_flags == 2) {   // only certificates missed: the tile's bounds may be too wide
                ++near_retries;
                ASB_TRY(replay_prepare(ctx, w, rows_d + lo * f, (int)(hi - lo), (int)f, (int)x, centroids_d, false));
                ASB_TRY(replay_run(ctx, w, rows_d + lo * f, (int)(hi - lo), (int)f, (int)x, x >= max_clusters ? 1 : 0, radius,
                                   centroids_d, centroids_d, assign_d + lo, sizes_d, &ok));
            }
            ms_top2 += ktimer_ms(ctx, "cluster_top2_kernel");
            ms_chain += ktimer_ms(ctx, "cluster_chain_kernel");
        }
        if (ok) {
            fails = 0;
            ++proven;
            rows_replayed += hi - lo;
            cur = cur * grow < chunk_max ? cur * grow : chunk_max;
        } else {
            // not provable (yet): walk sequentially, and twice as far after every further failure in a row -- data that
            // never settles costs O(log(n / chunk)) wasted attempts, data that settles late is picked up when it does
            ++fails;
            const int64_t span = chunk << (fails - 1 < 12 ? fails - 1 : 12);
            hi = lo + span < n ? lo + span : n;
            cur = chunk;
            const int64_t x_before = x;
            const auto t = now();
            ASB_TRY(asb_dev_cluster_seq(ctx, rows_d + lo * f, hi - lo, f, max_clusters, radius, centroids_d, assign_d + lo,
                                        sizes_d, &x, x_before));
            wall_seq += since(t);
            ms_seq += ktimer_ms(ctx, "cluster_kernel");
        }
        lo = hi;
    }
    *x_out_host = x;
    ctx->kernel_ms["cluster_replay_chunks"] = (double)tried;
    ctx->kernel_ms["cluster_replay_chunks_ok"] = (double)proven;
    ctx->kernel_ms["cluster_replay_rows"] = (double)rows_replayed;
    ctx->kernel_ms["cluster_replay_seq_ms"] = ms_seq;
    // host wall time per phase (prepare returns before its kernels end; run ends in a stream synchronisation)
    ctx->kernel_ms["cluster_wall_growth_ms"] = wall_growth;
    ctx->kernel_ms["cluster_wall_prefix_ms"] = wall_prefix;
    ctx->kernel_ms["cluster_wall_prepare_ms"] = wall_prepare;
    ctx->kernel_ms["cluster_wall_run_ms"] = wall_run;
    ctx->kernel_ms["cluster_wall_fallback_ms"] = wall_seq;
    ctx->kernel_ms["cluster_replay_top2_ms"] = ms_top2;
    ctx->kernel_ms["cluster_replay_chain_ms"] = ms_chain;
    ctx->kernel_ms["cluster_replay_near_retries"] = (double)near_retries;
    return ASB_OK;
}

// ---- row-sharded walk (one process per GPU, shard g = global rows [offset_g, offset_g + n_g) in rank order) --
// The walk is order dependent, so the K x F state still travels down the ranks -- but only the CHAINS are serial.  Rank 0
// walks the head of its shard and broadcasts that state P as the common snapshot; every later rank ranks ALL its rows
// against P (nearest / runner-up: the expensive contraction) while rank 0 is still finishing its shard; when the state
// C arrives from the rank before, the chains start from C with the displacement |C - P| already on their books, and the
// certification holds every row to the same proof as on one GPU.  A shard that cannot be proven against the stale
// snapshot (or that receives a state still opening centroids) falls back to the single-GPU algorithm from the
// received state -- the bits are the walk's either way, as asb_cluster_incremental_resume already guarantees.
// The packed state is [x : int64][sizes : uint64 x max_clusters][centroids : f64 x max_clusters x f].
namespace {
size_t shard_state_bytes(int64_t max_clusters, int64_t f) { return 8 + (size_t)max_clusters * 8 + (size_t)max_clusters * f * 8; }

int shard_pack(asb_ctx *ctx, unsigned char *pack_d, int64_t x, const double *cent_d, const unsigned long long *sizes_d,
               int64_t max_clusters, int64_t f) {
    ASB_CUDA(ctx, cudaMemcpyAsync(pack_d, &x, 8, cudaMemcpyHostToDevice, ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(pack_d + 8, sizes_d, (size_t)max_clusters * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(pack_d + 8 + (size_t)max_clusters * 8, cent_d, (size_t)max_clusters * f * 8,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // &x is a stack address
    return ASB_OK;
}
int shard_unpack(asb_ctx *ctx, const unsigned char *pack_d, int64_t *x, double *cent_d, unsigned long long *sizes_d,
                 int64_t max_clusters, int64_t f) {
    ASB_CUDA(ctx, cudaMemcpyAsync(x, pack_d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (sizes_d)
        ASB_CUDA(ctx, cudaMemcpyAsync(sizes_d, pack_d + 8, (size_t)max_clusters * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ASB_CUDA(ctx, cudaMemcpyAsync(cent_d, pack_d + 8 + (size_t)max_clusters * 8, (size_t)max_clusters * f * 8,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASB_OK;
}
}  // namespace

int asb_dev_cluster_sharded(asb_ctx *ctx, asb_comm *comm, const double *rows_d, int64_t n_local, int64_t f,
                            int64_t max_clusters, double radius, double *centroids_d, int64_t *assign_d,
                            unsigned long long *sizes_d, int64_t *x_out_host) {
    if (!comm || comm->nranks == 1)
        return asb_dev_cluster(ctx, rows_d, n_local, f, max_clusters, radius, centroids_d, assign_d, sizes_d, x_out_host, 0);
    const int R = comm->nranks, r = comm->rank;
    if (f <= 0 || max_clusters <= 0 || n_local < 0) ASB_FAIL(ctx, ASB_ERR_INVALID, "cluster_sharded: bad sizes");
    const size_t sb = shard_state_bytes(max_clusters, f);
    DevTmp<unsigned char> pack;
    DevTmp<double> snap;
    ASB_TRY(pack.init(ctx, sb));
    ASB_TRY(snap.init(ctx, (size_t)max_clusters * f));
    ctx->kernel_ms["cluster_shard_speculative"] = 0.0;   // pieces of this shard proven against the common snapshot
    ctx->kernel_ms["cluster_shard_fallback"] = 0.0;      // pieces that were not, and were re-ranked from the fresh state
    int64_t x = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *key) {   // host wall time since the call began (the phases below end in stream syncs)
        ASB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->kernel_ms[key] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return (int)ASB_OK;
    };
    if (r == 0) {
        int64_t head = (int64_t)opt_or(ctx, "cluster_shard_snapshot_rows", 262144.0);
        if (head > n_local) head = n_local;
        if (head > 0)
            ASB_TRY(asb_dev_cluster(ctx, rows_d, head, f, max_clusters, radius, centroids_d, assign_d, sizes_d, &x, 0));
        ASB_TRY(shard_pack(ctx, pack.ptr, x, centroids_d, sizes_d, max_clusters, f));
        ASB_TRY(asb_comm_bcast_bytes(ctx, comm, pack.ptr, sb, 0));
        ASB_TRY(lap("shard_t_snapshot_ms"));
        if (n_local > head) {
            const int64_t x_before = x;
            ASB_TRY(asb_dev_cluster(ctx, rows_d + head * f, n_local - head, f, max_clusters, radius, centroids_d,
                                    assign_d + head, sizes_d, &x, x_before));
        }
    } else {
        ASB_TRY(asb_comm_bcast_bytes(ctx, comm, pack.ptr, sb, 0));
        int64_t x_snap = 0;
        ASB_TRY(shard_unpack(ctx, pack.ptr, &x_snap, snap.ptr, nullptr, max_clusters, f));
        // speculative ranking against the common snapshot, in parallel with the ranks still walking.  The shard is cut
        // into pieces that are certified one by one, so a piece with an uncertifiable row (two centroids sharing a blob
        // leave rows on their bisector) costs a walk of that piece, not of the shard.
        const bool speculate = opt_or(ctx, "cluster_replay", 1.0) != 0.0 && opt_or(ctx, "cluster_shard_speculate", 1.0) != 0.0 &&
                               x_snap == max_clusters && x_snap >= 2 && n_local >= 1024 && max_clusters * f <= (1ll << 27);
        int64_t piece = (int64_t)opt_or(ctx, "cluster_shard_piece", 262144.0);
        if (piece < 4096) piece = 4096;
        if (piece > (1 << 24)) piece = 1 << 24;
        const int64_t npieces = speculate ? (n_local + piece - 1) / piece : 0;
        std::vector<ReplayWs> ws((size_t)npieces);
        for (int64_t i = 0; i < npieces; ++i) {
            const int64_t lo = i * piece, m = std::min<int64_t>(piece, n_local - lo);
            ASB_TRY(replay_ws_init(ctx, ws[(size_t)i], (int)m, (int)max_clusters, (int)f));
            ASB_TRY(replay_prepare(ctx, ws[(size_t)i], rows_d + lo * f, (int)m, (int)f, (int)x_snap, snap.ptr));
        }
        ASB_TRY(lap("shard_t_ranked_ms"));
        ASB_TRY(asb_comm_recv_bytes(ctx, comm, pack.ptr, sb, r - 1));
        ASB_TRY(shard_unpack(ctx, pack.ptr, &x, centroids_d, sizes_d, max_clusters, f));
        ASB_TRY(lap("shard_t_state_in_ms"));
        int64_t proven = 0, walked = 0;
        if (npieces == 0 && n_local > 0) {
            const int64_t x_before = x;
            ASB_TRY(asb_dev_cluster(ctx, rows_d, n_local, f, max_clusters, radius, centroids_d, assign_d, sizes_d, &x, x_before));
            walked = 1;
        }
        for (int64_t i = 0; i < npieces; ++i) {
            const int64_t lo = i * piece, m = std::min<int64_t>(piece, n_local - lo);
            ReplayWs &w = ws[(size_t)i];
            int ok = 0;
            if (x == x_snap) {
                // hint for the thin-margin test: what the piece before measured (drift since the snapshot included), or
                // the drift at the start of this one
                w.last_disp = (i > 0 && ws[(size_t)i - 1].last_disp < INFINITY && ws[(size_t)i - 1].last_disp >= 0.0)
                                  ? ws[(size_t)i - 1].last_disp : -1.0;
                ASB_TRY(replay_run(ctx, w, rows_d + lo * f, (int)m, (int)f, (int)x, 1, radius, snap.ptr, centroids_d,
                                   assign_d + lo, sizes_d, &ok));
                if (!ok && w.use_near && w.last_flags == 2) {   // the tile's bounds were too wide: exact top-2 of the snapshot
                    ASB_TRY(replay_prepare(ctx, w, rows_d + lo * f, (int)m, (int)f, (int)x_snap, snap.ptr, false));
                    w.last_disp = -1.0;
                    ASB_TRY(replay_run(ctx, w, rows_d + lo * f, (int)m, (int)f, (int)x, 1, radius, snap.ptr, centroids_d,
                                       assign_d + lo, sizes_d, &ok));
                }
            }
            if (ok) {
                ++proven;
            } else {
                const int64_t x_before = x;
                ASB_TRY(asb_dev_cluster(ctx, rows_d + lo * f, m, f, max_clusters, radius, centroids_d, assign_d + lo, sizes_d,
                                        &x, x_before));
                ++walked;
            }
        }
        ctx->kernel_ms["cluster_shard_speculative"] = (double)proven;
        ctx->kernel_ms["cluster_shard_fallback"] = (double)walked;
    }
    ASB_TRY(lap("shard_t_walked_ms"));
    ASB_TRY(shard_pack(ctx, pack.ptr, x, centroids_d, sizes_d, max_clusters, f));
    if (r < R - 1) ASB_TRY(asb_comm_send_bytes(ctx, comm, pack.ptr, sb, r + 1));
    ASB_TRY(asb_comm_bcast_bytes(ctx, comm, pack.ptr, sb, R - 1));
    ASB_TRY(shard_unpack(ctx, pack.ptr, &x, centroids_d, sizes_d, max_clusters, f));
    ASB_TRY(lap("shard_t_done_ms"));
    *x_out_host = x;
    return ASB_OK;
}
