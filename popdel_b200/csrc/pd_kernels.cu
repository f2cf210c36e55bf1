// pd_kernels.cu -- sm_100a kernels of the `popdel call` window scan and their orchestration.
//
//   k_screen      (K1, HBM-bound)  every (sample, window): streams the packed read-pair words once with 128-bit
//                                   loads (one warp = 8 tiles = 256 windows of one sample), and only where a read pair
//                                   above the smallest initial-length threshold can be active counts, per window, the
//                                   active read pairs and those above the threshold; flags windows in which some
//                                   sample's upper-half median CAN exceed the threshold (exact necessary condition
//                                   for initialize_deletion_lengths to return a candidate, reference
//                                   genotype_deletion_popdel_call.h:33-87).
//   k_gather      (K2a)            flagged windows only: exact active sets per (window, read group), coverage /
//                                   high-coverage state and the per-sample Q3 (upperHalfMedian, :15-27).
//   k_candidates  (K2b)            per flagged window: sort Q3s over samples, gap-50 clustering, rank-indexed
//                                   thresholds -> candidate initial lengths (:58-86).
//   k_em          (K3-K5)          per (window, initial length): allele-frequency initialisation (:93-133), EM over
//                                   deletion length / reference shifts / allele frequency (:556-664), final genotype
//                                   likelihoods, LAD/DAD/FL, supporting read percentiles, LR test, PL (:665-727).
//   k_compact                      gathers the per-sample rows of the emitted calls into one contiguous buffer.
// No tensor cores: the path is lookup-and-reduce (BASELINE.json north_star).
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "pd_context.h"

#define PD_CUDA(c, call)                                                                          \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return pd_fail((c), PD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int SCREEN_TPW = 8;                                     // tiles per warp in k_screen
constexpr double LN2_D = 0.693147180559945309417232121458;       // the reference evaluates log(2.0) in double
constexpr double LOG10_2_D = 0.301029995663981195213738894724;
// The reference accumulates in long double and subtracts the DOUBLE constants log(2.0) / log10(2.0) from
// logl(ref+del) / log10l(ref+del). For a read pair with ref == del this leaves ln2 - fl(ln2) (resp. the log10
// analogue) per read pair, so three otherwise identical sums are NOT equal there and the "all equal -> assume
// reference" overrides (:246-251, :314-319, :330-335) do not fire. We keep such read pairs out of the double
// sums and re-apply the residue as a tie-break.
constexpr double LN2_RESIDUE = 2.3190468138462996e-17;            // ln 2 - fl64(ln 2)
constexpr double LOG10_2_RESIDUE = -2.8037281277851704e-18;       // log10 2 - fl64(log10 2)
// expl() underflows to 0 below ln(2^-16446): the reference's `res == 0` test on long double (:240, :324)
constexpr double LD_EXP_ZERO = -11399.4985314888605;
constexpr double LN1E10 = -23.025850929940457;                    // ln(1e-10)

enum { CNT_JOBS = 0, CNT_PAIRS = 1, CNT_POOL = 2 };

struct PoolEntry { uint32_t pos_rel; int32_t dev; };
struct PdPair { uint32_t job; int32_t L0; };

__device__ __forceinline__ PdTile load_tile(const PdTile * p)
{
    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
    return PdTile{q.x, q.y, q.z, q.w};
}

// ------------------------------------------------------------------------------------------------------------------
// K1 phase B: lane = window of `tile`; adds to n_g the active read pairs of read group g and to x_g those with
// dev > t_min (stream words of the look-back tiles, then the wide list of long read pairs).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void count_tile(const PdDev & a, uint32_t g, const PdRgConst & k, uint32_t tile, int lane,
                                           uint32_t & n_g, uint32_t & x_g)
{
    const PdTile * tl = a.tiles + (size_t)g * (a.NT + 1);
    const int32_t w0 = (int32_t)(tile * PD_TILE_WINDOWS), w = w0 + lane;
    const uint32_t t_lo = tile > k.lookback_tiles ? tile - k.lookback_tiles : 0;
    for (uint32_t tt = t_lo; tt <= tile; ++tt) {
        const TileSeg ts = tile_seg(tt, a.window_buffer);
        const uint32_t r_lo = tl[tt].off, r_hi = tl[tt + 1].off;
        for (uint32_t base = r_lo; base < r_hi; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t word = i < r_hi ? __ldg(a.words + i) : PD_PAD_WORD;
            int32_t s = 0, e = 0, dev = 0; uint32_t pr;
            bool valid = word_interval(word, ts, k.inner_off, s, e, dev, pr);
            valid = valid && e >= w0 && s <= w0 + 31;
            const uint32_t pk = (uint32_t)e | (dev > a.t_min ? 0x80000000u : 0u);
            uint32_t mask = __ballot_sync(FULL, valid);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const int32_t ss = __shfl_sync(FULL, s, src);
                const uint32_t ee = __shfl_sync(FULL, pk, src);
                const bool hit = ss <= w && w <= (int32_t)(ee & 0x7FFFFFFFu);
                n_g += hit;
                x_g += hit & (ee >> 31);
            }
        }
    }
    const uint32_t l_lo = tl[tile].long_lo, l_hi = tl[tile].long_hi;
    for (uint32_t base = l_lo; base < l_hi; base += 32) {
        const uint32_t i = base + lane;
        PdLong L = i < l_hi ? a.longs[i] : PdLong{0xFFFFFFFFu, 0, 0, 0};
        const bool valid = i < l_hi && (int64_t)L.e >= w0;
        const uint32_t pk = L.e | (L.dev > a.t_min ? 0x80000000u : 0u);
        uint32_t mask = __ballot_sync(FULL, valid);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int32_t ss = (int32_t)__shfl_sync(FULL, L.s, src);
            const uint32_t ee = __shfl_sync(FULL, pk, src);
            const bool hit = ss <= w && w <= (int32_t)(ee & 0x7FFFFFFFu);
            n_g += hit;
            x_g += hit & (ee >> 31);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K1: screen. One warp = SCREEN_TPW consecutive tiles of one sample.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_screen(PdDev a, uint32_t tile_begin, uint32_t tile_end, uint32_t * __restrict__ flags,
                                                uint32_t * __restrict__ jobs, uint32_t * __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const uint32_t T0 = tile_begin + (blockIdx.x * 8 + (threadIdx.x >> 5)) * SCREEN_TPW;
    if (T0 >= tile_end) return;
    const uint32_t T1 = min(T0 + SCREEN_TPW, tile_end);          // exclusive
    const uint32_t smp = blockIdx.y;
    const uint32_t g0 = a.sample_rg[smp], g1 = a.sample_rg[smp + 1];
    const int32_t thr = (int32_t)(((uint32_t)a.t_min << 11) | 0x7FFu);      // (int)word > thr  <=>  dev > t_min

    // ---- phase A: stream every word once; which of my tiles can see a read pair above the threshold?
    uint32_t tmask = 0;                                           // bit i: tile T0+i needs exact counts
    for (uint32_t g = g0; g < g1; ++g) {
        const PdTile * tl = a.tiles + (size_t)g * (a.NT + 1);
        const uint32_t lb = a.rgc[g].lookback_tiles;              // <= PD_MAX_LOOKBACK_TILES
        const uint32_t t_lo = T0 > lb ? T0 - lb : 0;
        // lanes hold the tile table entries t_lo .. T1 (at most 8 + 8 + 1 = 17)
        const uint32_t te = t_lo + lane;
        const bool have = te <= T1;
        PdTile my = PdTile{0xFFFFFFFFu, 0, 0, 0};
        if (have) my = load_tile(tl + te);
        const uint32_t r0 = __shfl_sync(FULL, my.off, 0), r1 = __shfl_sync(FULL, my.off, (int)(T1 - t_lo));
        // long read pairs: any of my tiles with a non-empty wide-list range
        const uint32_t lmask = __ballot_sync(FULL, have && te >= T0 && te < T1 && my.long_lo < my.long_hi);
        tmask |= (lmask >> (T0 - t_lo)) & ((1u << SCREEN_TPW) - 1u);
        for (uint32_t i0 = r0; i0 < r1; i0 += 128 * 4) {
            bool ex[4]; uint32_t idx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                idx[u] = i0 + u * 128 + lane * 4;
                ex[u] = false;
                if (idx[u] < r1) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.words + idx[u]));
                    ex[u] = ((int32_t)v.x > thr) | ((int32_t)v.y > thr) | ((int32_t)v.z > thr) | ((int32_t)v.w > thr);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t m = __ballot_sync(FULL, ex[u]);
                while (m) {                                       // rare: find the tile of the exceeding word
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t wi = __shfl_sync(FULL, idx[u], src);
                    // tile index = (number of table entries with off <= wi) - 1
                    const uint32_t nle = __popc(__ballot_sync(FULL, have && my.off <= wi));
                    const int32_t rel = (int32_t)(t_lo + nle - 1) - (int32_t)T0;     // negative: look-back tile
                    const int32_t lo = rel < 0 ? 0 : rel, hi = min(rel + (int32_t)lb, SCREEN_TPW - 1);
                    if (hi >= lo) tmask |= ((1u << (hi + 1)) - 1u) & ~((1u << lo) - 1u);
                }
            }
        }
    }
    if (tmask == 0) return;

    // ---- phase B: exact counts per window (lane) for the tiles in tmask
    for (uint32_t ti = 0; ti < (uint32_t)SCREEN_TPW; ++ti) {
        if (!((tmask >> ti) & 1u)) continue;
        const uint32_t tile = T0 + ti;
        if (tile >= T1) break;
        uint32_t cov = 0, n = 0, x = 0;
        for (uint32_t g = g0; g < g1; ++g) {
            const PdRgConst k = a.rgc[g];
            uint32_t n_g = 0, x_g = 0;
            count_tile(a, g, k, tile, lane, n_g, x_g);
            cov += n_g;
            if (n_g < k.max_load) { n += n_g; x += x_g; }
        }
        const uint32_t w = tile * PD_TILE_WINDOWS + lane;
        const bool in_range = w >= a.w_begin && w < a.w_end;
        const bool pass = in_range && cov >= 2 && n >= 1 && x >= pd_q3_need(n);
        if (pass) {
            const uint32_t wi = w - a.w_begin;
            if (flags[wi] == 0 && atomicExch(&flags[wi], 1u) == 0) {
                uint32_t slot = atomicAdd(&counters[CNT_JOBS], 1u);
                jobs[slot] = w;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K2a: gather. One warp = one (flagged window, sample).
// ------------------------------------------------------------------------------------------------------------------
struct GatherArgs {
    const uint32_t * jobs; uint32_t job0, njobs;
    PoolEntry * pool; uint32_t pool_cap;
    uint32_t * counters;
    uint32_t * act_off, * act_cnt;       // [njobs][R]
    int32_t * q3; uint8_t * sstat;       // [njobs][N]   sstat: 0 = low coverage, 1 = no usable values, 2 = Q3 valid
};

// calls f(valid, pos_rel, dev) for every 32-wide batch of read pairs of read group g that may be active at window w;
// `valid` marks lanes whose read pair IS active at w. Batches preserve stream order (stream first, then wide list).
template <typename F>
__device__ __forceinline__ void for_active_batches(const PdDev & a, uint32_t g, const PdRgConst & k, int32_t w, int lane, F f)
{
    const PdTile * tl = a.tiles + (size_t)g * (a.NT + 1);
    const uint32_t tile = (uint32_t)w / PD_TILE_WINDOWS;
    const uint32_t t_lo = tile > k.lookback_tiles ? tile - k.lookback_tiles : 0;
    for (uint32_t tt = t_lo; tt <= tile; ++tt) {
        const TileSeg ts = tile_seg(tt, a.window_buffer);
        const uint32_t r_lo = tl[tt].off, r_hi = tl[tt + 1].off;
        for (uint32_t base = r_lo; base < r_hi; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t word = i < r_hi ? __ldg(a.words + i) : PD_PAD_WORD;
            int32_t s = 0, e = 0, dev = 0; uint32_t pr = 0;
            bool valid = word_interval(word, ts, k.inner_off, s, e, dev, pr);
            valid = valid && s <= w && w <= e;
            f(valid, pr, dev);
        }
    }
    const uint32_t l_lo = tl[tile].long_lo, l_hi = tl[tile].long_hi;
    for (uint32_t base = l_lo; base < l_hi; base += 32) {
        const uint32_t i = base + lane;
        PdLong L = i < l_hi ? a.longs[i] : PdLong{0xFFFFFFFFu, 0, 0, 0};
        const bool valid = i < l_hi && (int64_t)L.s <= w && (int64_t)L.e >= w;
        f(valid, L.pos_rel, L.dev);
    }
}

constexpr int GATHER_STAGE = 96;                                   // staged read pairs per warp (shared memory)

__global__ void __launch_bounds__(128) k_gather(PdDev a, GatherArgs ga)
{
    __shared__ PoolEntry stage[4][GATHER_STAGE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t job = blockIdx.x;
    const uint32_t smp = blockIdx.y * 4 + wib;
    if (smp >= a.N) return;
    const int32_t w = (int32_t)ga.jobs[ga.job0 + job];
    const uint32_t g0 = a.sample_rg[smp], g1 = a.sample_rg[smp + 1];
    uint32_t * cnt = ga.act_cnt + (size_t)job * a.R;
    uint32_t * off = ga.act_off + (size_t)job * a.R;
    PoolEntry * st = stage[wib];

    // single pass: count per read group and stage the usable read pairs in shared memory (stream order)
    uint32_t cov = 0, nvals = 0;
    bool overflow = false;
    for (uint32_t g = g0; g < g1; ++g) {
        const uint32_t max_load = __ldg(&a.rgc[g].max_load);
        const PdRgConst k = a.rgc[g];
        uint32_t n_g = 0;
        const uint32_t start = nvals;
        for_active_batches(a, g, k, w, lane, [&](bool valid, uint32_t pr, int32_t dev) {
            const uint32_t mask = __ballot_sync(FULL, valid);
            const uint32_t slot = start + n_g + __popc(mask & ((1u << lane) - 1u));
            if (valid && slot < (uint32_t)GATHER_STAGE) st[slot] = PoolEntry{pr, dev};
            n_g += __popc(mask);
        });
        if (lane == 0) { cnt[g] = n_g; off[g] = start; }                 // off = offset inside the sample's segment for now
        cov += n_g;
        if (n_g < max_load) { nvals += n_g; if (nvals > (uint32_t)GATHER_STAGE) overflow = true; }
    }
    uint32_t base = 0;
    if (lane == 0 && nvals) base = atomicAdd(&ga.counters[CNT_POOL], nvals);
    base = __shfl_sync(FULL, base, 0);
    const bool fits = (uint64_t)base + nvals <= ga.pool_cap;
    __syncwarp();
    if (lane == 0) for (uint32_t g = g0; g < g1; ++g) off[g] += base;
    if (fits) {
        if (!overflow) {
            for (uint32_t i = lane; i < nvals; i += 32) ga.pool[base + i] = st[i];
        } else {                                                         // rare: more usable pairs than the stage holds
            uint32_t cur = base;
            for (uint32_t g = g0; g < g1; ++g) {
                const PdRgConst k = a.rgc[g];
                const uint32_t n_g = __shfl_sync(FULL, lane == 0 ? cnt[g] : 0u, 0);
                if (n_g >= k.max_load) continue;
                for_active_batches(a, g, k, w, lane, [&](bool valid, uint32_t pr, int32_t dev) {
                    const uint32_t mask = __ballot_sync(FULL, valid);
                    if (valid) ga.pool[cur + __popc(mask & ((1u << lane) - 1u))] = PoolEntry{pr, dev};
                    cur += __popc(mask);
                });
            }
        }
    }
    __syncwarp();
    // coverage state and Q3 (upperHalfMedian :15-27: n<4 -> maximum; else interpolate at (3n+2+n%2)/4-1)
    uint8_t stt; int32_t q = 0;
    if (cov < 2u) stt = 0;
    else if (nvals == 0 || !fits) stt = 1;
    else {
        stt = 2;
        const uint32_t nn = nvals;
        uint32_t l; double r = 0;
        if (nn < 4) { l = nn - 1; }
        else { double pos = (3.0 * nn + 2.0 + (nn % 2)) / 4.0 - 1.0; l = (uint32_t)pos; r = pos - l; }
        const uint32_t l2 = (l + 1 < nn) ? l + 1 : l;
        int32_t lo_v = 0, hi_v = 0;
        if (nn <= 32) {
            const int32_t myval = (uint32_t)lane < nn ? st[lane].dev : INT_MAX;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < nn; ++j) {
                const int32_t vj = __shfl_sync(FULL, myval, (int)j);
                rank += (vj < myval) || (vj == myval && (int)j < lane);
            }
            const uint32_t m1 = __ballot_sync(FULL, (uint32_t)lane < nn && rank == l);
            const uint32_t m2 = __ballot_sync(FULL, (uint32_t)lane < nn && rank == l2);
            lo_v = __shfl_sync(FULL, myval, __ffs(m1) - 1);
            hi_v = __shfl_sync(FULL, myval, __ffs(m2) - 1);
        } else {
            const volatile PoolEntry * v = ga.pool + base;              // written above by this warp
            for (uint32_t i0 = 0; i0 < nn; i0 += 32) {
                const uint32_t i = i0 + lane;
                const int32_t vi = i < nn ? v[i].dev : INT_MAX;
                uint32_t rank = 0;
                for (uint32_t j = 0; j < nn; ++j) { const int32_t vj = v[j].dev; rank += (vj < vi) || (vj == vi && j < i); }
                const uint32_t m1 = __ballot_sync(FULL, i < nn && rank == l);
                const uint32_t m2 = __ballot_sync(FULL, i < nn && rank == l2);
                if (m1) lo_v = __shfl_sync(FULL, vi, __ffs(m1) - 1);
                if (m2) hi_v = __shfl_sync(FULL, vi, __ffs(m2) - 1);
            }
        }
        if (nn < 4) q = (int32_t)floor((double)lo_v + 0.5);
        else q = (int32_t)floor((1 - r) * lo_v + r * hi_v + 0.5);
    }
    if (lane == 0) { ga.q3[(size_t)job * a.N + smp] = q; ga.sstat[(size_t)job * a.N + smp] = stt; }
}

// ------------------------------------------------------------------------------------------------------------------
// K2b: candidates. One block per flagged window.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_candidates(PdDev a, const int32_t * __restrict__ q3, const uint8_t * __restrict__ sstat,
                                                    uint32_t job0, uint32_t * counters, PdPair * pairs, uint32_t pair_cap, uint32_t npad)
{
    extern __shared__ int32_t sv[];
    __shared__ uint32_t s_n;
    const uint32_t job = blockIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) sv[i] = INT_MAX;
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < a.N; s += blockDim.x)
        if (sstat[(size_t)job * a.N + s] == 2) sv[atomicAdd(&s_n, 1u)] = q3[(size_t)job * a.N + s];
    __syncthreads();
    const uint32_t nv = s_n;
    if (nv == 0) return;
    uint32_t np2 = 1; while (np2 < nv) np2 <<= 1;                 // sort only the occupied power of two
    for (uint32_t k = 2; k <= np2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    int32_t x = sv[i], y = sv[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { sv[i] = y; sv[ixj] = x; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        // genotype_deletion_popdel_call.h:62-84; thresholds are indexed by RANK in the sorted array (quirk)
        int sum = sv[0], n = 1;
        uint32_t thr = a.rgc[0].min_init;
        auto emit = [&](int mean) {
            uint32_t slot = atomicAdd(&counters[CNT_PAIRS], 1u);
            if (slot < pair_cap) pairs[slot] = PdPair{job0 + job, mean};
        };
        for (uint32_t i = 1; i < nv; ++i) {
            if (sv[i - 1] + 50 > sv[i]) { sum += sv[i]; ++n; thr = min(thr, a.rgc[i].min_init); }
            else { if (sum / n > (int)thr) emit(sum / n); sum = sv[i]; n = 1; thr = a.rgc[i].min_init; }
        }
        if (sum / n > (int)thr) emit(sum / n);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K3-K5: EM + final pass. One block per (flagged window, initial length); thread = sample (strided).
// ------------------------------------------------------------------------------------------------------------------
struct EmArgs {
    const uint32_t * jobs; const PdPair * pairs; uint32_t pair0, npairs, job_base;   // pairs[].job is absolute; scratch uses job - job_base
    const PoolEntry * pool; const uint32_t * act_off, * act_cnt; const uint8_t * sstat;
    double * dlx;            // [block][N][3]   data likelihoods, log domain, max = 0
    int32_t * shifts;        // [block][R]
    uint32_t * ps;           // [block][N][13]  per-sample output rows (pair-indexed)
    pd_call * calls;         // [block]         call headers (pair-indexed)
    uint8_t * valid;         // [block]
    uint32_t iterations, min_len; double min_lr, min_sample_fraction; int somatic, window_wise; uint32_t anchor;
    uint32_t * dbg;          // optional [npairs][4]: reason, len, iterations, supp (PD_DEBUG)
    int dbg_window;          // device printf of the EM trajectory of this window (PD_DEBUG_WINDOW), -1 = off
};

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// deterministic block sum of two doubles (fixed shuffle tree, then warps in index order); result on all threads
__device__ __forceinline__ void block_sum2(double & a, double & b, double * red)
{
    a = warp_sum(a); b = warp_sum(b);
    const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[2 * wid] = a; red[2 * wid + 1] = b; }
    __syncthreads();
    double sa = 0, sb = 0;
    for (int i = 0; i < nw; ++i) { sa += red[2 * i]; sb += red[2 * i + 1]; }
    a = sa; b = sb;
}
__device__ __forceinline__ void block_sum2u(unsigned long long & a, unsigned long long & b, unsigned long long * red)
{
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(FULL, a, o); b += __shfl_xor_sync(FULL, b, o); }
    const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[2 * wid] = a; red[2 * wid + 1] = b; }
    __syncthreads();
    unsigned long long sa = 0, sb = 0;
    for (int i = 0; i < nw; ++i) { sa += red[2 * i]; sb += red[2 * i + 1]; }
    a = sa; b = sb;
}

// table entry of insert-size deviation `dev` (I(), insert_histogram_popdel.h:1157-1163): floor entry outside the table
__device__ __forceinline__ const PdTab * tab_at(const PdTab * __restrict__ tab, const PdRgConst & k, int dev)
{
    const int i = dev + k.hist_base;
    const bool in = i > 0 && i + 1 < (int)k.hist_len;
    return tab + k.hist_off + (in ? i + 1 : 0);
}

struct Gt { double a, b, c; };
__device__ __forceinline__ Gt gt_prior(double f, int somatic)       // :343-380
{
    const double ps = 0.0000000001;
    Gt g;
    if (!somatic) { g.a = fmax((1 - f) * (1 - f), ps); g.b = fmax(2 * f * (1 - f), ps); g.c = fmax(f * f, ps); }
    else if (f <= 0.4) { g.a = fmax(1 - 2 * f + ps, ps); g.b = fmax(2 * f - 2 * ps, ps); g.c = ps; }
    else if (f < 0.75) { g.a = ps; g.b = 1.; g.c = ps; }
    else { g.a = ps; g.b = ps; g.c = 1.; }
    return g;
}

struct EmShared {
    double red[64];
    unsigned long long redu[64];
    double rgw[3];                     // log-domain per-RG likelihoods of read group 0 (quirk: drives all reference shifts)
    double freq, prevFreq;
    Gt gt, prevGt;
    double ea0Rg, ea1Rg;
    uint32_t len, prevLen, it;
    int visited_len[64]; double visited_freq[64]; int nvisited;
    int stop;
};

// Normalises three log-likelihood sums like the reference (:235-251): subtract the maximum, apply the long-double
// tie-break of `ndeg` read pairs with ref == del to the heterozygous sum, then the two overrides.
__device__ __forceinline__ void finish_triple(double l0, double l1, double l2, uint32_t ndeg, double & x0, double & x1, double & x2)
{
    const double m = fmax(fmax(l0, l1), l2);
    x0 = l0 - m; x1 = l1 - m; x2 = l2 - m;
    if (ndeg) {
        x1 += ndeg * LN2_RESIDUE;
        const double m2 = fmax(fmax(x0, x1), x2);
        x0 -= m2; x1 -= m2; x2 -= m2;
    }
    if (x0 < LD_EXP_ZERO || x1 < LD_EXP_ZERO || x2 < LD_EXP_ZERO) { x0 = 0; x1 = LN1E10; x2 = LN1E10; }
    if (x0 == x1 && x0 == x2) { x0 = 0; x1 = LN1E10; x2 = LN1E10; }
}

struct RgLite { int hist_base; uint32_t hist_len, hist_off, max_load; double min_prob; };
__device__ __forceinline__ RgLite rg_lite(const PdRgConst * r)
{
    RgLite k;
    k.hist_base = __ldg(&r->hist_base); k.hist_len = __ldg(&r->hist_len); k.hist_off = __ldg(&r->hist_off);
    k.max_load = __ldg(&r->max_load); k.min_prob = __ldg(&r->min_prob);
    return k;
}
__device__ __forceinline__ const PdTab * tab_at(const PdTab * __restrict__ tab, const RgLite & k, int dev)
{
    const int i = dev + k.hist_base;
    const bool in = i > 0 && i + 1 < (int)k.hist_len;
    return tab + k.hist_off + (in ? i + 1 : 0);
}

// EM state handed from k_em to k_final
struct EmState { uint32_t len, it, alive, pad; double freq; double gt[3]; };

// ------------------------------------------------------------------------------------------------------------------
// EM kernel: LPS lanes cooperate on one sample (read pairs strided over the lanes, fixed-order shuffle reduction);
// the two table values of every read pair computed by the data-likelihood pass are cached in shared memory for the
// length update of the next iteration (same deletion length and reference shifts -> identical look-ups).
// ------------------------------------------------------------------------------------------------------------------
constexpr int EM_CACHE_SLOTS = 8;

template <int LPS>
__device__ __forceinline__ double group_sum(double v, uint32_t gmask)
{
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}

// data likelihoods of every sample for (L, shifts) -> dlx; compute_data_likelihoods (EM overload) :179-253
// Returns this thread's part of update_allele_frequency's sum (:467-485) under the genotype priors `gtf`.
template <int LPS>
__device__ double compute_dl(const PdDev & a, const EmArgs & e, EmShared & sh, const uint32_t * cnt, const uint32_t * off,
                             double * dlx, const int32_t * shifts, bool zero_shifts, uint32_t L, double * cache, bool fill_cache,
                             const Gt gtf)
{
    double fs = 0;
    const int tid = threadIdx.x, sub = tid % LPS, grp = tid / LPS, ngrp = blockDim.x / LPS;
    const uint32_t gmask = (LPS == 32) ? FULL : (((1u << LPS) - 1u) << ((tid & 31) & ~(LPS - 1)));
    for (uint32_t s = grp; s < a.N; s += ngrp) {
        double l0 = 0, l1 = 0, l2 = 0, w0 = 0, w1 = 0, w2 = 0, nd = 0;
        int slot = 0;
        bool high0 = false;
        for (uint32_t g = a.sample_rg[s]; g < a.sample_rg[s + 1]; ++g) {
            const RgLite k = rg_lite(a.rgc + g);
            const uint32_t n = cnt[g];
            if (n >= k.max_load) { if (g == 0) high0 = true; continue; }
            const int shift = zero_shifts ? 0 : shifts[g];
            const PoolEntry * p = e.pool + off[g];
            for (uint32_t i = sub; i < n; i += LPS, ++slot) {
                const int d = __ldg(&p[i].dev);
                const PdTab * tr = tab_at(a.tab, k, d - shift);
                const PdTab * td = tab_at(a.tab, k, d - (int)L);
                const double ref = __ldg(&tr->val), del = __ldg(&td->val);
                const double g0 = __ldg(&tr->ln), g2 = __ldg(&td->ln);
                double g1;
                if (ref == del) { g1 = g0; nd += 1; }                    // + LN2_RESIDUE, applied in finish_triple
                else if (del == k.min_prob) g1 = __ldg(&tr->lnp);        // ln(ref + min_prob) - ln2_d
                else if (ref == k.min_prob) g1 = __ldg(&td->lnp);
                else g1 = log(ref + del) - LN2_D;
                if (fill_cache && slot < EM_CACHE_SLOTS) { cache[(2 * slot) * blockDim.x + tid] = del; cache[(2 * slot + 1) * blockDim.x + tid] = ref; }
                l0 += g0; l1 += g1; l2 += g2;
                if (g == 0) { w0 += g0; w1 += g1; w2 += g2; }
            }
        }
        l0 = group_sum<LPS>(l0, gmask); l1 = group_sum<LPS>(l1, gmask); l2 = group_sum<LPS>(l2, gmask); nd = group_sum<LPS>(nd, gmask);
        if (s == a.rgc[0].sample) {                                      // read group 0 belongs to this sample
            w0 = group_sum<LPS>(w0, gmask); w1 = group_sum<LPS>(w1, gmask); w2 = group_sum<LPS>(w2, gmask);
            if (sub == 0) {
                if (high0) { sh.rgw[0] = sh.rgw[1] = sh.rgw[2] = -INFINITY; }        // Triple(0,0,0) in the reference
                else { const double m = fmax(fmax(w0, w1), w2); sh.rgw[0] = w0 - m; sh.rgw[1] = w1 - m; sh.rgw[2] = w2 - m; }
            }
        }
        if (sub == 0) {
            double x0, x1, x2;
            finish_triple(l0, l1, l2, (uint32_t)nd, x0, x1, x2);
            dlx[3 * s] = x0; dlx[3 * s + 1] = x1; dlx[3 * s + 2] = x2;
            const double p0 = exp(x0) * gtf.a, p1 = exp(x1) * gtf.b, p2 = exp(x2) * gtf.c;
            fs += (p1 + 2 * p2) / (p0 + p1 + p2);
        }
    }
    __syncthreads();
    return fs;
}

// deletion_likelihood_ratio :490-508 (block-wide; result on all threads)
__device__ double block_lr(const PdDev & a, EmShared & sh, const double * dlx, const Gt gt)
{
    double del = 0, nodel = 0;
    for (uint32_t s = threadIdx.x; s < a.N; s += blockDim.x) {
        const double x0 = dlx[3 * s], A = exp(x0), B = exp(dlx[3 * s + 1]), C = exp(dlx[3 * s + 2]);
        const double p0 = A * gt.a, p1 = B * gt.b, p2 = C * gt.c, pAll = p0 + p1 + p2;
        const double a0 = p0 / pAll, a1 = p1 / pAll, a2 = p2 / pAll;
        del += log(a0 * A + a1 * B + a2 * C);
        nodel += x0;
    }
    block_sum2(del, nodel, sh.red);
    return del - nodel;
}

template <int LPS>
__global__ void __launch_bounds__(LPS >= 8 ? 1024 : 512, LPS >= 8 ? 1 : 2) k_em(PdDev a, EmArgs e, EmState * states)
{
    __shared__ EmShared sh;
    extern __shared__ double cache[];                 // [2 * EM_CACHE_SLOTS][blockDim.x]
    const uint32_t pi = e.pair0 + blockIdx.x;
    const PdPair pr = e.pairs[pi];
    const uint32_t job = pr.job - e.job_base;
    const uint32_t L0 = (uint32_t)pr.L0;
    const uint32_t w = e.jobs[pr.job];
    const uint32_t * cnt = e.act_cnt + (size_t)job * a.R;
    const uint32_t * off = e.act_off + (size_t)job * a.R;
    double * dlx = e.dlx + (size_t)blockIdx.x * 3 * a.N;
    int32_t * shifts = e.shifts + (size_t)blockIdx.x * a.R;
    const int tid = threadIdx.x, sub = tid % LPS, grp = tid / LPS, ngrp = blockDim.x / LPS;
    const uint32_t gmask = (LPS == 32) ? FULL : (((1u << LPS) - 1u) << ((tid & 31) & ~(LPS - 1)));
    auto finish = [&](uint32_t alive, uint32_t reason) {
        if (tid == 0) {
            EmState st; st.len = sh.len; st.it = sh.it; st.alive = alive; st.pad = 0; st.freq = sh.freq;
            st.gt[0] = sh.gt.a; st.gt[1] = sh.gt.b; st.gt[2] = sh.gt.c;
            states[blockIdx.x] = st;
            if (!alive) {
                e.valid[blockIdx.x] = 0;
                if (e.dbg) { e.dbg[4 * blockIdx.x] = reason; e.dbg[4 * blockIdx.x + 1] = sh.len; e.dbg[4 * blockIdx.x + 2] = sh.it; }
            }
        }
    };

    for (uint32_t g = tid; g < a.R; g += blockDim.x) shifts[g] = 0;

    // ---- initialize_allele_frequency :93-133
    {
        unsigned long long c = 0, t = 0;
        for (uint32_t g = tid; g < a.R; g += blockDim.x) {
            const uint32_t n = cnt[g];
            if (n == 0 || n >= __ldg(&a.rgc[g].max_load)) continue;
            t += n;
            const double sd = __ldg(&a.rgc[g].stddev);
            const int wb = max((int)L0 / 2, (int)floor((double)L0 - 2 * sd + 0.5));
            const int we = (int)((double)L0 + 2 * sd);
            const PoolEntry * p = e.pool + off[g];
            for (uint32_t i = 0; i < n; ++i) { const int d = p[i].dev; c += (d > wb && d < we); }
        }
        block_sum2u(c, t, sh.redu);
        if (tid == 0) {
            sh.freq = t == 0 ? 0.0 : (double)c / (double)t;
            sh.len = L0; sh.it = 0; sh.nvisited = 0; sh.stop = 0;
            sh.gt = gt_prior(sh.freq, e.somatic);
        }
        __syncthreads();
    }
    if (sh.freq == 0) { finish(0, 1); return; }
    compute_dl<LPS>(a, e, sh, cnt, off, dlx, shifts, false, L0, cache, true, sh.gt);
    if (tid == 0) { sh.prevFreq = sh.freq; sh.prevLen = sh.len; sh.prevGt = sh.gt; }
    __syncthreads();

    // ---- EM loop :598-660
    while (true) {
        const bool cont = sh.len >= e.min_len && sh.it < e.iterations;
        __syncthreads();                      // everyone has evaluated the loop condition before thread 0 touches it
        if (!cont) break;
        if (tid == 0) {
            ++sh.it;
            int key = (int)sh.len, f = -1;
            for (int i = 0; i < sh.nvisited; ++i) if (sh.visited_len[i] == key) f = i;
            if (f < 0) { f = sh.nvisited++; sh.visited_len[f] = key; }
            sh.visited_freq[f] = sh.freq;
            sh.prevLen = sh.len; sh.prevFreq = sh.freq; sh.prevGt = sh.gt;
            // weights of read group 0 (rgDlIt is never advanced, :401,424-431)
            const double r0 = exp(sh.rgw[0]), r1 = exp(sh.rgw[1]), r2 = exp(sh.rgw[2]);
            const double aSumRg = r0 * sh.gt.a + r1 * sh.gt.b + r2 * sh.gt.c;
            const double a0Rg = sh.rgw[0] + log(sh.gt.a) - log(aSumRg);
            const double a1Rg = sh.rgw[1] + log(sh.gt.b) - log(aSumRg);
            sh.ea0Rg = exp(a0Rg); sh.ea1Rg = exp(a1Rg);
        }
        __syncthreads();
        // update_deletion_length :388-462 (table values of (len, shifts) come from the cache filled by compute_dl)
        const Gt gt = sh.gt;
        const int L = (int)sh.len;
        const double ea0Rg = sh.ea0Rg, ea1Rg = sh.ea1Rg;
        double sumDel = 0, wDel = 0;
        for (uint32_t s = grp; s < a.N; s += ngrp) {
            const double x1 = dlx[3 * s + 1], x2 = dlx[3 * s + 2];
            const double aSum = exp(dlx[3 * s]) * gt.a + exp(x1) * gt.b + exp(x2) * gt.c;
            const double a1 = x1 + log(gt.b) - log(aSum);
            const double a2 = x2 + log(gt.c) - log(aSum);
            const double ea1 = exp(a1), ea2 = exp(a2);
            int slot = 0;
            for (uint32_t g = a.sample_rg[s]; g < a.sample_rg[s + 1]; ++g) {
                const RgLite k = rg_lite(a.rgc + g);
                const uint32_t n = cnt[g];
                if (n >= k.max_load) continue;
                double sumRef = 0, wRef = 0;
                const int shift = shifts[g];
                const PoolEntry * p = e.pool + off[g];
                for (uint32_t i = sub; i < n; i += LPS, ++slot) {
                    const int d = __ldg(&p[i].dev);
                    double del, nod;
                    if (slot < EM_CACHE_SLOTS) { del = cache[(2 * slot) * blockDim.x + tid]; nod = cache[(2 * slot + 1) * blockDim.x + tid]; }
                    else { del = __ldg(&tab_at(a.tab, k, d - L)->val); nod = __ldg(&tab_at(a.tab, k, d - shift)->val); }
                    const double pd = ea1 * del / (del + nod) + ea2;
                    const double prf = ea1Rg * nod / (del + nod) + ea0Rg;
                    sumDel += pd; sumRef += prf;
                    wDel += pd * d; wRef += prf * d;
                }
                sumRef = group_sum<LPS>(sumRef, gmask); wRef = group_sum<LPS>(wRef, gmask);
                if (sub == 0) {
                    const double q = wRef / sumRef;
                    int sft = (q != q) ? 0 : (q >= 2147483647.0 ? INT_MAX : (q <= -2147483648.0 ? INT_MIN : (int)q));
                    const double sd = __ldg(&a.rgc[g].stddev);
                    if (sft > sd || sft < -1 * sd) sft = 0;
                    shifts[g] = sft;
                }
            }
        }
        block_sum2(sumDel, wDel, sh.red);
        if (tid == 0 && (int)w == e.dbg_window)
            printf("GPU w %u L0 %u it %u len %u freq %.17g sumDel %.17g wDel %.17g\n", w, L0, sh.it, sh.len, sh.freq, sumDel, wDel);
        if (tid == 0) {
            uint32_t nl;
            if (sumDel == 0) nl = 0;
            else { double len = wDel / sumDel; nl = len < 0 ? 0u : (uint32_t)round(len); }
            sh.len = nl;
        }
        __syncthreads();
        // data likelihoods at the new length + update_allele_frequency :467-485 (priors of the previous iteration)
        double fs = compute_dl<LPS>(a, e, sh, cnt, off, dlx, shifts, false, sh.len, cache, true, gt), dummy = 0;
        block_sum2(fs, dummy, sh.red);
        if (tid == 0) {
            sh.freq = fs / 2.0 / a.N;
            sh.stop = 0;
            if (sh.freq == 0) sh.stop = 1;
            else {
                sh.gt = gt_prior(sh.freq, e.somatic);
                int key = (int)sh.len;
                for (int i = 0; i < sh.nvisited; ++i)
                    if (sh.visited_len[i] == key && fabs(sh.visited_freq[i] - sh.freq) <= 0.0001) sh.stop = 2;
            }
        }
        __syncthreads();
        if (sh.stop == 1) break;
        if (sh.stop == 2) {
            // convergence :632-658: compare with the previous estimate evaluated with the initial (zero) shifts
            const double lr = block_lr(a, sh, dlx, sh.gt);
            compute_dl<LPS>(a, e, sh, cnt, off, dlx, shifts, true, sh.prevLen, cache, false, sh.prevGt);
            const double plr = block_lr(a, sh, dlx, sh.prevGt);
            __syncthreads();
            if (plr > lr) {
                if (tid == 0) { sh.len = sh.prevLen; sh.freq = sh.prevFreq; }
                for (uint32_t g = tid; g < a.R; g += blockDim.x) shifts[g] = 0;
            }
            __syncthreads();
            break;
        }
    }
    __syncthreads();
    const bool alive = !(sh.freq < 0.0000000001 || sh.len < e.min_len);
    finish(alive ? 1u : 0u, 2);
}

// ------------------------------------------------------------------------------------------------------------------
// final pass: one block per (window, initial length) that survived the EM; thread = sample (strided)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_final(PdDev a, EmArgs e, const EmState * states)
{
    constexpr uint32_t SUPP_CAP = 1536;               // supporting read pairs kept in shared memory for the percentiles
    __shared__ EmShared sh;
    __shared__ uint32_t s_first[SUPP_CAP], s_last[SUPP_CAP];
    __shared__ uint32_t s_nsupp;
    const EmState stt = states[blockIdx.x];
    if (!stt.alive) return;
    if (threadIdx.x == 0) s_nsupp = 0;
    const uint32_t pi = e.pair0 + blockIdx.x;
    const PdPair pr = e.pairs[pi];
    const uint32_t job = pr.job - e.job_base;
    const uint32_t L0 = (uint32_t)pr.L0;
    const uint32_t w = e.jobs[pr.job];
    const uint32_t * cnt = e.act_cnt + (size_t)job * a.R;
    const uint32_t * off = e.act_off + (size_t)job * a.R;
    const uint8_t * sstat = e.sstat + (size_t)job * a.N;
    double * dlx = e.dlx + (size_t)blockIdx.x * 3 * a.N;
    const int32_t * shifts = e.shifts + (size_t)blockIdx.x * a.R;
    uint32_t * ps = e.ps + (size_t)blockIdx.x * 13 * a.N;
    const int tid = threadIdx.x;
    if (tid == 0) { sh.len = stt.len; sh.it = stt.it; sh.freq = stt.freq; sh.gt = Gt{stt.gt[0], stt.gt[1], stt.gt[2]}; }
    __syncthreads();
    auto reject = [&](uint32_t reason) {
        if (tid == 0) {
            e.valid[blockIdx.x] = 0;
            if (e.dbg) { e.dbg[4 * blockIdx.x] = reason; e.dbg[4 * blockIdx.x + 1] = sh.len; e.dbg[4 * blockIdx.x + 2] = sh.it; }
        }
    };
    // ---- final pass :665-727 (compute_data_likelihoods final overload :255-337)
    const int len = (int)sh.len;
    unsigned long long supp = 0, ndata = 0;
    uint32_t smin = 0xFFFFFFFFu, smax = 0, lmin = 0xFFFFFFFFu, lmax = 0;   // ranges of supporting starts / ends
    for (uint32_t s = tid; s < a.N; s += blockDim.x) {
        uint32_t lad[3] = {0, 0, 0}, dad[5] = {0, 0, 0, 0, 0};
        uint32_t fl_min = 0xFFFFFFFFu, fl_max = 0, ndeg = 0;
        double l0 = 0, l1 = 0, l2 = 0, t0 = 0, t1 = 0, t2 = 0;
        int delLower = INT_MAX, delUpper = 0;
        const uint32_t g0 = a.sample_rg[s], g1 = a.sample_rg[s + 1];
        for (uint32_t g = g0; g < g1; ++g) {
            const PdRgConst k = a.rgc[g];
            const uint32_t n = cnt[g];
            if (n >= k.max_load) continue;
            const int shift = shifts[g];
            delLower = len - k.lower_q; delUpper = len + k.upper_q;
            const PoolEntry * p = e.pool + off[g];
            for (uint32_t i = 0; i < n; ++i) {
                const int d = p[i].dev;
                if (d > k.upper_q) { if (d < delLower) ++dad[2]; else if (d <= delUpper) ++dad[3]; else ++dad[4]; }
                else { if (d < delUpper) ++dad[0]; else ++dad[1]; }
                const PdTab * tr = tab_at(a.tab, k, d - shift);
                const PdTab * td = tab_at(a.tab, k, d - len);
                const double ref = tr->val, del = td->val;
                if (ref >= 2 * del) ++lad[0]; else if (del >= 2 * ref) ++lad[2]; else ++lad[1];
                l0 += tr->ln; t0 += tr->l10;
                l2 += td->ln; t2 += td->l10;
                if (ref == del) { l1 += tr->ln; t1 += tr->l10; ++ndeg; }              // residues applied below
                else if (del == k.min_prob) { l1 += tr->lnp; t1 += tr->l10p; }
                else if (ref == k.min_prob) { l1 += td->lnp; t1 += td->l10p; }
                else { l1 += log(ref + del) - LN2_D; t1 += log10(ref + del) - LOG10_2_D; }
                const uint32_t first = p[i].pos_rel + e.anchor;
                const int isz = max(0, d + k.inner_off);
                const uint32_t last = first + (uint32_t)isz;
                fl_min = min(fl_min, first); fl_max = max(fl_max, last);
            }
        }
        // supporting read pairs: borders of the LAST usable read group apply to all of the sample's read groups (quirk)
        for (uint32_t g = g0; g < g1; ++g) {
            const PdRgConst k = a.rgc[g];
            const uint32_t n = cnt[g];
            if (n >= k.max_load) continue;
            const PoolEntry * p = e.pool + off[g];
            for (uint32_t i = 0; i < n; ++i) {
                const int d = p[i].dev;
                if (d >= delLower && d <= delUpper) {
                    const uint32_t first = p[i].pos_rel + e.anchor;
                    const uint32_t last = first + (uint32_t)max(0, d + k.inner_off);
                    ++supp; smin = min(smin, first); smax = max(smax, first); lmin = min(lmin, last); lmax = max(lmax, last);
                    const uint32_t slot = atomicAdd(&s_nsupp, 1u);
                    if (slot < SUPP_CAP) { s_first[slot] = first; s_last[slot] = last; }
                }
            }
        }
        if (fl_min == 0xFFFFFFFFu) fl_min = 0;
        double x0, x1, x2, g0l = t0, g1l = t1, g2l = t2;
        if (t0 + t1 + t2 == 0.0) { x0 = 0; x1 = LN1E10; x2 = LN1E10; }       // sum(gtLogs) == 0 :307-308
        else {
            const double mg = fmax(fmax(t0, t1), t2);
            g0l -= mg; g1l -= mg; g2l -= mg;
            if (ndeg) { g1l += ndeg * LOG10_2_RESIDUE; const double m2 = fmax(fmax(g0l, g1l), g2l); g0l -= m2; g1l -= m2; g2l -= m2; }
            if (g0l == g1l && g0l == g2l) { g0l = 0; g1l = -10; g2l = -10; }
            finish_triple(l0, l1, l2, ndeg, x0, x1, x2);
        }
        dlx[3 * s] = x0; dlx[3 * s + 1] = x1; dlx[3 * s + 2] = x2;
        // calculatePhredGL utils_popdel.h:1511-1528
        const double gTot = log10(exp(g0l) + exp(g1l) + exp(g2l));
        const double q0 = -10 * (g0l - gTot), q1 = -10 * (g1l - gTot), q2 = -10 * (g2l - gTot);
        const double mn = fmin(fmin(q0, q1), q2);
        uint32_t * o = ps + 13 * s;
        const bool low = sstat[s] == 0;
        o[0] = low ? 0u : (uint32_t)round(q0 - mn);
        o[1] = low ? 0u : (uint32_t)round(q1 - mn);
        o[2] = low ? 0u : (uint32_t)round(q2 - mn);
        o[3] = lad[0]; o[4] = lad[1]; o[5] = lad[2];
        o[6] = dad[0]; o[7] = dad[1]; o[8] = dad[2]; o[9] = dad[3]; o[10] = dad[4];
        o[11] = fl_min; o[12] = fl_max;
        if (!low) ++ndata;
    }
    __syncthreads();
    block_sum2u(supp, ndata, sh.redu);
    if (supp == 0) { reject(3); return; }
    // percentiles of the supporting starts (80th) and ends (20th): getSuppFirstLast :514-529, by value bisection
    uint32_t sF, sL;
    {
        // block-wide min/max of the supporting positions (shuffles, then warps in index order)
        for (int o = 16; o > 0; o >>= 1) {
            smin = min(smin, __shfl_xor_sync(FULL, smin, o)); smax = max(smax, __shfl_xor_sync(FULL, smax, o));
            lmin = min(lmin, __shfl_xor_sync(FULL, lmin, o)); lmax = max(lmax, __shfl_xor_sync(FULL, lmax, o));
        }
        uint32_t * r32 = reinterpret_cast<uint32_t *>(sh.redu);
        const int wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
        __syncthreads();
        if ((tid & 31) == 0) { r32[4 * wid] = smin; r32[4 * wid + 1] = smax; r32[4 * wid + 2] = lmin; r32[4 * wid + 3] = lmax; }
        __syncthreads();
        for (int i = 0; i < nw; ++i) { smin = min(smin, r32[4 * i]); smax = max(smax, r32[4 * i + 1]); lmin = min(lmin, r32[4 * i + 2]); lmax = max(lmax, r32[4 * i + 3]); }
        __syncthreads();
        const unsigned long long kF = (unsigned long long)round((double)(supp - 1) * 0.8);
        const unsigned long long kL = (unsigned long long)round((double)(supp - 1) * (1 - 0.8));
        uint32_t loF = smin, hiF = smax, loL = lmin, hiL = lmax;
        while (loF < hiF || loL < hiL) {
            const uint32_t midF = loF + (hiF - loF) / 2, midL = loL + (hiL - loL) / 2;
            unsigned long long cF = 0, cL = 0;
            if (supp <= SUPP_CAP) {
                for (uint32_t i = tid; i < (uint32_t)supp; i += blockDim.x) { cF += s_first[i] <= midF; cL += s_last[i] <= midL; }
            } else
            for (uint32_t s = tid; s < a.N; s += blockDim.x) {
                int delLower = INT_MAX, delUpper = 0;
                const uint32_t g0 = a.sample_rg[s], g1 = a.sample_rg[s + 1];
                for (uint32_t g = g0; g < g1; ++g) { const PdRgConst k = a.rgc[g]; if (cnt[g] >= k.max_load) continue; delLower = len - k.lower_q; delUpper = len + k.upper_q; }
                for (uint32_t g = g0; g < g1; ++g) {
                    const PdRgConst k = a.rgc[g];
                    const uint32_t n = cnt[g];
                    if (n >= k.max_load) continue;
                    const PoolEntry * p = e.pool + off[g];
                    for (uint32_t i = 0; i < n; ++i) {
                        const int d = p[i].dev;
                        if (d >= delLower && d <= delUpper) {
                            const uint32_t first = p[i].pos_rel + e.anchor;
                            const uint32_t last = first + (uint32_t)max(0, d + k.inner_off);
                            cF += first <= midF; cL += last <= midL;
                        }
                    }
                }
            }
            block_sum2u(cF, cL, sh.redu);
            if (loF < hiF) { if (cF >= kF + 1) hiF = midF; else loF = midF + 1; }
            if (loL < hiL) { if (cL >= kL + 1) hiL = midL; else loL = midL + 1; }
        }
        sF = loF; sL = loL;
    }
    if (sF == 0 && sL == 0) { reject(4); return; }
    const double lr = block_lr(a, sh, dlx, sh.gt);
    if (tid == 0) {
        const bool ok = lr >= e.min_lr;
        e.valid[blockIdx.x] = ok ? 1 : 0;
        if (e.dbg) { e.dbg[4 * blockIdx.x] = ok ? 0 : 5; e.dbg[4 * blockIdx.x + 1] = sh.len; e.dbg[4 * blockIdx.x + 2] = sh.it; e.dbg[4 * blockIdx.x + 3] = (uint32_t)supp; }
        if (ok) {
            pd_call c;
            c.initial_length = L0; c.iterations = sh.it; c.deletion_length = sh.len;
            c.filter = ((double)ndata / a.N >= e.min_sample_fraction) ? 0u : 4u;
            c.lr = lr; c.frequency = sh.freq;
            const uint32_t cur = e.anchor + w * PD_WIN;
            c.window_position = cur - 1;
            c.position = e.window_wise ? cur - 1 : sF;
            c.end_position = e.window_wise ? 0u : sL;
            c.segment = (uint32_t)(((uint64_t)w * PD_WIN) / a.window_buffer);
            e.calls[blockIdx.x] = c;
        }
    }
}

// gathers the per-sample rows of the valid pairs: out[k] = ps[idx[k]]
__global__ void __launch_bounds__(256) k_compact(const uint32_t * __restrict__ ps, const uint32_t * __restrict__ idx,
                                                 uint32_t * __restrict__ out, uint32_t row_words)
{
    const uint32_t * src = ps + (size_t)idx[blockIdx.x] * row_words;
    uint32_t * dst = out + (size_t)blockIdx.x * row_words;
    for (uint32_t i = threadIdx.x; i < row_words; i += blockDim.x) dst[i] = src[i];
}

template <typename T>
int grow_scratch(pd_ctx * c, int slot, T *& p, size_t need)
{
    size_t bytes = need * sizeof(T);
    if (bytes > c->cap_scratch[slot] || !c->d_scratch[slot]) {
        if (c->d_scratch[slot]) cudaFree(c->d_scratch[slot]);
        c->d_scratch[slot] = nullptr; c->cap_scratch[slot] = 0;
        size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
        PD_CUDA(c, cudaMalloc(&c->d_scratch[slot], want));
        c->cap_scratch[slot] = want;
    }
    p = reinterpret_cast<T *>(c->d_scratch[slot]);
    return 0;
}

}  // namespace

int pd_run_scan(pd_ctx * c, uint64_t first_window, uint64_t n_windows, pd_result * out)
{
    PD_CUDA(c, cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    const uint64_t total = c->n_windows_total;
    uint64_t w_begin = std::min<uint64_t>(first_window, total);
    uint64_t w_end = n_windows ? std::min<uint64_t>(first_window + n_windows, total) : total;
    c->res_calls.clear();
    out->n_reads = c->n_reads;
    out->algorithmic_bytes = 4ull * c->n_reads;
    out->h2d_bytes = c->h2d_bytes;
    out->calls = c->res_calls.data(); out->per_sample = c->res_ps;
    if (w_end <= w_begin) return 0;
    const uint32_t N = c->N, R = c->R;
    const uint32_t W = (uint32_t)(w_end - w_begin);
    const size_t row = 13ull * N;

    PdDev a;
    a.words = c->d_words; a.tiles = c->d_tiles; a.longs = c->d_longs;
    a.rgc = c->d_rgc; a.sample_rg = c->d_sample_rg; a.tab = c->d_tab;
    a.NT = c->NT; a.N = N; a.R = R; a.window_buffer = c->grid.window_buffer; a.t_min = c->t_min;
    a.w_begin = (uint32_t)w_begin; a.w_end = (uint32_t)w_end;

    uint32_t * d_flags, * d_jobs, * d_counters;
    if (grow_scratch(c, 0, d_flags, (size_t)W)) return c->status;
    if (grow_scratch(c, 1, d_jobs, (size_t)W)) return c->status;
    if (grow_scratch(c, 2, d_counters, (size_t)16)) return c->status;
    cudaStream_t st = c->stream;
    PD_CUDA(c, cudaEventRecord(c->ev[2], st));
    PD_CUDA(c, cudaMemsetAsync(d_flags, 0, (size_t)W * 4, st));
    PD_CUDA(c, cudaMemsetAsync(d_counters, 0, 16 * 4, st));
    const uint32_t tile_begin = (uint32_t)(w_begin / PD_TILE_WINDOWS);
    const uint32_t tile_end = (uint32_t)std::min<uint64_t>((w_end + PD_TILE_WINDOWS - 1) / PD_TILE_WINDOWS, c->NT);
    PD_CUDA(c, cudaEventRecord(c->ev[3], st));
    {
        const uint32_t tiles_per_block = 8 * SCREEN_TPW;
        dim3 grid((tile_end - tile_begin + tiles_per_block - 1) / tiles_per_block, N);
        k_screen<<<grid, 256, 0, st>>>(a, tile_begin, tile_end, d_flags, d_jobs, d_counters);
        ++out->n_kernel_launches;
        PD_CUDA(c, cudaGetLastError());
    }
    PD_CUDA(c, cudaEventRecord(c->ev[4], st));
    uint32_t h_cnt[16];
    PD_CUDA(c, cudaMemcpyAsync(h_cnt, d_counters, 16 * 4, cudaMemcpyDeviceToHost, st));
    PD_CUDA(c, cudaStreamSynchronize(st));
    const uint32_t n_jobs = h_cnt[CNT_JOBS];
    out->n_windows = W;
    out->n_flagged_windows = n_jobs;

    std::vector<uint32_t> h_jobs(n_jobs);
    if (n_jobs) {
        PD_CUDA(c, cudaMemcpyAsync(h_jobs.data(), d_jobs, (size_t)n_jobs * 4, cudaMemcpyDeviceToHost, st));
        PD_CUDA(c, cudaStreamSynchronize(st));
        std::sort(h_jobs.begin(), h_jobs.end());
        PD_CUDA(c, cudaMemcpyAsync(d_jobs, h_jobs.data(), (size_t)n_jobs * 4, cudaMemcpyHostToDevice, st));
    }

    // ---- genotyping stage, in batches of flagged windows
    const size_t job_bytes = 8ull * R + 5ull * N + 400ull * N;
    const uint32_t JB = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(3000000000ull / job_bytes, 64), 16384);
    const size_t pair_bytes = 24ull * N + 4ull * R + 52ull * N + 52ull * N;
    const uint32_t PB = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(3000000000ull / pair_bytes, 64), 32768);
    uint32_t npad = 1; while (npad < N) npad <<= 1;
    if ((size_t)npad * 4 > 200 * 1024) return pd_fail(c, PD_ERR_CAPACITY, "more than 51200 samples per context: candidate sort does not fit shared memory (shard by sample)");
    if ((size_t)npad * 4 > 48 * 1024)
        PD_CUDA(c, cudaFuncSetAttribute(k_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(npad * 4)));
    const uint32_t em_threads = N <= 64 ? 64 : 128;                 // k_final: thread = sample (strided)
    uint64_t n_pairs_total = 0, n_calls = 0;
    size_t pool_cap = std::max<size_t>((size_t)std::min<uint32_t>(JB, std::max(n_jobs, 1u)) * R * 40, 1u << 20);
    const bool dbg = getenv("PD_DEBUG") != nullptr;
    const int dbg_window = getenv("PD_DEBUG_WINDOW") ? atoi(getenv("PD_DEBUG_WINDOW")) : -1;

    for (uint32_t job0 = 0; job0 < n_jobs; job0 += JB) {
        const uint32_t nj = std::min(JB, n_jobs - job0);
        uint32_t * d_act_off, * d_act_cnt; int32_t * d_q3; uint8_t * d_sstat; PoolEntry * d_pool = nullptr; PdPair * d_pairs;
        if (grow_scratch(c, 3, d_act_off, (size_t)nj * R)) return c->status;
        if (grow_scratch(c, 4, d_act_cnt, (size_t)nj * R)) return c->status;
        if (grow_scratch(c, 5, d_q3, (size_t)nj * N)) return c->status;
        if (grow_scratch(c, 6, d_sstat, (size_t)nj * N)) return c->status;
        const uint32_t pair_cap = nj * 8 + 1024;
        if (grow_scratch(c, 8, d_pairs, (size_t)pair_cap)) return c->status;
        uint32_t n_pairs = 0;
        for (int attempt = 0; ; ++attempt) {
            if (pool_cap > 0xFFFFFFF0ull) pool_cap = 0xFFFFFFF0ull;
            if (grow_scratch(c, 7, d_pool, pool_cap)) return c->status;
            PD_CUDA(c, cudaMemsetAsync(d_counters + CNT_PAIRS, 0, 2 * 4, st));      // pairs + pool cursor
            GatherArgs ga{d_jobs, job0, nj, d_pool, (uint32_t)pool_cap, d_counters, d_act_off, d_act_cnt, d_q3, d_sstat};
            k_gather<<<dim3(nj, (N + 3) / 4), 128, 0, st>>>(a, ga);
            k_candidates<<<nj, 256, npad * 4, st>>>(a, d_q3, d_sstat, job0, d_counters, d_pairs, pair_cap, npad);
            out->n_kernel_launches += 2;
            PD_CUDA(c, cudaGetLastError());
            PD_CUDA(c, cudaMemcpyAsync(h_cnt, d_counters, 16 * 4, cudaMemcpyDeviceToHost, st));
            PD_CUDA(c, cudaStreamSynchronize(st));
            if (h_cnt[CNT_POOL] <= pool_cap) { n_pairs = h_cnt[CNT_PAIRS]; break; }
            if (attempt == 1) return pd_fail(c, PD_ERR_CAPACITY, "active read-pair pool overflow");
            pool_cap = (size_t)h_cnt[CNT_POOL] + 1024;                              // exact size known now: run again
        }
        if (n_pairs > pair_cap) return pd_fail(c, PD_ERR_CAPACITY, "more than 8 candidate lengths per flagged window on average");
        // deterministic order: (window, initial length)
        std::vector<PdPair> h_pairs(n_pairs);
        if (n_pairs) {
            PD_CUDA(c, cudaMemcpyAsync(h_pairs.data(), d_pairs, (size_t)n_pairs * sizeof(PdPair), cudaMemcpyDeviceToHost, st));
            PD_CUDA(c, cudaStreamSynchronize(st));
            std::sort(h_pairs.begin(), h_pairs.end(), [](const PdPair & x, const PdPair & y) { return x.job != y.job ? x.job < y.job : x.L0 < y.L0; });
            PD_CUDA(c, cudaMemcpyAsync(d_pairs, h_pairs.data(), (size_t)n_pairs * sizeof(PdPair), cudaMemcpyHostToDevice, st));
        }
        n_pairs_total += n_pairs;
        // pairs in chunks: the row compaction + D2H of chunk k (stream2) overlaps the EM of chunk k+1 (stream)
        const uint32_t CH = std::min<uint32_t>(PB, std::max<uint32_t>(2048, (n_pairs + 3) / 4));
        uint32_t chunk_no = 0;
        for (uint32_t p0 = 0; p0 < n_pairs; p0 += CH, ++chunk_no) {
            const uint32_t np = std::min(CH, n_pairs - p0);
            const int par = (int)(chunk_no & 1);
            double * d_dlx; int32_t * d_shifts; uint32_t * d_ps, * d_idx, * d_out_ps; pd_call * d_calls; uint8_t * d_valid; EmState * d_states;
            if (grow_scratch(c, 9, d_dlx, (size_t)np * 3 * N)) return c->status;
            if (grow_scratch(c, 10, d_shifts, (size_t)np * R)) return c->status;
            // double-buffered: read by stream2 while the next chunk is computed
            PD_CUDA(c, cudaEventSynchronize(c->ev[6 + par]));                      // chunk k-2 has left these buffers
            if (grow_scratch(c, 11 + 8 * par, d_ps, (size_t)CH * row)) return c->status;            // slots 11 / 19
            if (grow_scratch(c, 13 + 8 * par, d_out_ps, (size_t)CH * row)) return c->status;        // slots 13 / 21
            if (grow_scratch(c, 15 + 8 * par, d_idx, (size_t)CH)) return c->status;                 // slots 15 / 23
            if (grow_scratch(c, 12, d_calls, (size_t)np)) return c->status;
            if (grow_scratch(c, 14, d_valid, (size_t)np + 16 + (dbg ? (size_t)np * 16 : 0))) return c->status;
            if (grow_scratch(c, 16, d_states, (size_t)np)) return c->status;
            EmArgs e;
            e.jobs = d_jobs; e.pairs = d_pairs; e.pair0 = p0; e.npairs = np; e.job_base = job0;
            e.pool = d_pool; e.act_off = d_act_off; e.act_cnt = d_act_cnt; e.sstat = d_sstat;
            e.dlx = d_dlx; e.shifts = d_shifts; e.ps = d_ps; e.calls = d_calls; e.valid = d_valid;
            e.iterations = c->params.iterations; e.min_len = c->params.min_len; e.min_lr = c->params.min_lr;
            e.min_sample_fraction = c->params.min_sample_fraction; e.somatic = c->params.somatic; e.window_wise = c->params.window_wise;
            e.anchor = c->grid.anchor;
            e.dbg = nullptr; e.dbg_window = dbg_window;
            if (dbg) {
                e.dbg = reinterpret_cast<uint32_t *>(d_valid + (((size_t)np + 15) & ~(size_t)15));
                PD_CUDA(c, cudaMemsetAsync(e.dbg, 0xFF, (size_t)np * 16, st));
            }
            {
                uint32_t lps = N * 8 <= 512 ? 8 : (N * 4 <= 512 ? 4 : (N * 2 <= 512 ? 2 : 1));
                if (getenv("PD_EM_LPS")) lps = (uint32_t)atoi(getenv("PD_EM_LPS"));          // tuning knob
                const uint32_t T = std::min<uint32_t>(lps >= 8 ? 1024 : 512, ((N * lps + 31) / 32) * 32);
                const size_t smem = (size_t)2 * EM_CACHE_SLOTS * T * sizeof(double);
                auto launch = [&](auto kern) -> cudaError_t {
                    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (err != cudaSuccess) return err;
                    kern<<<np, T, smem, st>>>(a, e, d_states);
                    return cudaGetLastError();
                };
                cudaError_t err = lps == 8 ? launch(k_em<8>) : lps == 4 ? launch(k_em<4>) : lps == 2 ? launch(k_em<2>) : launch(k_em<1>);
                if (err != cudaSuccess) return pd_fail(c, PD_ERR_CUDA, std::string("k_em launch: ") + cudaGetErrorString(err));
                k_final<<<np, em_threads, 0, st>>>(a, e, d_states);
            }
            out->n_kernel_launches += 2;
            PD_CUDA(c, cudaGetLastError());
            std::vector<uint8_t> h_valid(np);
            std::vector<pd_call> h_calls(np);
            PD_CUDA(c, cudaMemcpyAsync(h_valid.data(), d_valid, np, cudaMemcpyDeviceToHost, st));
            PD_CUDA(c, cudaMemcpyAsync(h_calls.data(), d_calls, (size_t)np * sizeof(pd_call), cudaMemcpyDeviceToHost, st));
            PD_CUDA(c, cudaStreamSynchronize(st));
            if (e.dbg) {
                std::vector<uint32_t> hd((size_t)np * 4);
                cudaMemcpy(hd.data(), e.dbg, (size_t)np * 16, cudaMemcpyDeviceToHost);
                for (uint32_t i = 0; i < np; ++i)
                    fprintf(stderr, "PD_DEBUG pair window %u L0 %d reason %u len %u it %u supp %u\n", h_jobs[h_pairs[p0 + i].job], h_pairs[p0 + i].L0, hd[4 * i], hd[4 * i + 1], hd[4 * i + 2], hd[4 * i + 3]);
            }
            std::vector<uint32_t> & idx = c->idx_stage[par];                       // must outlive the async H2D below
            idx.clear();
            for (uint32_t i = 0; i < np; ++i) if (h_valid[i]) { idx.push_back(i); c->res_calls.push_back(h_calls[i]); }
            const uint32_t nc = (uint32_t)idx.size();
            if (nc) {
                const size_t need = (n_calls + nc) * row;
                if (need > c->cap_res_ps) {                       // grow the pinned result buffer (rare; drains stream2 first)
                    PD_CUDA(c, cudaStreamSynchronize(c->stream2));
                    size_t want = std::max<size_t>(need + need / 2, 1u << 20);
                    uint32_t * p = nullptr;
                    PD_CUDA(c, cudaMallocHost(&p, want * 4));
                    if (n_calls) memcpy(p, c->res_ps, n_calls * row * 4);
                    if (c->res_ps) cudaFreeHost(c->res_ps);
                    c->res_ps = p; c->cap_res_ps = want;
                }
                cudaStream_t s2 = c->stream2;
                PD_CUDA(c, cudaMemcpyAsync(d_idx, idx.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, s2));
                k_compact<<<nc, 256, 0, s2>>>(d_ps, d_idx, d_out_ps, (uint32_t)row);
                ++out->n_kernel_launches;
                PD_CUDA(c, cudaGetLastError());
                PD_CUDA(c, cudaMemcpyAsync(c->res_ps + n_calls * row, d_out_ps, (size_t)nc * row * 4, cudaMemcpyDeviceToHost, s2));
                PD_CUDA(c, cudaEventRecord(c->ev[6 + par], s2));
                out->d2h_bytes += (uint64_t)nc * (sizeof(pd_call) + row * 4);
                n_calls += nc;
            }
        }
        PD_CUDA(c, cudaStreamSynchronize(c->stream2));
    }
    PD_CUDA(c, cudaStreamSynchronize(c->stream2));
    PD_CUDA(c, cudaEventRecord(c->ev[5], st));
    PD_CUDA(c, cudaStreamSynchronize(st));
    out->n_calls = c->res_calls.size();
    out->calls = c->res_calls.data();
    out->per_sample = c->res_ps;
    out->n_candidates = n_pairs_total;
    PD_CUDA(c, cudaEventElapsedTime(&out->ms_screen, c->ev[3], c->ev[4]));
    PD_CUDA(c, cudaEventElapsedTime(&out->ms_genotype, c->ev[4], c->ev[5]));
    PD_CUDA(c, cudaEventElapsedTime(&out->ms_total, c->ev[2], c->ev[5]));
    out->ms_d2h = 0;
    return 0;
}
