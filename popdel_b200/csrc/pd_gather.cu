// pd_gather.cu -- K2: active sets, per-sample upper-half medians and candidate deletion lengths of the flagged windows.
//
//   k_tile_gather  one warp = (flagged tile, sample): the read pairs that can be active anywhere in the tile are decoded
//                  ONCE into shared memory (interval, position, deviation); every flagged window of the tile then only
//                  filters that staged list: active read pairs -> pool (structure of arrays), coverage state, and the
//                  sample's Q3 (upperHalfMedian, genotype_deletion_popdel_call.h:15-27) by a warp bitonic sort.
//   k_candidates   one block per flagged window: sort the Q3s over samples, gap-50 clustering with rank-indexed
//                  thresholds (:58-86) -> candidate initial lengths, kept inline per window.
//   k_cand_*       exclusive scan of the candidate counts -> (window, initial length) pairs in reference order.
#include "pd_device.cuh"

namespace {

constexpr int TG_WARPS = 4;
constexpr int TG_STAGE = 320;                                       // staged read pairs per warp
constexpr int TG_ACT = 96;                                          // active read pairs per (window, sample) via the fast path
constexpr int TG_MAXRG = 8;                                         // read groups per sample via the fast path

struct alignas(16) StagePair { int32_t s, e; uint32_t pos; int32_t dev; };

struct alignas(16) WarpStage {
    StagePair pair[TG_STAGE];
    uint16_t act[TG_ACT];
    uint32_t rg_first[TG_MAXRG + 1];
};

// upperHalfMedian (:15-27) from the order statistics l and l+1: n<4 -> the maximum; else interpolate at (3n+2+n%2)/4-1
__device__ __forceinline__ void q3_position(uint32_t nn, uint32_t & l, uint32_t & l2, double & r)
{
    r = 0;
    if (nn < 4) l = nn - 1;
    else { const double pos = (3.0 * nn + 2.0 + (nn % 2)) / 4.0 - 1.0; l = (uint32_t)pos; r = pos - l; }
    l2 = (l + 1 < nn) ? l + 1 : l;
}
__device__ __forceinline__ int32_t q3_value(uint32_t nn, double r, int32_t lo_v, int32_t hi_v)
{
    if (nn < 4) return (int32_t)floor((double)lo_v + 0.5);
    return (int32_t)floor((1 - r) * lo_v + r * hi_v + 0.5);
}

// ascending bitonic sort of one value per lane
__device__ __forceinline__ int32_t warp_sort(int32_t v, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int32_t o = __shfl_xor_sync(PD_FULL, v, j);
            const bool up = ((lane & k) == 0) == ((lane & j) == 0);      // keep the minimum
            v = up ? min(v, o) : max(v, o);
        }
    return v;
}

// Generic (slow) path of one (window, sample): streams the read groups twice, no shared memory. Used when the staged
// list or the active set does not fit the fast path.
__device__ __noinline__ void gather_window_slow(const PdDev & a, const GatherArgs & ga, uint32_t smp, int32_t w, uint32_t job, int lane)
{
    const uint32_t g0 = a.sample_rg[smp], g1 = a.sample_rg[smp + 1];
    uint32_t * cnt = ga.act_cnt + (size_t)job * a.R;
    uint32_t * off = ga.act_off + (size_t)job * a.R;
    const uint32_t tile = (uint32_t)w / PD_TILE_WINDOWS;
    uint32_t cov = 0, nvals = 0;
    for (uint32_t g = g0; g < g1; ++g) {
        const PdRgConst k = a.rgc[g];
        uint32_t n_g = 0;
        for_tile_batches(a, g, k, tile, lane, [&](bool valid, int32_t s, int32_t e, uint32_t, int32_t) {
            n_g += __popc(__ballot_sync(PD_FULL, valid && s <= w && w <= e));
        });
        if (lane == 0) { cnt[g] = n_g; off[g] = nvals; }
        cov += n_g;
        if (n_g < k.max_load) nvals += n_g;
    }
    uint32_t base = 0;
    if (lane == 0 && nvals) base = atomicAdd(&ga.counters[CNT_POOL], nvals);
    base = __shfl_sync(PD_FULL, base, 0);
    const bool fits = (uint64_t)base + nvals <= ga.pool_cap;
    __syncwarp();
    if (lane == 0) for (uint32_t g = g0; g < g1; ++g) off[g] += base;
    if (fits) {
        uint32_t cur = base;
        for (uint32_t g = g0; g < g1; ++g) {
            const PdRgConst k = a.rgc[g];
            const uint32_t n_g = __shfl_sync(PD_FULL, lane == 0 ? cnt[g] : 0u, 0);
            if (n_g >= k.max_load) continue;
            for_tile_batches(a, g, k, tile, lane, [&](bool valid, int32_t s, int32_t e, uint32_t pr, int32_t dev) {
                valid = valid && s <= w && w <= e;
                const uint32_t mask = __ballot_sync(PD_FULL, valid);
                if (valid) { const uint32_t i = cur + __popc(mask & ((1u << lane) - 1u)); ga.pool_pos[i] = pr; ga.pool_dev[i] = dev; }
                cur += __popc(mask);
            });
        }
    }
    __syncwarp();
    uint8_t stt; int32_t q = 0, dmx = INT_MIN;
    if (cov < 2u) stt = 0;
    else if (nvals == 0 || !fits) stt = 1;
    else {
        stt = 2;
        uint32_t l, l2; double r;
        q3_position(nvals, l, l2, r);
        int32_t lo_v = 0, hi_v = 0;
        const volatile int32_t * v = ga.pool_dev + base;                 // written above by this warp
        for (uint32_t i0 = 0; i0 < nvals; i0 += 32) {
            const uint32_t i = i0 + lane;
            const int32_t vi = i < nvals ? v[i] : INT_MAX;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < nvals; ++j) { const int32_t vj = v[j]; rank += (vj < vi) || (vj == vi && j < i); }
            const uint32_t m1 = __ballot_sync(PD_FULL, i < nvals && rank == l);
            const uint32_t m2 = __ballot_sync(PD_FULL, i < nvals && rank == l2);
            if (m1) lo_v = __shfl_sync(PD_FULL, vi, __ffs(m1) - 1);
            if (m2) hi_v = __shfl_sync(PD_FULL, vi, __ffs(m2) - 1);
        }
        q = q3_value(nvals, r, lo_v, hi_v);
    }
    if (fits && nvals) {
        const volatile int32_t * v = ga.pool_dev + base;
        for (uint32_t i = lane; i < nvals; i += 32) dmx = max(dmx, v[i]);
        for (int o = 16; o > 0; o >>= 1) dmx = max(dmx, __shfl_xor_sync(PD_FULL, dmx, o));
    }
    if (lane == 0) { ga.q3[(size_t)job * a.N + smp] = q; ga.sstat[(size_t)job * a.N + smp] = stt; ga.dmax[(size_t)job * a.N + smp] = dmx; }
}

__global__ void __launch_bounds__(TG_WARPS * 32) k_tile_gather(PdDev a, GatherArgs ga)
{
    __shared__ WarpStage stage_all[TG_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t smp = blockIdx.y * TG_WARPS + wib;
    if (smp >= a.N) return;
    const uint32_t tj = ga.tj0 + blockIdx.x;
    const uint32_t tile = ga.tj_tile[tj], wmask = ga.tj_mask[tj];
    const uint32_t job_first = ga.tj_wbase[tj] - ga.job_base;            // scratch row of the tile's first flagged window
    const uint32_t g0 = a.sample_rg[smp], g1 = a.sample_rg[smp + 1], nrg = g1 - g0;
    WarpStage & st = stage_all[wib];
    const int32_t w0 = (int32_t)(tile * PD_TILE_WINDOWS);

    // ---- stage every read pair whose interval intersects the tile (per read group, stream order)
    uint32_t total = 0;
    bool fast = nrg <= (uint32_t)TG_MAXRG;
    if (fast) {
        for (uint32_t g = g0; g < g1; ++g) {
            if (lane == 0) st.rg_first[g - g0] = total;
            const PdRgConst k = a.rgc[g];
            for_tile_batches(a, g, k, tile, lane, [&](bool valid, int32_t s, int32_t e, uint32_t pr, int32_t dev) {
                valid = valid && e >= w0 && s <= w0 + 31;
                const uint32_t mask = __ballot_sync(PD_FULL, valid);
                const uint32_t slot = total + __popc(mask & ((1u << lane) - 1u));
                if (valid && slot < (uint32_t)TG_STAGE) st.pair[slot] = StagePair{s, e, pr, dev};
                total += __popc(mask);
            });
        }
        if (lane == 0) st.rg_first[nrg] = total;
        fast = total <= (uint32_t)TG_STAGE;
    }
    __syncwarp();

    // one pool reservation per (tile, sample): the staged read pairs' active windows among the flagged ones (an upper
    // bound of what is written: a high-coverage read group is left out of the pool)
    uint32_t tile_base = 0; bool tile_fits = true;
    if (fast) {
        uint32_t tot = 0;
        for (uint32_t i = lane; i < total; i += 32) {
            const int2 se = *reinterpret_cast<const int2 *>(&st.pair[i]);
            const int sr = max(se.x - w0, 0), er = min(se.y - w0, 31);
            tot += __popc(wmask & (er == 31 ? 0xFFFFFFFFu : ((1u << (er + 1)) - 1u)) & ~((1u << sr) - 1u));
        }
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(PD_FULL, tot, o);
        if (lane == 0 && tot) tile_base = atomicAdd(&ga.counters[CNT_POOL], tot);
        tile_base = __shfl_sync(PD_FULL, tile_base, 0);
        tile_fits = (uint64_t)tile_base + tot <= ga.pool_cap;
    }

    uint32_t rank = 0;
    for (uint32_t m = wmask; m; m &= m - 1, ++rank) {
        const int32_t w = w0 + (__ffs(m) - 1);
        const uint32_t job = job_first + rank;
        if (!fast) { gather_window_slow(a, ga, smp, w, job, lane); continue; }
        uint32_t * cnt = ga.act_cnt + (size_t)job * a.R;
        uint32_t * off = ga.act_off + (size_t)job * a.R;
        uint32_t cov = 0, nvals = 0;
        for (uint32_t gi = 0; gi < nrg; ++gi) {
            const uint32_t i_lo = st.rg_first[gi], i_hi = st.rg_first[gi + 1];
            uint32_t n_g = 0;
            for (uint32_t i0 = i_lo; i0 < i_hi; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool valid = false;
                if (i < i_hi) { const int2 se = *reinterpret_cast<const int2 *>(&st.pair[i]); valid = se.x <= w && w <= se.y; }
                const uint32_t mask = __ballot_sync(PD_FULL, valid);
                const uint32_t slot = nvals + n_g + __popc(mask & ((1u << lane) - 1u));
                if (valid && slot < (uint32_t)TG_ACT) st.act[slot] = (uint16_t)i;
                n_g += __popc(mask);
            }
            if (lane == 0) { cnt[g0 + gi] = n_g; off[g0 + gi] = nvals; }
            cov += n_g;
            if (n_g < __ldg(&a.rgc[g0 + gi].max_load)) nvals += n_g;          // a high-coverage read group is left out
        }
        if (nvals > (uint32_t)TG_ACT) { gather_window_slow(a, ga, smp, w, job, lane); continue; }
        const uint32_t base = tile_base;
        const bool fits = tile_fits;
        tile_base += nvals;
        __syncwarp();
        if ((uint32_t)lane < nrg) off[g0 + lane] += base;
        int32_t dmx = INT_MIN;
        for (uint32_t i = lane; i < nvals; i += 32) {
            const StagePair p = st.pair[st.act[i]];
            dmx = max(dmx, p.dev);
            if (fits) { ga.pool_pos[base + i] = p.pos; ga.pool_dev[base + i] = p.dev; }
        }
        for (int o = 16; o > 0; o >>= 1) dmx = max(dmx, __shfl_xor_sync(PD_FULL, dmx, o));
        uint8_t stt; int32_t q = 0;
        if (cov < 2u) stt = 0;
        else if (nvals == 0 || !fits) stt = 1;
        else {
            stt = 2;
            uint32_t l, l2; double r;
            q3_position(nvals, l, l2, r);
            int32_t lo_v, hi_v;
            if (nvals <= 32) {
                const int32_t v = warp_sort((uint32_t)lane < nvals ? st.pair[st.act[lane]].dev : INT_MAX, lane);
                lo_v = __shfl_sync(PD_FULL, v, (int)l);
                hi_v = __shfl_sync(PD_FULL, v, (int)l2);
            } else {                                                      // 33..TG_ACT values: rank counting in shared memory
                lo_v = hi_v = 0;
                for (uint32_t i0 = 0; i0 < nvals; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const int32_t vi = i < nvals ? st.pair[st.act[i]].dev : INT_MAX;
                    uint32_t rk = 0;
                    for (uint32_t j = 0; j < nvals; ++j) { const int32_t vj = st.pair[st.act[j]].dev; rk += (vj < vi) || (vj == vi && j < i); }
                    const uint32_t m1 = __ballot_sync(PD_FULL, i < nvals && rk == l);
                    const uint32_t m2 = __ballot_sync(PD_FULL, i < nvals && rk == l2);
                    if (m1) lo_v = __shfl_sync(PD_FULL, vi, __ffs(m1) - 1);
                    if (m2) hi_v = __shfl_sync(PD_FULL, vi, __ffs(m2) - 1);
                }
            }
            q = q3_value(nvals, r, lo_v, hi_v);
        }
        if (lane == 0) { ga.q3[(size_t)job * a.N + smp] = q; ga.sstat[(size_t)job * a.N + smp] = stt; ga.dmax[(size_t)job * a.N + smp] = dmx; }
        __syncwarp();                                                     // st.act is rewritten by the next window
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K2b: candidates. One block per flagged window. mode 0: count + inline list; mode 1: windows with more than
// PD_CAND_INLINE candidates write their pairs directly (after the scan of the counts).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_candidates(PdDev a, CandArgs ca, int mode)
{
    extern __shared__ int32_t sv[];
    __shared__ uint32_t s_n;
    const uint32_t job = blockIdx.x;
    if (mode == 1 && ca.cand_cnt[job] <= (uint32_t)PD_CAND_INLINE) return;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < a.N; s += blockDim.x)
        if (ca.sstat[(size_t)job * a.N + s] == 2) sv[atomicAdd(&s_n, 1u)] = ca.q3[(size_t)job * a.N + s];
    __syncthreads();
    const uint32_t nv = s_n;
    if (nv == 0) { if (threadIdx.x == 0 && mode == 0) ca.cand_cnt[job] = 0; return; }
    uint32_t np2 = 1; while (np2 < nv) np2 <<= 1;                 // sort only the occupied power of two
    for (uint32_t i = nv + threadIdx.x; i < np2; i += blockDim.x) sv[i] = INT_MAX;
    __syncthreads();
    for (uint32_t k = 2; k <= np2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const int32_t x = sv[i], y = sv[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { sv[i] = y; sv[ixj] = x; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        // genotype_deletion_popdel_call.h:62-84; thresholds are indexed by RANK in the sorted array (quirk)
        int sum = sv[0], n = 1;
        uint32_t thr = a.rgc[0].min_init, nc = 0;
        const uint32_t pair0 = mode == 1 ? ca.cand_off[job] : 0;
        auto emit = [&](int mean) {
            if (mode == 0) { if (nc < (uint32_t)PD_CAND_INLINE) ca.cand_inline[(size_t)job * PD_CAND_INLINE + nc] = mean; }
            else if (pair0 + nc < ca.pair_cap) ca.pairs[pair0 + nc] = PdPair{ca.job_base + job, mean};
            ++nc;
        };
        for (uint32_t i = 1; i < nv; ++i) {
            if (sv[i - 1] + 50 > sv[i]) { sum += sv[i]; ++n; thr = min(thr, a.rgc[i].min_init); }
            else { if (sum / n > (int)thr) emit(sum / n); sum = sv[i]; n = 1; thr = a.rgc[i].min_init; }
        }
        if (sum / n > (int)thr) emit(sum / n);
        if (mode == 0) ca.cand_cnt[job] = nc;
    }
}

// exclusive scan of cand_cnt -> cand_off, total -> counters[CNT_PAIRS]; inline candidates -> pairs
__global__ void __launch_bounds__(1024) k_cand_sums(CandArgs ca)
{
    __shared__ unsigned long long ws[33];
    const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    unsigned long long total;
    block_excl_scan(i < ca.njobs ? ca.cand_cnt[i] : 0u, ws, total);
    if (threadIdx.x == 0) ca.block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_cand_offsets(CandArgs ca, uint32_t nb)
{
    __shared__ unsigned long long ws[33];
    unsigned long long carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        unsigned long long total;
        const unsigned long long ex = block_excl_scan(b < nb ? ca.block_sums[b] : 0ull, ws, total);
        if (b < nb) ca.block_sums[b] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) ca.counters[CNT_PAIRS] = (uint32_t)min(carry, 0xFFFFFFFFull);
}
__global__ void __launch_bounds__(1024) k_cand_write(CandArgs ca)
{
    __shared__ unsigned long long ws[33];
    const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    const uint32_t c = i < ca.njobs ? ca.cand_cnt[i] : 0u;
    unsigned long long total;
    const uint32_t o = (uint32_t)(block_excl_scan(c, ws, total) + ca.block_sums[blockIdx.x]);
    if (i >= ca.njobs) return;
    ca.cand_off[i] = o;
    if (c <= (uint32_t)PD_CAND_INLINE)
        for (uint32_t k = 0; k < c; ++k)
            if (o + k < ca.pair_cap) ca.pairs[o + k] = PdPair{ca.job_base + i, ca.cand_inline[(size_t)i * PD_CAND_INLINE + k]};
}

}  // namespace

void pd_launch_gather(const PdDev & a, const GatherArgs & g, cudaStream_t st, uint64_t * launches)
{
    k_tile_gather<<<dim3(g.ntj, (a.N + TG_WARPS - 1) / TG_WARPS), TG_WARPS * 32, 0, st>>>(a, g);
    ++*launches;
}

int pd_launch_candidates(pd_ctx * c, const PdDev & a, const CandArgs & ca, cudaStream_t st, uint64_t * launches)
{
    if ((size_t)ca.npad * 4 > 48 * 1024)
        PD_CUDA(c, cudaFuncSetAttribute(k_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ca.npad * 4)));
    const uint32_t threads = ca.npad >= 512 ? 256 : (ca.npad >= 128 ? 64 : 32);
    const uint32_t nb = (ca.njobs + 1023) / 1024;
    k_candidates<<<ca.njobs, threads, ca.npad * 4, st>>>(a, ca, 0);
    k_cand_sums<<<nb, 1024, 0, st>>>(ca);
    k_cand_offsets<<<1, 1024, 0, st>>>(ca, nb);
    k_cand_write<<<nb, 1024, 0, st>>>(ca);
    k_candidates<<<ca.njobs, threads, ca.npad * 4, st>>>(a, ca, 1);
    *launches += 5;
    PD_CUDA(c, cudaGetLastError());
    return 0;
}
