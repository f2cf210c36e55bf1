// pd_em_common.cuh -- device helpers shared by the genotyping kernels (pd_em.cu, pd_shard.cu). Internal.
#ifndef PD_EM_COMMON_CUH_
#define PD_EM_COMMON_CUH_

#include <cfloat>
#include <cmath>

#include "pd_device.cuh"

namespace {

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
    double red[2][64];                 // double-buffered partial sums of the block reductions (<= 32 warps)
    unsigned long long redu[64];
    double rgw[3];                     // read group 0's likelihood triple, exp domain (quirk: drives all reference shifts)
    int visited_len[64]; double visited_freq[64]; int nvisited;
    uint32_t sel[2];                   // supporting start / end percentiles (k_final)
    unsigned long long xr_in[PD_MAX_WORLD][4];    // values received from every rank (sample sharding)
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PD_FULL, v, o);
    return v;
}
// deterministic block sums (fixed shuffle tree, then warps in index order); result on all threads. One barrier:
// consecutive calls alternate between the two halves of `red` (parity p).
__device__ __forceinline__ void block_sum2(double & a, double & b, double (*red)[64], int & p)
{
    a = warp_sum(a); b = warp_sum(b);
    const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double * r = red[p];
    p ^= 1;
    if ((threadIdx.x & 31) == 0) { r[2 * wid] = a; r[2 * wid + 1] = b; }
    __syncthreads();
    double sa = 0, sb = 0;
    for (int i = 0; i < nw; ++i) { sa += r[2 * i]; sb += r[2 * i + 1]; }
    a = sa; b = sb;
}
__device__ __forceinline__ void block_sum2u(unsigned long long & a, unsigned long long & b, unsigned long long * red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(PD_FULL, a, o); b += __shfl_xor_sync(PD_FULL, b, o); }
    const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[2 * wid] = a; red[2 * wid + 1] = b; }
    __syncthreads();
    unsigned long long sa = 0, sb = 0;
    for (int i = 0; i < nw; ++i) { sa += red[2 * i]; sb += red[2 * i + 1]; }
    a = sa; b = sb;
}

// ---------------------------------------------------------------------------------------------------------------
// Cross-rank reductions of the sample-sharded scan: every rank runs the SAME (window, length) pair in a
// co-resident block; the per-pair sufficient statistics are exchanged through peer memory inside the kernel -- rank r
// stores its four values and a sequence number into slot [pair][reduction parity][r] of EVERY rank's array (NVLink
// stores, release at system scope) and polls its own array (acquire) -- and are then combined in rank order, so every
// rank holds bit-identical results and takes identical branches. `v*` must be block-uniform on entry.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long * p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long * p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct XrBlock { uint32_t pair; unsigned long long seq; };          // per block: pair slot and reduction counter

// exchanges four raw 64-bit values; afterwards sh.xr_in[r][i] = value i of rank r (all threads may read it)
__device__ __forceinline__ void xr_exchange(const XrArgs & x, XrBlock & xb, EmShared & sh, unsigned long long v0, unsigned long long v1,
                                            unsigned long long v2, unsigned long long v3)
{
    ++xb.seq;
    __syncthreads();                                              // earlier readers of sh.xr_in are done
    const uint32_t tid = threadIdx.x;
    if (tid < x.world) {
        const size_t base = ((((size_t)(x.epoch & 1) * x.pairs_cap + xb.pair) * 2 + (xb.seq & 1)) * x.world);
        XrSlot * dst = x.peer[tid] + base + x.rank;
        st_relaxed_sys(&dst->v[0], v0); st_relaxed_sys(&dst->v[1], v1); st_relaxed_sys(&dst->v[2], v2); st_relaxed_sys(&dst->v[3], v3);
        st_release_sys(&dst->seq, xb.seq);
        const XrSlot * src = x.peer[x.rank] + base + tid;
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(&src->seq) != xb.seq) {
            if (*(volatile uint32_t *)x.err) { ok = false; break; }
            if (clock64() - t0 > 6000000000ll) { atomicExch(x.err, 1u); ok = false; break; }     // ~3 s: a peer is gone
            __nanosleep(40);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) sh.xr_in[tid][i] = ok ? ld_relaxed_sys(&src->v[i]) : 0ull;
    }
    __syncthreads();
}
__device__ __forceinline__ void xr_sum4(const XrArgs & x, XrBlock & xb, EmShared & sh, double & a, double & b, double & c, double & d)
{
    xr_exchange(x, xb, sh, __double_as_longlong(a), __double_as_longlong(b), __double_as_longlong(c), __double_as_longlong(d));
    a = b = c = d = 0;
    for (uint32_t r = 0; r < x.world; ++r) {
        a += __longlong_as_double(sh.xr_in[r][0]); b += __longlong_as_double(sh.xr_in[r][1]);
        c += __longlong_as_double(sh.xr_in[r][2]); d += __longlong_as_double(sh.xr_in[r][3]);
    }
}
__device__ __forceinline__ void xr_sum2(const XrArgs & x, XrBlock & xb, EmShared & sh, double & a, double & b)
{
    double c = 0, d = 0;
    xr_sum4(x, xb, sh, a, b, c, d);
}
__device__ __forceinline__ void xr_sum2u(const XrArgs & x, XrBlock & xb, EmShared & sh, unsigned long long & a, unsigned long long & b)
{
    xr_exchange(x, xb, sh, a, b, 0ull, 0ull);
    a = b = 0;
    for (uint32_t r = 0; r < x.world; ++r) { a += sh.xr_in[r][0]; b += sh.xr_in[r][1]; }
}
// minima of (a, c) and maxima of (b, d) over the ranks
__device__ __forceinline__ void xr_minmax(const XrArgs & x, XrBlock & xb, EmShared & sh, uint32_t & a, uint32_t & b, uint32_t & c, uint32_t & d)
{
    xr_exchange(x, xb, sh, a, b, c, d);
    for (uint32_t r = 0; r < x.world; ++r) {
        a = min(a, (uint32_t)sh.xr_in[r][0]); b = max(b, (uint32_t)sh.xr_in[r][1]);
        c = min(c, (uint32_t)sh.xr_in[r][2]); d = max(d, (uint32_t)sh.xr_in[r][3]);
    }
}

// every exit of the EM / final-pass bodies leaves valid / calls / ps of the pair final: publish that to the concurrent
// emitter (k_emit_stream)
__device__ __forceinline__ void publish_done(const EmArgs & e, uint32_t bid)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(e.done + bid), "r"(1u) : "memory");
}

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

struct RgLite { int hist_base; uint32_t hist_len, hist_off, max_load; double min_prob, ln_min_prob; };
__device__ __forceinline__ RgLite rg_lite(const PdRgConst * r)
{
    RgLite k;
    k.hist_base = __ldg(&r->hist_base); k.hist_len = __ldg(&r->hist_len); k.hist_off = __ldg(&r->hist_off);
    k.max_load = __ldg(&r->max_load); k.min_prob = __ldg(&r->min_prob); k.ln_min_prob = __ldg(&r->ln_min_prob);
    return k;
}
// I() (insert_histogram_popdel.h:1157-1163): table row of deviation `dev`; row 0 of a read group is the floor entry
template <typename K>
__device__ __forceinline__ bool tab_in(const K & k, int dev) { return (uint32_t)(dev + k.hist_base - 1) < k.hist_len - 2u; }
template <typename K>
__device__ __forceinline__ const PdTab * tab_at(const PdTab * __restrict__ tab, const K & k, int dev)
{
    return tab + k.hist_off + (tab_in(k, dev) ? dev + k.hist_base + 1 : 0);
}
__device__ __forceinline__ double2 ld2(const double * p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
// one 256-bit load = half a table row (32-byte aligned): a single L1 request per lane instead of two
struct D4 { double a, b, c, d; };
__device__ __forceinline__ D4 ld4(const double * p)
{
    D4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}

template <int LPS>
__device__ __forceinline__ double group_sum(double v, uint32_t gmask)
{
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}
template <int LPS>
__device__ __forceinline__ uint32_t group_sum(uint32_t v, uint32_t gmask)
{
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}
template <int LPS>
__device__ __forceinline__ uint32_t group_mask()
{
    return (LPS == 32) ? PD_FULL : (((1u << LPS) - 1u) << ((threadIdx.x & 31) & ~(LPS - 1)));
}

// posterior weight of the deletion hypothesis of one read pair, r = del / (del + ref), without the likelihood terms
// (uncached read pairs of the length update)
__device__ __forceinline__ double pair_weight(const PdTab * __restrict__ tab, const RgLite & k, int d, int shift, int L)
{
    const PdTab * tr = tab_at(tab, k, d - shift);
    const double ref = __ldg(&tr->val);
    if (!tab_in(k, d - L)) return ref == k.min_prob ? 0.5 : __ldg(&tr->fr);
    const PdTab * td = tab + k.hist_off + (d - L + k.hist_base + 1);
    const double del = __ldg(&td->val);
    if (ref == del) return 0.5;
    if (del == k.min_prob) return __ldg(&tr->fr);
    if (ref == k.min_prob) return 1.0 - __ldg(&td->fr);
    return del / (del + ref);
}

template <int LPS>
__device__ __forceinline__ uint32_t group_min(uint32_t v, uint32_t gmask)
{
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(gmask, v, o));
    return v;
}
template <int LPS>
__device__ __forceinline__ uint32_t group_max(uint32_t v, uint32_t gmask)
{
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(gmask, v, o));
    return v;
}

}  // namespace

#endif
