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
