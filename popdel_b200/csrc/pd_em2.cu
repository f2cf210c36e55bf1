// pd_em2.cu -- K3-K5, second generation: the genotyping of every (window, initial length) pair as a BULK-SYNCHRONOUS,
// SAMPLE-MAJOR pipeline (replaces the pair-major kernels of pd_em.cu for cohorts that are not sharded by sample).
//
// Why: the pair-major kernels gather one 32-byte likelihood-table row per read pair and pass from tables that only
// fit L2 (100 read groups x 302 rows x 64 B); every lane of a warp belongs to another read group, so every gather is
// its own L1 wavefront (~2 cycles each, profiles/r02/em_stats_k_em_one_baseline.json: 4.7 cycles per gather measured).
// Here a block owns ONE read group: its table (32 B per row for the EM, 24 B for the final pass) is staged in
// shared memory once per block and the gathers become shared-memory loads.
//
//   k_e2_prep         per (read group, group of 2048 pairs): counting sort of the group's items (pair, read group) by
//                     (carrier of the initial length, number of active read pairs), so that the 32 lanes of a warp get
//                     items of equal length and the items that need later passes share warps; copies the items'
//                     deviations / positions from the window-major pool into a lane-interleaved layout [j][lane].
//   k_e2_reads<A>     sample-major data-likelihood pass (compute_data_likelihoods, EM overload,
//                     genotype_deletion_popdel_call.h:179-253): lane = item, sums ln ref / ln((ref+del)/2) / ln del and
//                     the posterior-weight moments sum r, sum r*d (r = del/(del+ref)) in read-pair order. Pass A (initial
//                     length, zero shifts) also counts initialize_allele_frequency's window (:93-133). An item is only
//                     recomputed when its reference shift changed or it carries read pairs inside the deletion
//                     hypothesis' histogram at the old or the new length.
//   k_e2_pair         pair-major, one block per pair: per-sample triples (finish_triple), allele frequency
//                     (:467-485), convergence test and the partial "previous state" comparison (:598-660), length and
//                     reference-shift update from the moments (:388-462, rgDlIt quirk), requests the next pass.
//   k_e2_final_reads  sample-major final pass (:255-337 without the ln sums, which equal the last EM pass): log10
//                     likelihood sums, LAD, DAD (:137-172), first/last positions, supporting read pairs.
//   k_e2_final_pair   per surviving pair: PL (utils_popdel.h:1511-1528), percentiles of the supporting read pairs
//                     (:514-529), likelihood ratio (:490-508), the Call.
// Control flow between the kernels travels in device memory (per-item control words, per-pair state); the host
// enqueues the fixed sequence prep, A, pair, (reads, pair) x (iterations + 1), final_reads, final_pair without
// synchronising; blocks without work exit at once.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "pd_em_common.cuh"

namespace {

constexpr int E2_SUB = 2048;                       // pairs per sort group
constexpr int E2_WBS = E2_SUB / 32;                // warp blocks per sort group
constexpr int E2_T = 256;
constexpr uint32_t E2_LMASK = (1u << 30) - 1;
enum { E2_NONE = 0, E2_M1 = 1, E2_M2 = 2, E2_FIN = 3 };
enum { PH_A = 0, PH_M1 = 1, PH_M2 = 2, PH_ALIVE = 3, PH_DEAD = 4, PH_DONE = 5 };
constexpr uint32_t E2_SUPP_CAP = 1536;             // supporting read pairs kept per pair for the percentiles

struct E2Item { uint32_t off; uint32_t n; int32_t dmx; uint32_t pair; };             // sorted-major, static
struct E2Ctl { int32_t shift; uint32_t lmode; int32_t supp_lo, supp_hi; };           // sorted-major, written by k_e2_pair
struct E2Cur { int32_t Lc, Sc, LcA, pad; };                                          // what rec / recA were computed with
struct E2Rec { double l0, l1, l2, sr, srd; uint32_t nd, pad; };                      // pair-major [pair][rg]
struct E2Stat { int32_t sd; uint32_t c; };                                           // sum of deviations, window count (:93-133)
struct E2Fin { double t0, t1, t2; uint32_t lad[3], dad[5], fl_min, fl_max, nsupp, pad; };
struct E2Pair {
    int32_t L0; uint32_t len, it, prev_len, phase, nvisited, src_a, pad;
    double freq, prev_freq, lr_conv, gt[3];
    int32_t vlen[64]; double vfreq[64];
};

struct E2Args {
    E2Item * item; E2Ctl * ctl; E2Cur * cur;         // [groups * E2_SUB]
    E2Rec * rec, * recA; E2Stat * stat; uint32_t * inv; E2Fin * fin;      // [pairs * R]
    E2Pair * pst;                                    // [pairs]
    uint32_t * blk_off; uint32_t * blk_nmax;         // [groups * E2_WBS]
    int32_t * devT; uint32_t * posT; uint32_t devt_cap;      // lane-interleaved copies of the items' read pairs
    uint32_t * devt_used;                            // device counter (words)
    uint32_t * ovf;                                  // set when devT is too small (the scan is repeated with more room)
    uint32_t * suppn; uint32_t * supp_first, * supp_last;     // [pairs], [pairs * E2_SUPP_CAP]
    uint32_t nsub;
};

__device__ __forceinline__ void st_release_gpu(uint32_t * p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(E2_T) k_e2_prep(PdDev a, EmArgs e, E2Args x)
{
    __shared__ uint32_t s_hist[264], s_start[264];
    __shared__ uint16_t s_key[E2_SUB], s_rank[E2_SUB];
    __shared__ uint32_t s_n[E2_SUB];
    __shared__ uint32_t s_nmax[E2_WBS], s_boff[E2_WBS];
    __shared__ uint32_t s_base;
    __shared__ unsigned long long s_ws[33];
    const uint32_t tid = threadIdx.x, sub = blockIdx.x, g = blockIdx.y;
    const uint32_t p_lo = sub * E2_SUB, np = min((uint32_t)E2_SUB, e.npairs - p_lo);
    const uint32_t grp = g * x.nsub + sub;
    const size_t sbase = (size_t)grp * E2_SUB;
    const PdRgConst * rg = a.rgc + g;
    const uint32_t max_load = __ldg(&rg->max_load), smp = __ldg(&rg->sample);
    const int hb = __ldg(&rg->hist_base);
    for (uint32_t i = tid; i < 264; i += E2_T) s_hist[i] = 0;
    for (uint32_t i = tid; i < (uint32_t)E2_SUB; i += E2_T) s_n[i] = 0;
    __syncthreads();
    // ---- keys
    for (uint32_t i = tid; i < np; i += E2_T) {
        const PdPair pr = e.pairs[e.pair0 + p_lo + i];
        const uint32_t job = pr.job - e.job_base;
        const uint32_t cj = e.cjob_of[job] - e.cj_base;
        const uint32_t n = e.act_cnt[(size_t)cj * a.R + g];
        const int32_t dmx = e.dmax[(size_t)job * a.N + smp];
        uint32_t key;
        if (n == 0 || n >= max_load) key = 256;
        else key = (dmx >= pr.L0 - hb + 1 ? 0u : 128u) + (127u - min(n, 127u));
        s_key[i] = (uint16_t)key;
        s_rank[i] = (uint16_t)atomicAdd(&s_hist[key], 1u);
        if (g == 0) {                                            // per-pair state is initialised by the blocks of read group 0
            E2Pair & st = x.pst[p_lo + i];
            st.L0 = pr.L0; st.len = (uint32_t)pr.L0; st.it = 0; st.prev_len = (uint32_t)pr.L0; st.phase = PH_A; st.nvisited = 0; st.src_a = 0;
            st.freq = 0; st.prev_freq = 0; st.lr_conv = 0;
        }
    }
    __syncthreads();
    {
        static_assert(E2_T == 256, "one thread per key bin");
        unsigned long long total;
        const unsigned long long ex = block_excl_scan((unsigned long long)s_hist[tid], s_ws, total);
        s_start[tid] = (uint32_t)ex;
        if (tid == 0) s_start[256] = (uint32_t)total;                       // empty / high-coverage items go last
    }
    __syncthreads();
    // ---- scatter into sorted order
    for (uint32_t i = tid; i < np; i += E2_T) {
        const uint32_t p = p_lo + i;
        const PdPair pr = e.pairs[e.pair0 + p];
        const uint32_t job = pr.job - e.job_base;
        const uint32_t cj = e.cjob_of[job] - e.cj_base;
        const uint32_t key = s_key[i];
        const uint32_t slot = s_start[key] + s_rank[i];
        uint32_t n = e.act_cnt[(size_t)cj * a.R + g];
        if (key == 256) n = 0;
        E2Item it;
        it.off = e.act_off[(size_t)cj * a.R + g]; it.n = n; it.dmx = e.dmax[(size_t)job * a.N + smp]; it.pair = p;
        x.item[sbase + slot] = it;
        x.ctl[sbase + slot] = E2Ctl{0, (uint32_t)pr.L0 & E2_LMASK, 0, 0};
        x.cur[sbase + slot] = E2Cur{INT_MIN, 0, pr.L0, 0};
        x.inv[(size_t)p * a.R + g] = (uint32_t)(sbase + slot);
        s_n[slot] = n;
    }
    for (uint32_t i = np + tid; i < (uint32_t)E2_SUB; i += E2_T) {              // padding slots of the last group
        x.item[sbase + i] = E2Item{0, 0, INT_MIN, 0xFFFFFFFFu};
        x.ctl[sbase + i] = E2Ctl{0, 0, 0, 0};
    }
    __syncthreads();
    // ---- warp blocks: longest item, slab offsets
    if (tid < (uint32_t)E2_WBS) {
        uint32_t m = 0;
        for (int k = 0; k < 32; ++k) m = max(m, s_n[tid * 32 + k]);
        s_nmax[tid] = m;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < E2_WBS; ++w) { s_boff[w] = tot; tot += 32u * s_nmax[w]; }
        uint32_t base = tot ? atomicAdd(x.devt_used, tot) : 0u;
        if ((uint64_t)base + tot > x.devt_cap) { atomicExch(x.ovf, 1u); base = 0xFFFFFFFFu; }
        s_base = base;
    }
    __syncthreads();
    const uint32_t base = s_base;
    if (tid < (uint32_t)E2_WBS) {
        x.blk_off[(size_t)grp * E2_WBS + tid] = base == 0xFFFFFFFFu ? 0u : base + s_boff[tid];
        x.blk_nmax[(size_t)grp * E2_WBS + tid] = base == 0xFFFFFFFFu ? 0u : s_nmax[tid];
    }
    if (base == 0xFFFFFFFFu) return;
    // ---- copy the read pairs: lane = item, [j][lane]
    const uint32_t lane = tid & 31, warp = tid >> 5;
    for (uint32_t wb = warp; wb < (uint32_t)E2_WBS; wb += E2_T / 32) {
        if (s_nmax[wb] == 0) continue;
        const uint32_t slot = wb * 32 + lane;
        const uint32_t n = s_n[slot];
        if (n == 0) continue;
        const E2Item it = x.item[sbase + slot];
        const size_t o = (size_t)base + s_boff[wb] + lane;
        for (uint32_t j = 0; j < n; ++j) {
            x.devT[o + (size_t)j * 32] = __ldg(e.pool_dev + it.off + j);
            x.posT[o + (size_t)j * 32] = __ldg(e.pool_pos + it.off + j);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// data-likelihood pass
// ------------------------------------------------------------------------------------------------------------------
template <bool PASS_A>
__global__ void __launch_bounds__(E2_T) k_e2_reads(PdDev a, EmArgs e, E2Args x)
{
    extern __shared__ double2 s_tab[];                   // [rows] {ln, lnp}, then [rows] {val, fr}
    const uint32_t tid = threadIdx.x, sub = blockIdx.x, g = blockIdx.y;
    const uint32_t grp = g * x.nsub + sub;
    const uint32_t * nmaxs = x.blk_nmax + (size_t)grp * E2_WBS;
    if (nmaxs[0] == 0) return;                           // sorted: an empty first warp block means an empty group
    if (*(volatile uint32_t *)x.ovf) return;
    const PdRgConst * rg = a.rgc + g;
    const uint32_t hist_len = __ldg(&rg->hist_len), rows = hist_len + 1;
    const int hb = __ldg(&rg->hist_base);
    const double minp = __ldg(&rg->min_prob), lnminp = __ldg(&rg->ln_min_prob);
    double2 * sA = s_tab, * sB = s_tab + rows;
    {
        const PdTab * t = a.tab + __ldg(&rg->hist_off);
        for (uint32_t r = tid; r < rows; r += E2_T) {
            const D4 v = ld4(&t[r].val);
            sA[r] = make_double2(v.b, v.c);
            sB[r] = make_double2(v.a, v.d);
        }
    }
    __syncthreads();
    const uint32_t hl2 = hist_len - 2u;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const size_t sbase = (size_t)grp * E2_SUB;
    const double sd = __ldg(&rg->stddev);
    for (uint32_t wb = warp; wb < (uint32_t)E2_WBS; wb += E2_T / 32) {
        if (nmaxs[wb] == 0) break;
        const size_t slot = sbase + wb * 32 + lane;
        const E2Item it = x.item[slot];
        const uint2 cw = *reinterpret_cast<const uint2 *>(&x.ctl[slot]);
        const uint32_t mode = PASS_A ? (uint32_t)E2_M2 : (cw.y >> 30);
        const int L = (int)(cw.y & E2_LMASK);
        int S = 0;
        bool need = it.n > 0;
        E2Cur cur = E2Cur{INT_MIN, 0, L, 0};
        if (!PASS_A) {
            cur = x.cur[slot];
            const int thr = hb - 1;                                            // carrier of length X: dmx >= X - hb + 1
            if (mode == E2_M1) {
                S = (int)cw.x;
                need = need && (cur.Lc == INT_MIN || S != cur.Sc || (L != cur.Lc && (it.dmx + thr >= L || it.dmx + thr >= cur.Lc)));
            } else if (mode == E2_M2) {
                need = need && (L != cur.LcA && (it.dmx + thr >= L || it.dmx + thr >= cur.LcA));
            } else need = false;
            if (mode == E2_M1 || mode == E2_M2) x.ctl[slot].lmode = cw.y & E2_LMASK;      // request consumed
        }
        uint32_t jmax = need ? it.n : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jmax = max(jmax, __shfl_xor_sync(PD_FULL, jmax, o));
        if (jmax == 0) continue;
        const int32_t * dp = x.devT + x.blk_off[(size_t)grp * E2_WBS + wb] + lane;
        const uint32_t ub = (uint32_t)(hb - 1 - S), vb = (uint32_t)(hb - 1 - L);
        const uint32_t n = need ? it.n : 0u;
        double l0 = 0, l1 = 0, l2 = 0, sr = 0, srd = 0;
        uint32_t nd = 0, c = 0; int32_t sdv = 0;
        int wlo = 0, whi = 0;
        if (PASS_A) { wlo = max(L / 2, (int)floor((double)L - 2 * sd + 0.5)); whi = (int)((double)L + 2 * sd); }
        for (uint32_t j = 0; j < jmax; ++j) {
            if (j < n) {
                const int d = dp[(size_t)j * 32];
                const uint32_t u = (uint32_t)d + ub, v = (uint32_t)d + vb;
                const uint32_t ir = u < hl2 ? u + 2u : 0u;
                const double2 A = sA[ir], B = sB[ir];
                double g1, r;
                if (v < hl2) {
                    const double2 dA = sA[v + 2u], dB = sB[v + 2u];
                    if (B.x == dB.x) { g1 = A.x; ++nd; r = 0.5; }                  // ref == del: + LN2_RESIDUE in finish_triple
                    else if (dB.x == minp) { g1 = A.y; r = B.y; }
                    else if (B.x == minp) { g1 = dA.y; r = 1.0 - dB.y; }
                    else { g1 = log(B.x + dB.x) - LN2_D; r = dB.x / (dB.x + B.x); }
                    l2 += dA.x;
                } else {                                                          // deletion hypothesis on the floor: the bulk
                    if (B.x == minp) { g1 = A.x; ++nd; r = 0.5; }
                    else { g1 = A.y; r = B.y; }
                    l2 += lnminp;
                }
                l0 += A.x; l1 += g1; sr += r; srd += r * d;
                if (PASS_A) { c += (d > wlo && d < whi); sdv += d; }
            }
        }
        if (need) {
            E2Rec * out = ((PASS_A || mode == E2_M2) ? x.recA : x.rec) + (size_t)it.pair * a.R + g;
            E2Rec r; r.l0 = l0; r.l1 = l1; r.l2 = l2; r.sr = sr; r.srd = srd; r.nd = nd; r.pad = 0;
            *out = r;
            if (PASS_A) x.stat[(size_t)it.pair * a.R + g] = E2Stat{sdv, c};
            else {
                if (mode == E2_M1) { cur.Lc = L; cur.Sc = S; } else cur.LcA = L;
                x.cur[slot] = cur;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// pair phase
// ------------------------------------------------------------------------------------------------------------------
struct PairShared {
    double red[2][64];
    unsigned long long redu[64];
    double rgw[3];
    int flag;
};

// triple of sample s from the records `src` of pair p; also the per-read-group constants the length update needs
struct SampleDl { double x0, E0, E1, E2; };

__device__ __forceinline__ SampleDl sample_dl(const PdDev & a, const uint32_t * __restrict__ cnt, const E2Rec * __restrict__ src, uint32_t s)
{
    double l0 = 0, l1 = 0, l2 = 0; uint32_t nd = 0;
    const uint32_t g0 = a.sample_rg[s], g1 = a.sample_rg[s + 1];
    for (uint32_t g = g0; g < g1; ++g) {
        const uint32_t n = cnt[g];
        if (n == 0 || n >= __ldg(&a.rgc[g].max_load)) continue;
        const E2Rec r = src[g];
        l0 += r.l0; l1 += r.l1; l2 += r.l2; nd += r.nd;
    }
    SampleDl o; double x1, x2;
    finish_triple(l0, l1, l2, nd, o.x0, x1, x2);
    o.E0 = o.x0 == 0 ? 1.0 : exp(o.x0); o.E1 = x1 == 0 ? 1.0 : exp(x1); o.E2 = x2 == 0 ? 1.0 : exp(x2);
    return o;
}

template <bool ONE>
__global__ void __launch_bounds__(1024) k_e2_pair(PdDev a, EmArgs e, E2Args x, int round_no)
{
    __shared__ PairShared sh;
    const uint32_t p = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    E2Pair & st = x.pst[p];
    const uint32_t phase = st.phase;
    if (phase >= PH_ALIVE) return;
    if (*(volatile uint32_t *)x.ovf) return;
    const PdPair pr = e.pairs[e.pair0 + p];
    const uint32_t job = pr.job - e.job_base;
    const uint32_t cj = e.cjob_of[job] - e.cj_base;
    const uint32_t * cnt = e.act_cnt + (size_t)cj * a.R;
    const E2Rec * src = (phase == PH_M1 ? x.rec : x.recA) + (size_t)p * a.R;
    const E2Rec * rec1 = x.rec + (size_t)p * a.R;
    const E2Stat * stat = x.stat + (size_t)p * a.R;
    const uint32_t * inv = x.inv + (size_t)p * a.R;
    int par = 0;
    uint32_t len = st.len, it = st.it;
    double freq = st.freq;
    Gt gt = Gt{st.gt[0], st.gt[1], st.gt[2]};
    SampleDl mine = SampleDl{0, 1, 1, 1};
    if (ONE) { if (tid < a.N) mine = sample_dl(a, cnt, src, tid); }
    auto dl_of = [&](uint32_t s) -> SampleDl { return ONE ? mine : sample_dl(a, cnt, src, s); };
    if (tid == T - 1) {
        // likelihood triple of read group 0 (rgDlIt is never advanced, :401,424-431: it drives every reference shift)
        const uint32_t n0 = cnt[0];
        if (n0 >= __ldg(&a.rgc[0].max_load)) sh.rgw[0] = sh.rgw[1] = sh.rgw[2] = 0;      // Triple(0,0,0) in the reference
        else {
            double l0 = 0, l1 = 0, l2 = 0;
            if (n0) { const E2Rec r = src[0]; l0 = r.l0; l1 = r.l1; l2 = r.l2; }
            const double m = fmax(fmax(l0, l1), l2);
            sh.rgw[0] = exp(l0 - m); sh.rgw[1] = exp(l1 - m); sh.rgw[2] = exp(l2 - m);
        }
    }
    auto lr_now = [&](const Gt g) {                      // deletion_likelihood_ratio :490-508
        double del = 0, nodel = 0;
        for (uint32_t s = tid; s < a.N; s += T) {
            const SampleDl d = dl_of(s);
            const double p0 = d.E0 * g.a, p1 = d.E1 * g.b, p2 = d.E2 * g.c, pAll = p0 + p1 + p2;
            del += log(p0 / pAll * d.E0 + p1 / pAll * d.E1 + p2 / pAll * d.E2);
            nodel += d.x0;
        }
        block_sum2(del, nodel, sh.red, par);
        return del - nodel;
    };
    auto finish = [&](bool alive, uint32_t src_a, uint32_t reason) {
        // request the final pass (alive) or publish the rejection
        if (alive) {
            for (uint32_t s = tid; s < a.N; s += T) {
                const uint32_t g0 = a.sample_rg[s], g1 = a.sample_rg[s + 1];
                int lo = INT_MAX, hi = 0;                // borders of the LAST usable read group apply to the whole sample (quirk)
                for (uint32_t g = g0; g < g1; ++g) {
                    if (cnt[g] >= __ldg(&a.rgc[g].max_load)) continue;
                    lo = (int)len - __ldg(&a.rgc[g].lower_q); hi = (int)len + __ldg(&a.rgc[g].upper_q);
                }
                for (uint32_t g = g0; g < g1; ++g) {
                    E2Ctl * c = x.ctl + inv[g];
                    if (src_a) c->shift = 0;
                    c->supp_lo = lo; c->supp_hi = hi;
                    c->lmode = (len & E2_LMASK) | ((uint32_t)E2_FIN << 30);
                }
            }
        }
        if (tid == 0) {
            st.len = len; st.it = it; st.freq = freq; st.gt[0] = gt.a; st.gt[1] = gt.b; st.gt[2] = gt.c;
            st.src_a = src_a; st.phase = alive ? PH_ALIVE : PH_DEAD;
            if (!alive) {
                e.valid[p] = 0;
                if (e.dbg) { e.dbg[4 * p] = reason; e.dbg[4 * p + 1] = len; e.dbg[4 * p + 2] = it; }
                __threadfence();
                st_release_gpu(e.done + p, 1u);
            }
        }
    };
    __syncthreads();                                     // sh.rgw of this round

    if (phase == PH_A) {
        // ---- initialize_allele_frequency :93-133
        unsigned long long c = 0, t = 0;
        for (uint32_t s = tid; s < a.N; s += T)
            for (uint32_t g = a.sample_rg[s]; g < a.sample_rg[s + 1]; ++g) {
                const uint32_t n = cnt[g];
                if (n == 0 || n >= __ldg(&a.rgc[g].max_load)) continue;
                t += n; c += stat[g].c;
            }
        block_sum2u(c, t, sh.redu);
        freq = t == 0 ? 0.0 : (double)c / (double)t;
        gt = gt_prior(freq, e.somatic);
        if (freq == 0) { finish(false, 0, 1); return; }
    } else if (phase == PH_M1) {
        // ---- update_allele_frequency :467-485 (priors of the previous iteration)
        double fs = 0, dummy = 0;
        for (uint32_t s = tid; s < a.N; s += T) {
            const SampleDl d = dl_of(s);
            const double p0 = d.E0 * gt.a, p1 = d.E1 * gt.b, p2 = d.E2 * gt.c;
            fs += (p1 + 2 * p2) / (p0 + p1 + p2);
        }
        block_sum2(fs, dummy, sh.red, par);
        freq = fs / 2.0 / a.N;
        if (freq == 0) { finish(false, 0, 2); return; }
        gt = gt_prior(freq, e.somatic);
        bool conv = false;
        for (uint32_t i = 0; i < st.nvisited; ++i)
            if (st.vlen[i] == (int)len && fabs(st.vfreq[i] - freq) <= 0.0001) conv = true;
        if (conv) {
            // convergence :632-658: compare with the previous estimate evaluated with the initial (zero) shifts
            const double lr = lr_now(gt);
            const uint32_t lm = (st.prev_len & E2_LMASK) | ((uint32_t)E2_M2 << 30);
            for (uint32_t g = tid; g < a.R; g += T) x.ctl[inv[g]].lmode = lm;
            if (tid == 0) {
                st.lr_conv = lr; st.freq = freq; st.gt[0] = gt.a; st.gt[1] = gt.b; st.gt[2] = gt.c; st.phase = PH_M2;
            }
            return;
        }
    } else {                                             // PH_M2: the records in recA hold the previous estimate
        const double plr = lr_now(gt_prior(st.prev_freq, e.somatic));
        uint32_t src_a = 0;
        if (plr > st.lr_conv) { len = st.prev_len; freq = st.prev_freq; src_a = 1; }
        const bool alive = !(freq < 0.0000000001 || len < e.min_len);
        finish(alive, src_a, 2);
        return;
    }
    if (!(len >= e.min_len && it < e.iterations)) {
        const bool alive = !(freq < 0.0000000001 || len < e.min_len);
        finish(alive, phase == PH_A ? 1u : 0u, 2);           // without an iteration the records of pass A are the final state
        return;
    }
    // ---- next iteration: update_deletion_length :388-462 from the moments of the posterior weights
    ++it;
    const uint32_t prevLen = len; const double prevFreq = freq;
    const double r0 = sh.rgw[0], r1 = sh.rgw[1], r2 = sh.rgw[2];
    const double aSumRg = r0 * gt.a + r1 * gt.b + r2 * gt.c;
    const double ea0Rg = r0 * gt.a / aSumRg, ea1Rg = r1 * gt.b / aSumRg;          // NaN when read group 0 is high-coverage
    double sumDel = 0, wDel = 0;
    for (uint32_t s = tid; s < a.N; s += T) {
        const SampleDl d = dl_of(s);
        const double invp = 1.0 / (d.E0 * gt.a + d.E1 * gt.b + d.E2 * gt.c);
        const double ea1 = d.E1 * gt.b * invp, ea2 = d.E2 * gt.c * invp;
        for (uint32_t g = a.sample_rg[s]; g < a.sample_rg[s + 1]; ++g) {
            const uint32_t n = cnt[g];
            if (n >= __ldg(&a.rgc[g].max_load)) continue;
            double Sr = 0, Srd = 0, Sd = 0; const double dn = (double)n;
            if (n) { const E2Rec r = src[g]; Sr = r.sr; Srd = r.srd; Sd = (double)stat[g].sd; }
            sumDel += ea1 * Sr + ea2 * dn; wDel += ea1 * Srd + ea2 * Sd;
            const double sumRef = ea1Rg * (dn - Sr) + ea0Rg * dn, wRef = ea1Rg * (Sd - Srd) + ea0Rg * Sd;
            const double q = wRef / sumRef;
            int sft = (q != q) ? 0 : (q >= 2147483647.0 ? INT_MAX : (q <= -2147483648.0 ? INT_MIN : (int)q));
            const double sdg = __ldg(&a.rgc[g].stddev);
            if (sft > sdg || sft < -1 * sdg) sft = 0;
            x.ctl[inv[g]].shift = sft;
        }
    }
    block_sum2(sumDel, wDel, sh.red, par);
    if (sumDel == 0) len = 0;
    else { const double nlen = wDel / sumDel; len = nlen < 0 ? 0u : (uint32_t)round(nlen); }
    const uint32_t lm = (len & E2_LMASK) | ((uint32_t)E2_M1 << 30);
    for (uint32_t g = tid; g < a.R; g += T) x.ctl[inv[g]].lmode = lm;
    (void)rec1;
    if (tid == 0) {
        int f = -1;                                      // visited[prevLen] = prevFreq (:600)
        const int nv = (int)st.nvisited;
        for (int i = 0; i < nv; ++i) if (st.vlen[i] == (int)prevLen) f = i;
        if (f < 0) { f = nv; st.vlen[f] = (int)prevLen; st.nvisited = (uint32_t)nv + 1; }
        st.vfreq[f] = prevFreq;
        st.len = len; st.it = it; st.prev_len = prevLen; st.freq = freq; st.prev_freq = prevFreq;
        st.gt[0] = gt.a; st.gt[1] = gt.b; st.gt[2] = gt.c; st.phase = PH_M1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// final pass
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(E2_T) k_e2_final_reads(PdDev a, EmArgs e, E2Args x)
{
    extern __shared__ double2 s_tab[];                   // [rows] {log10, log10p}, then [rows] val (double)
    const uint32_t tid = threadIdx.x, sub = blockIdx.x, g = blockIdx.y;
    const uint32_t grp = g * x.nsub + sub;
    const uint32_t * nmaxs = x.blk_nmax + (size_t)grp * E2_WBS;
    if (nmaxs[0] == 0) return;
    if (*(volatile uint32_t *)x.ovf) return;
    const PdRgConst * rg = a.rgc + g;
    const uint32_t hist_len = __ldg(&rg->hist_len), rows = hist_len + 1;
    const int hb = __ldg(&rg->hist_base);
    const double minp = __ldg(&rg->min_prob), l10minp = __ldg(&rg->l10_min_prob);
    const int lower_q = __ldg(&rg->lower_q), upper_q = __ldg(&rg->upper_q), inner_off = __ldg(&rg->inner_off);
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const size_t sbase = (size_t)grp * E2_SUB;
    // any item of this group in a surviving pair? (cheap test before the table is staged)
    {
        int any = 0;
        for (uint32_t i = tid; i < (uint32_t)E2_SUB; i += E2_T) any |= (x.ctl[sbase + i].lmode >> 30) == E2_FIN && x.item[sbase + i].n > 0;
        if (!__syncthreads_or(any)) return;
    }
    double2 * sC = s_tab; double * sV = reinterpret_cast<double *>(s_tab + rows);
    {
        const PdTab * t = a.tab + __ldg(&rg->hist_off);
        for (uint32_t r = tid; r < rows; r += E2_T) {
            sC[r] = ld2(&t[r].l10);
            sV[r] = __ldg(&t[r].val);
        }
    }
    __syncthreads();
    const uint32_t hl2 = hist_len - 2u;
    for (uint32_t wb = warp; wb < (uint32_t)E2_WBS; wb += E2_T / 32) {
        if (nmaxs[wb] == 0) break;
        const size_t slot = sbase + wb * 32 + lane;
        const E2Item it = x.item[slot];
        const E2Ctl cw = x.ctl[slot];
        const bool need = (cw.lmode >> 30) == E2_FIN && it.n > 0;
        uint32_t jmax = need ? it.n : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jmax = max(jmax, __shfl_xor_sync(PD_FULL, jmax, o));
        if (jmax == 0) continue;
        const int flen = (int)(cw.lmode & E2_LMASK);
        const size_t bo = (size_t)x.blk_off[(size_t)grp * E2_WBS + wb] + lane;
        const uint32_t ub = (uint32_t)(hb - 1 - cw.shift), vb = (uint32_t)(hb - 1 - flen);
        const uint32_t n = need ? it.n : 0u;
        const int delLower = flen - lower_q, delUpper = flen + upper_q;           // DAD: the read group's own borders
        double t0 = 0, t1 = 0, t2 = 0;
        uint32_t lad0 = 0, lad1 = 0, lad2 = 0, dad0 = 0, dad1 = 0, dad2 = 0, dad3 = 0, dad4 = 0, fl_min = 0xFFFFFFFFu, fl_max = 0, nsupp = 0;
        for (uint32_t j = 0; j < jmax; ++j) {
            if (j < n) {
                const int d = x.devT[bo + (size_t)j * 32];
                if (d > upper_q) { if (d < delLower) ++dad2; else if (d <= delUpper) ++dad3; else ++dad4; }
                else { if (d < delUpper) ++dad0; else ++dad1; }
                const uint32_t u = (uint32_t)d + ub, v = (uint32_t)d + vb;
                const uint32_t ir = u < hl2 ? u + 2u : 0u;
                const double2 rc = sC[ir]; const double ref = sV[ir];
                double del = minp; double2 dc = make_double2(l10minp, 0.0);
                if (v < hl2) { dc = sC[v + 2u]; del = sV[v + 2u]; }
                if (ref >= 2 * del) ++lad0; else if (del >= 2 * ref) ++lad2; else ++lad1;
                t0 += rc.x; t2 += dc.x;
                if (ref == del) t1 += rc.x;                                       // residue applied in k_e2_final_pair
                else if (del == minp) t1 += rc.y;
                else if (ref == minp) t1 += dc.y;
                else t1 += log10(ref + del) - LOG10_2_D;
                const uint32_t first = x.posT[bo + (size_t)j * 32] + e.anchor;
                const uint32_t last = first + (uint32_t)max(0, d + inner_off);
                fl_min = min(fl_min, first); fl_max = max(fl_max, last);
                if (d >= cw.supp_lo && d <= cw.supp_hi) {                         // supporting read pair
                    ++nsupp;
                    const uint32_t k = atomicAdd(x.suppn + it.pair, 1u);
                    if (k < E2_SUPP_CAP) { x.supp_first[(size_t)it.pair * E2_SUPP_CAP + k] = first; x.supp_last[(size_t)it.pair * E2_SUPP_CAP + k] = last; }
                }
            }
        }
        if (need) {
            E2Fin f;
            f.t0 = t0; f.t1 = t1; f.t2 = t2; f.lad[0] = lad0; f.lad[1] = lad1; f.lad[2] = lad2;
            f.dad[0] = dad0; f.dad[1] = dad1; f.dad[2] = dad2; f.dad[3] = dad3; f.dad[4] = dad4;
            f.fl_min = fl_min; f.fl_max = fl_max; f.nsupp = nsupp; f.pad = 0;
            x.fin[(size_t)it.pair * a.R + g] = f;
        }
    }
}

template <bool ONE>
__global__ void __launch_bounds__(1024) k_e2_final_pair(PdDev a, EmArgs e, E2Args x)
{
    __shared__ PairShared sh;
    __shared__ uint32_t s_first[E2_SUPP_CAP], s_last[E2_SUPP_CAP];
    __shared__ uint32_t s_sel[2];
    const uint32_t p = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    E2Pair & st = x.pst[p];
    if (st.phase != PH_ALIVE) return;
    if (*(volatile uint32_t *)x.ovf) return;
    const PdPair pr = e.pairs[e.pair0 + p];
    const uint32_t job = pr.job - e.job_base;
    const uint32_t w = e.job_window[pr.job];
    const uint32_t cj = e.cjob_of[job] - e.cj_base;
    const uint32_t * cnt = e.act_cnt + (size_t)cj * a.R;
    const uint32_t * off = e.act_off + (size_t)cj * a.R;
    const E2Rec * src = (st.src_a ? x.recA : x.rec) + (size_t)p * a.R;
    const E2Fin * fin = x.fin + (size_t)p * a.R;
    const uint8_t * sstat = e.sstat + (size_t)job * a.N;
    uint32_t * ps = e.ps + (size_t)p * 13 * a.N;
    const int len = (int)st.len;
    const Gt gt = Gt{st.gt[0], st.gt[1], st.gt[2]};
    int par = 0;
    auto reject = [&](uint32_t reason) {
        if (tid == 0) {
            e.valid[p] = 0; st.phase = PH_DONE;
            if (e.dbg) { e.dbg[4 * p] = reason; e.dbg[4 * p + 1] = st.len; e.dbg[4 * p + 2] = st.it; }
            __threadfence();
            st_release_gpu(e.done + p, 1u);
        }
    };
    unsigned long long supp = 0, ndata = 0;
    double del = 0, nodel = 0;
    for (uint32_t s = tid; s < a.N; s += T) {
        uint32_t lad0 = 0, lad1 = 0, lad2 = 0, dad0 = 0, dad1 = 0, dad2 = 0, dad3 = 0, dad4 = 0, fl_min = 0xFFFFFFFFu, fl_max = 0, ndeg = 0;
        double l0 = 0, l1 = 0, l2 = 0, t0 = 0, t1 = 0, t2 = 0;
        for (uint32_t g = a.sample_rg[s]; g < a.sample_rg[s + 1]; ++g) {
            const uint32_t n = cnt[g];
            if (n == 0 || n >= __ldg(&a.rgc[g].max_load)) continue;
            const E2Rec r = src[g]; const E2Fin f = fin[g];
            l0 += r.l0; l1 += r.l1; l2 += r.l2; ndeg += r.nd;
            t0 += f.t0; t1 += f.t1; t2 += f.t2;
            lad0 += f.lad[0]; lad1 += f.lad[1]; lad2 += f.lad[2];
            dad0 += f.dad[0]; dad1 += f.dad[1]; dad2 += f.dad[2]; dad3 += f.dad[3]; dad4 += f.dad[4];
            fl_min = min(fl_min, f.fl_min); fl_max = max(fl_max, f.fl_max);
            supp += f.nsupp;
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
        const double E0 = exp(x0), E1 = exp(x1), E2 = exp(x2);
        // calculatePhredGL utils_popdel.h:1511-1528
        const double gTot = log10(exp(g0l) + exp(g1l) + exp(g2l));
        const double q0 = -10 * (g0l - gTot), q1 = -10 * (g1l - gTot), q2 = -10 * (g2l - gTot);
        const double mn = fmin(fmin(q0, q1), q2);
        uint32_t * o = ps + 13 * s;
        const bool low = sstat[s] == 0;
        o[0] = low ? 0u : (uint32_t)round(q0 - mn);
        o[1] = low ? 0u : (uint32_t)round(q1 - mn);
        o[2] = low ? 0u : (uint32_t)round(q2 - mn);
        o[3] = lad0; o[4] = lad1; o[5] = lad2;
        o[6] = dad0; o[7] = dad1; o[8] = dad2; o[9] = dad3; o[10] = dad4;
        o[11] = fl_min; o[12] = fl_max;
        if (!low) ++ndata;
        // deletion_likelihood_ratio :490-508
        const double p0 = E0 * gt.a, p1 = E1 * gt.b, p2 = E2 * gt.c, pAll = p0 + p1 + p2;
        del += log(p0 / pAll * E0 + p1 / pAll * E1 + p2 / pAll * E2);
        nodel += x0;
    }
    block_sum2u(supp, ndata, sh.redu);
    if (supp == 0) { reject(3); return; }
    // percentiles of the supporting starts (80th) and ends (20th): getSuppFirstLast :514-529, by value bisection
    const unsigned long long kF = (unsigned long long)round((double)(supp - 1) * 0.8);
    const unsigned long long kL = (unsigned long long)round((double)(supp - 1) * (1 - 0.8));
    uint32_t sF, sL;
    if (supp <= E2_SUPP_CAP) {
        const uint32_t ns = (uint32_t)supp;
        for (uint32_t i = tid; i < ns; i += T) { s_first[i] = x.supp_first[(size_t)p * E2_SUPP_CAP + i]; s_last[i] = x.supp_last[(size_t)p * E2_SUPP_CAP + i]; }
        __syncthreads();
        if (tid < 32) {
            uint32_t loF = 0xFFFFFFFFu, hiF = 0, loL = 0xFFFFFFFFu, hiL = 0;
            for (uint32_t i = tid; i < ns; i += 32) { loF = min(loF, s_first[i]); hiF = max(hiF, s_first[i]); loL = min(loL, s_last[i]); hiL = max(hiL, s_last[i]); }
            for (int o = 16; o > 0; o >>= 1) {
                loF = min(loF, __shfl_xor_sync(PD_FULL, loF, o)); hiF = max(hiF, __shfl_xor_sync(PD_FULL, hiF, o));
                loL = min(loL, __shfl_xor_sync(PD_FULL, loL, o)); hiL = max(hiL, __shfl_xor_sync(PD_FULL, hiL, o));
            }
            while (loF < hiF || loL < hiL) {
                const uint32_t midF = loF + (hiF - loF) / 2, midL = loL + (hiL - loL) / 2;
                uint32_t cF = 0, cL = 0;
                for (uint32_t i = tid; i < ns; i += 32) { cF += s_first[i] <= midF; cL += s_last[i] <= midL; }
                for (int o = 16; o > 0; o >>= 1) { cF += __shfl_xor_sync(PD_FULL, cF, o); cL += __shfl_xor_sync(PD_FULL, cL, o); }
                if (loF < hiF) { if (cF >= kF + 1) hiF = midF; else loF = midF + 1; }
                if (loL < hiL) { if (cL >= kL + 1) hiL = midL; else loL = midL + 1; }
            }
            if (tid == 0) { s_sel[0] = loF; s_sel[1] = loL; }
        }
        __syncthreads();
        sF = s_sel[0]; sL = s_sel[1];
    } else {
        // more supporting read pairs than the list holds: bisect on the value, recounting from the pool each step
        uint32_t loF = 0, hiF = 0xFFFFFFFFu, loL = 0, hiL = 0xFFFFFFFFu;
        while (loF < hiF || loL < hiL) {
            const uint32_t midF = loF + (hiF - loF) / 2, midL = loL + (hiL - loL) / 2;
            unsigned long long cF = 0, cL = 0;
            for (uint32_t s = tid; s < a.N; s += T) {
                int delLower = INT_MAX, delUpper = 0;
                const uint32_t g0 = a.sample_rg[s], g1 = a.sample_rg[s + 1];
                for (uint32_t g = g0; g < g1; ++g) {
                    if (cnt[g] >= __ldg(&a.rgc[g].max_load)) continue;
                    delLower = len - __ldg(&a.rgc[g].lower_q); delUpper = len + __ldg(&a.rgc[g].upper_q);
                }
                for (uint32_t g = g0; g < g1; ++g) {
                    const uint32_t n = cnt[g];
                    if (n >= __ldg(&a.rgc[g].max_load)) continue;
                    const int inner_off = __ldg(&a.rgc[g].inner_off);
                    const uint32_t * pp = e.pool_pos + off[g];
                    const int32_t * pd = e.pool_dev + off[g];
                    for (uint32_t i = 0; i < n; ++i) {
                        const int d = pd[i];
                        if (d >= delLower && d <= delUpper) {
                            const uint32_t first = pp[i] + e.anchor;
                            const uint32_t last = first + (uint32_t)max(0, d + inner_off);
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
    block_sum2(del, nodel, sh.red, par);
    const double lr = del - nodel;
    if (tid == 0) {
        const bool ok = lr >= e.min_lr;
        e.valid[p] = ok ? 1 : 0;
        st.phase = PH_DONE;
        if (e.dbg) { e.dbg[4 * p] = ok ? 0 : 5; e.dbg[4 * p + 1] = st.len; e.dbg[4 * p + 2] = st.it; e.dbg[4 * p + 3] = (uint32_t)supp; }
        if (ok) {
            pd_call c;
            c.initial_length = (uint32_t)pr.L0; c.iterations = st.it; c.deletion_length = st.len;
            c.filter = ((double)ndata / a.N >= e.min_sample_fraction) ? 0u : 4u;
            c.lr = lr; c.frequency = st.freq;
            const uint32_t cur = e.anchor + w * PD_WIN;
            c.window_position = cur - 1;
            c.position = e.window_wise ? cur - 1 : sF;
            c.end_position = e.window_wise ? 0u : sL;
            c.segment = (uint32_t)(((uint64_t)w * PD_WIN) / a.window_buffer);
            e.calls[p] = c;
        }
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release_gpu(e.done + p, 1u); }
}

template <typename T>
int e2_grow(pd_ctx * c, int slot, T *& p, size_t count)
{
    void * q = nullptr;
    if (pd_grow_scratch(c, slot, std::max<size_t>(count, 1) * sizeof(T), &q)) return c->status;
    p = reinterpret_cast<T *>(q);
    return 0;
}

}  // namespace

// bytes of scratch one pair needs in the pipeline above (sizes the EM chunks in pd_scan.cu)
size_t pd_em2_pair_bytes(uint32_t R, double reads_per_pair)
{
    return (size_t)R * (3 * 16 + 2 * 48 + 8 + 4 + 72 + 8) + sizeof(E2Pair) + (size_t)E2_SUPP_CAP * 8 + (size_t)(reads_per_pair * 8 * 1.3) + 64;
}

bool pd_em2_usable(const pd_ctx * c)
{
    if (getenv("PD_EM_V1") || getenv("PD_EM_GENERAL")) return false;
    uint32_t rows = 0;
    for (const auto & k : c->rgc) rows = std::max(rows, k.hist_len + 1);
    return (size_t)rows * 32 <= 160 * 1024;
}

int pd_launch_em2(pd_ctx * c, const PdDev & a, const EmArgs & e, double reads_per_pair, cudaStream_t st, uint64_t * launches)
{
    const uint32_t np = e.npairs, R = a.R, N = a.N;
    if (np == 0) return 0;
    E2Args x;
    x.nsub = (np + E2_SUB - 1) / E2_SUB;
    const size_t groups = (size_t)R * x.nsub, slots = groups * E2_SUB, items = (size_t)np * R;
    if (e2_grow(c, PD_S_E2_ITEM, x.item, slots) || e2_grow(c, PD_S_E2_CTL, x.ctl, slots) || e2_grow(c, PD_S_E2_CUR, x.cur, slots) ||
        e2_grow(c, PD_S_E2_REC, x.rec, items) || e2_grow(c, PD_S_E2_RECA, x.recA, items) || e2_grow(c, PD_S_E2_STAT, x.stat, items) ||
        e2_grow(c, PD_S_E2_INV, x.inv, items) || e2_grow(c, PD_S_E2_FIN, x.fin, items) || e2_grow(c, PD_S_E2_PST, x.pst, (size_t)np) ||
        e2_grow(c, PD_S_E2_BOFF, x.blk_off, groups * E2_WBS) || e2_grow(c, PD_S_E2_BNMAX, x.blk_nmax, groups * E2_WBS) ||
        e2_grow(c, PD_S_E2_SUPPN, x.suppn, (size_t)np + 8) || e2_grow(c, PD_S_E2_SUPPF, x.supp_first, (size_t)np * E2_SUPP_CAP) ||
        e2_grow(c, PD_S_E2_SUPPL, x.supp_last, (size_t)np * E2_SUPP_CAP))
        return c->status;
    // lane-interleaved read-pair copies: capacity from the pool density of this batch, kept across scans; k_e2_prep reports
    // an overflow and the scan is repeated with the exact need (pd_scan.cu)
    const size_t want = (size_t)(reads_per_pair * np * 1.25) + groups * 2048 + (1u << 20);
    if (want > c->e2_devt_cap) c->e2_devt_cap = want;
    if (c->e2_devt_cap > 0xFFFFFF00ull) c->e2_devt_cap = 0xFFFFFF00ull;
    if (e2_grow(c, PD_S_E2_DEVT, x.devT, c->e2_devt_cap) || e2_grow(c, PD_S_E2_POST, x.posT, c->e2_devt_cap)) return c->status;
    x.devt_cap = (uint32_t)c->e2_devt_cap;
    uint32_t * cnts;
    if (e2_grow(c, PD_S_E2_CNT, cnts, (size_t)8)) return c->status;
    x.devt_used = cnts; x.ovf = cnts + 1;
    PD_CUDA(c, cudaMemsetAsync(x.devt_used, 0, 4, st));                     // (ovf is sticky for the scan: cleared by pd_run_scan)
    PD_CUDA(c, cudaMemsetAsync(x.suppn, 0, (size_t)np * 4, st));

    uint32_t rows = 0;
    for (const auto & k : c->rgc) rows = std::max(rows, k.hist_len + 1);
    const size_t smem_em = (size_t)rows * 32, smem_fin = (size_t)rows * 24;
    if (smem_em > 48 * 1024) {
        PD_CUDA(c, cudaFuncSetAttribute(k_e2_reads<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_em));
        PD_CUDA(c, cudaFuncSetAttribute(k_e2_reads<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_em));
    }
    if (smem_fin > 48 * 1024) PD_CUDA(c, cudaFuncSetAttribute(k_e2_final_reads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fin));
    const dim3 gg(x.nsub, R);
    const bool one = N <= 1024;
    const uint32_t TP = one ? std::max<uint32_t>(32, ((N + 31) / 32) * 32) : 1024;
    k_e2_prep<<<gg, E2_T, 0, st>>>(a, e, x);
    k_e2_reads<true><<<gg, E2_T, smem_em, st>>>(a, e, x);
    if (one) k_e2_pair<true><<<np, TP, 0, st>>>(a, e, x, 0); else k_e2_pair<false><<<np, TP, 0, st>>>(a, e, x, 0);
    *launches += 3;
    for (uint32_t r = 0; r < e.iterations + 1; ++r) {
        k_e2_reads<false><<<gg, E2_T, smem_em, st>>>(a, e, x);
        if (one) k_e2_pair<true><<<np, TP, 0, st>>>(a, e, x, (int)r + 1); else k_e2_pair<false><<<np, TP, 0, st>>>(a, e, x, (int)r + 1);
        *launches += 2;
    }
    k_e2_final_reads<<<gg, E2_T, smem_fin, st>>>(a, e, x);
    if (one) k_e2_final_pair<true><<<np, TP, 0, st>>>(a, e, x); else k_e2_final_pair<false><<<np, TP, 0, st>>>(a, e, x);
    *launches += 2;
    PD_CUDA(c, cudaGetLastError());
    return 0;
}

// overflow of the interleaved read-pair copies during the last scan? (exact need in *need_words)
int pd_em2_overflow(pd_ctx * c, bool * ovf, size_t * need_words)
{
    *ovf = false; *need_words = 0;
    if (!c->d_scratch[PD_S_E2_CNT]) return 0;
    uint32_t h[2] = {0, 0};
    PD_CUDA(c, cudaMemcpy(h, c->d_scratch[PD_S_E2_CNT], 8, cudaMemcpyDeviceToHost));
    *ovf = h[1] != 0; *need_words = h[0];
    return 0;
}
