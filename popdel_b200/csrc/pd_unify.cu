// pd_unify.cu -- segment-level merge of the window calls ON THE DEVICE.
//
// Reference: unifyCalls utils_popdel.h:567-654 with similar / delSizeSimilar / enoughOverlap / checkAndExtend
// (:237-320), lowerCall (:327-343), mergeWindowRange (:512-559), setGenotypes (:441-503), setFreqFromGTs (:344-366),
// called once per processSegment() (workflow_popdel.h:48). With pd_set_unify the scan keeps the window calls and
// their N x 13 per-sample rows in device memory and only the MERGED variants cross PCIe.
//
//   k_seg_bounds     window calls are emitted in window order, so the calls of one processSegment() are a run of equal
//                    `segment` tags: first / last index per segment.
//   k_unify_plan     one block per segment. All threads: stable rank sort by (position, deletion length, LR descending)
//                    and a working copy of the call headers in that order. Thread 0: the reference's merge loop -- a
//                    sequential state machine over a few dozen headers whose quirks must be kept (the running lists of
//                    starts / sizes and the window counters are only reset by a SUCCESSFUL mergeWindowRange, the first
//                    call of a later group is not in the lists, `last` itself is outside the merged range). It
//                    fixes position / length / LR / filter of every merged variant and lists its in-range windows.
//   k_unify_offsets  exclusive scan of the variants per segment -> output slots (segment order).
//   k_unify_emit     one block per segment, one thread per sample: PL sums -> setGenotypes, median LAD / DAD by value
//                    bisection, allele count -> frequency; header + row go to mapped host memory.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pd_device.cuh"

namespace {

struct UnifyArgs {
    const pd_call * calls; const uint32_t * ps;        // window calls of the scan (device), window order
    uint32_t n_raw, row_words, N, nseg;
    double sd, min_cover; int output_failed;
    uint32_t * seg_first, * seg_last;                  // [nseg]
    uint32_t * seg_keep;                               // [nseg + 1] variants per segment, then their exclusive prefix
    // per window call; a segment owns the entries [seg_first, seg_last) of each array
    uint32_t * order;                                  // sorted slot -> index of the window call
    pd_call * wc;                                      // working copy of the headers in sorted order
    uint32_t * starts, * sizes;                        // the merge loop's running lists (append-only, "clear" = new offset)
    uint32_t * inr;                                    // window calls in range of the variant starting at a sorted slot
    uint32_t * gwc, * sig;                             // per sorted slot: in-range windows / significant windows (0: no merge)
    uint32_t * total;                                  // mapped host: number of variants
    pd_call * out_calls; uint32_t * out_ps; uint32_t * out_sig;     // mapped host
};

__device__ __forceinline__ bool all_pass(const pd_call & c) { return (c.filter & 31u) == 0; }        // :207-214

__device__ __forceinline__ bool size_similar(uint32_t a, uint32_t b, double sd)                      // :237-257
{
    const uint32_t l = min(a, b), r = max(a, b);
    return (l + 2 * sd >= r) || (l >= 0.5 * r);
}
// :306-320; may extend a.end_position (checkAndExtend :269-280)
__device__ __forceinline__ bool similar_calls(pd_call & a, const pd_call & b, double sd)
{
    if (!size_similar(a.deletion_length, b.deletion_length, sd)) return false;
    const uint32_t aSpan = a.end_position - a.position, bSpan = b.end_position - b.position;
    const uint32_t minLen = min(aSpan, bSpan);
    const uint32_t left = max(a.position, b.position), right = min(a.position + aSpan, b.position + bSpan);
    const int overlap = (int)(right - left);
    if (overlap >= 0.25 * minLen || overlap + 2 * sd >= minLen) return true;
    if ((aSpan < a.deletion_length || bSpan < b.deletion_length) &&
        (b.position - a.position < (min(a.deletion_length, b.deletion_length) + 4 * sd))) { a.end_position = b.end_position; return true; }
    return false;
}

__global__ void k_seg_bounds(UnifyArgs u)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u.n_raw) return;
    const uint32_t s = u.calls[i].segment;
    if (s >= u.nseg) return;
    if (i == 0 || u.calls[i - 1].segment != s) u.seg_first[s] = i;
    if (i == u.n_raw - 1 || u.calls[i + 1].segment != s) u.seg_last[s] = i + 1;
}

__device__ void sort_small(uint32_t * v, uint32_t n)                 // insertion sort (lists of a few dozen entries)
{
    for (uint32_t i = 1; i < n; ++i) {
        const uint32_t x = v[i];
        uint32_t j = i;
        for (; j > 0 && v[j - 1] > x; --j) v[j] = v[j - 1];
        v[j] = x;
    }
}

__global__ void __launch_bounds__(256) k_unify_plan(UnifyArgs u)
{
    const uint32_t seg = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const uint32_t first = u.seg_first[seg], n = u.seg_last[seg] - first;
    if (n <= 1) { if (tid == 0) u.seg_keep[seg] = 0; return; }        // `calls.size() <= 1 -> false`: nothing is written
    // ---- std::sort by lowerCall; ties keep window order (the oracle's choice, equal keys are interchangeable headers)
    for (uint32_t i = tid; i < n; i += T) {
        const pd_call a = u.calls[first + i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const pd_call * b = u.calls + first + j;
            const uint32_t bp = __ldg(&b->position), bl = __ldg(&b->deletion_length);
            bool lower;
            if (bp != a.position) lower = bp < a.position;
            else if (bl != a.deletion_length) lower = bl < a.deletion_length;
            else { const double blr = __ldg(&b->lr); lower = blr > a.lr || (blr == a.lr && j < i); }
            rank += lower;
        }
        u.order[first + rank] = first + i;
        u.wc[first + rank] = a;
        u.gwc[first + rank] = 0; u.sig[first + rank] = 0;
    }
    __syncthreads();
    if (tid != 0) return;
    pd_call * W = u.wc + first;
    uint32_t * S = u.starts + first, * Z = u.sizes + first, * I = u.inr + first, * G = u.gwc + first, * Q = u.sig + first;
    const uint32_t * ord = u.order + first;
    const uint32_t last = n - 1;
    uint32_t cur = 0;
    auto drop_all = [&]() { u.seg_keep[seg] = 0; };
    if (!u.output_failed) {
        while (!all_pass(W[cur])) { if (cur == last) { drop_all(); return; } ++cur; }
        if (cur == last) { drop_all(); return; }
    }
    const uint32_t first_idx = cur;
    for (uint32_t k = 0; k < first_idx; ++k) W[k].filter = 255;
    uint32_t loff = 0, ln = 0;                                         // running lists = S/Z[loff, loff + ln)
    S[0] = W[cur].position; Z[0] = W[cur].deletion_length; ln = 1;
    double lr = W[cur].lr;                                             // (long double in the reference)
    uint32_t winCount = 1, sigWin = 1;
    auto merge_range = [&](uint32_t start, uint32_t lastx) {           // mergeWindowRange :512-559 without the per-sample part
        pd_call & st = W[start];
        sort_small(S + loff, ln); sort_small(Z + loff, ln);
        st.position = S[loff + ln / 2]; st.deletion_length = Z[loff + ln / 2];
        loff += ln; ln = 0;
        st.lr = lr / winCount;
        uint32_t g = 0;
        for (uint32_t k = start; k < lastx; ++k) {
            const uint32_t wp = W[k].window_position;
            if (wp > st.position && wp - 30 < st.position + st.deletion_length) I[start + g++] = ord[k];
        }
        if (g == 0) { st.filter = 255; return; }                       // (returns before the counters are reset)
        G[start] = g; Q[start] = sigWin;
        if (30.0 * sigWin / st.deletion_length < u.min_cover) st.filter |= 16;
        winCount = 1; sigWin = 1; lr = 0.0;
    };
    uint32_t it = first_idx + 1;
    while (true) {
        if (similar_calls(W[cur], W[it], u.sd)) {
            if (all_pass(W[it])) { S[loff + ln] = W[it].position; Z[loff + ln] = W[it].deletion_length; ++ln; ++sigWin; }
            ++winCount;
            lr += W[it].lr;
            W[it].filter = 255;
            if (it == last) { if (ln) merge_range(cur, it); break; }
        } else {
            if (winCount != 1 && ln) merge_range(cur, it);
            else W[cur].filter = 255;
            cur = it;
        }
        if (it != last) ++it;
        else { if (winCount == 1) W[cur].filter = 255; break; }
    }
    uint32_t keep = 0;
    for (uint32_t k = first_idx; k <= last; ++k) keep += W[k].filter != 255;
    u.seg_keep[seg] = keep;
}

__global__ void __launch_bounds__(1024) k_unify_offsets(UnifyArgs u)
{
    __shared__ unsigned long long ws[33];
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < u.nseg; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < u.nseg ? u.seg_keep[i] : 0ull;
        unsigned long long tot;
        const unsigned long long ex = block_excl_scan(v, ws, tot);
        if (i < u.nseg) u.seg_keep[i] = (uint32_t)(carry + ex);
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { u.seg_keep[u.nseg] = (uint32_t)carry; *u.total = (uint32_t)carry; }
}

__global__ void __launch_bounds__(256) k_unify_emit(UnifyArgs u)
{
    __shared__ unsigned long long ws[33];
    const uint32_t seg = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const uint32_t first = u.seg_first[seg], n = u.seg_last[seg] - first;
    if (n <= 1 || u.seg_keep[seg + 1] == u.seg_keep[seg]) return;
    uint32_t slot = u.seg_keep[seg];
    for (uint32_t k = 0; k < n; ++k) {
        const pd_call hdr = u.wc[first + k];
        if (hdr.filter == 255) continue;                               // block-uniform
        const uint32_t g = u.gwc[first + k];
        const uint32_t * own = u.ps + (size_t)u.order[first + k] * u.row_words;
        uint32_t * orow = u.out_ps + (size_t)slot * u.row_words;
        if (g == 0) {                                                  // kept without a merge (stale counters at the end of the loop)
            for (uint32_t i = tid; i < u.row_words; i += T) orow[i] = own[i];
            if (tid == 0) { u.out_calls[slot] = hdr; u.out_sig[slot] = 0; }
            ++slot;
            continue;
        }
        const uint32_t * inr = u.inr + first + k;
        unsigned long long alleles = 0;
        for (uint32_t s = tid; s < u.N; s += T) {
            // ---- setGenotypes :441-503
            uint32_t p0 = 0, p1 = 0, p2 = 0;
            for (uint32_t i = 0; i < g; ++i) {
                const uint32_t * r = u.ps + (size_t)inr[i] * u.row_words + 13 * s;
                p0 += r[0]; p1 += r[1]; p2 += r[2];
            }
            const double mn = (double)min(min(p0, p1), p2);
            const double ref = (p0 - mn) / g, het = (p1 - mn) / g, hom = (p2 - mn) / g;
            uint32_t o[13];
            o[0] = (uint32_t)round(ref); o[1] = (uint32_t)round(het); o[2] = (uint32_t)round(hom);
            // ---- median LAD / DAD over the in-range windows: element g/2 of the sorted values, by bisection on the value
            for (int j = 0; j < 8; ++j) {
                uint32_t lo = 0xFFFFFFFFu, hi = 0;
                for (uint32_t i = 0; i < g; ++i) {
                    const uint32_t v = u.ps[(size_t)inr[i] * u.row_words + 13 * s + 3 + j];
                    lo = min(lo, v); hi = max(hi, v);
                }
                const uint32_t need = g / 2 + 1;
                while (lo < hi) {
                    const uint32_t mid = lo + (hi - lo) / 2;
                    uint32_t cnt = 0;
                    for (uint32_t i = 0; i < g; ++i) cnt += u.ps[(size_t)inr[i] * u.row_words + 13 * s + 3 + j] <= mid;
                    if (cnt >= need) hi = mid; else lo = mid + 1;
                }
                o[3 + j] = lo;
            }
            o[11] = own[13 * s + 11]; o[12] = own[13 * s + 12];
            if (o[1] == o[2]) { if (het > hom) ++o[1]; else ++o[2]; }
            else if (o[0] == o[1]) { if (ref > het) ++o[0]; else ++o[1]; }
            if (o[0] != 0) alleles += (o[1] == 0) ? 1 : 2;             // setFreqFromGTs :344-366
            for (int j = 0; j < 13; ++j) orow[13 * s + j] = o[j];
        }
        unsigned long long tot;
        block_excl_scan(alleles, ws, tot);
        if (tid == 0) {
            pd_call h = hdr;
            h.frequency = (double)tot / (u.N * 2.0);
            u.out_calls[slot] = h; u.out_sig[slot] = u.sig[first + k];
        }
        ++slot;
    }
}

template <typename T>
int grow(pd_ctx * c, int slot, T *& p, size_t need)
{
    const size_t bytes = std::max<size_t>(need, 1) * sizeof(T);
    if (bytes > c->cap_unify[slot] || !c->d_unify[slot]) {
        if (c->d_unify[slot]) cudaFree(c->d_unify[slot]);
        c->d_unify[slot] = nullptr; c->cap_unify[slot] = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
        PD_CUDA(c, cudaMalloc(&c->d_unify[slot], want));
        c->cap_unify[slot] = want;
    }
    p = reinterpret_cast<T *>(c->d_unify[slot]);
    return 0;
}

}  // namespace

// device-side buffers of the window calls (unify mode): grown so that `need_calls` fit; the first `keep_calls` survive
int pd_unify_ensure_raw(pd_ctx * c, size_t need_calls, size_t row, size_t keep_calls)
{
    if (need_calls > c->cap_u_calls || !c->d_u_calls) {
        const size_t floor_calls = getenv("PD_UNIFY_CAP") ? (size_t)std::max(1, atoi(getenv("PD_UNIFY_CAP"))) : 1024;      // test knob
        const size_t want = std::max<size_t>(need_calls + need_calls / 2, floor_calls);
        pd_call * p = nullptr;
        PD_CUDA(c, cudaMalloc(&p, want * sizeof(pd_call)));
        if (keep_calls) PD_CUDA(c, cudaMemcpy(p, c->d_u_calls, keep_calls * sizeof(pd_call), cudaMemcpyDeviceToDevice));
        cudaFree(c->d_u_calls);
        c->d_u_calls = p; c->cap_u_calls = want;
    }
    if (need_calls * row > c->cap_u_ps || !c->d_u_ps) {
        const size_t want = std::max<size_t>((need_calls + need_calls / 2) * row, getenv("PD_UNIFY_CAP") ? 64 : 1u << 18);
        uint32_t * p = nullptr;
        if (cudaMalloc(&p, want * 4) != cudaSuccess) {
            cudaGetLastError();
            return pd_fail(c, PD_ERR_CAPACITY, "device-side unify: the window calls of this scan do not fit device memory; scan a smaller window range");
        }
        if (keep_calls) PD_CUDA(c, cudaMemcpy(p, c->d_u_ps, keep_calls * row * 4, cudaMemcpyDeviceToDevice));
        cudaFree(c->d_u_ps);
        c->d_u_ps = p; c->cap_u_ps = want;
    }
    return 0;
}

void pd_unify_release(pd_ctx * c)
{
    cudaFree(c->d_u_calls); cudaFree(c->d_u_ps);
    for (auto & p : c->d_unify) cudaFree(p);
    if (c->res_sig) cudaFreeHost(c->res_sig);
}

// merges the n_raw window calls held in d_u_calls / d_u_ps; variants go to res_calls / res_ps / res_sig (mapped host)
int pd_run_unify(pd_ctx * c, uint32_t n_raw, size_t row, uint32_t nseg, int (*ensure_results)(pd_ctx *, size_t, size_t, size_t),
                 uint64_t * n_out, uint64_t * launches)
{
    *n_out = 0;
    if (n_raw == 0) return 0;
    cudaStream_t st = c->stream;
    UnifyArgs u;
    memset(&u, 0, sizeof(u));
    u.calls = c->d_u_calls; u.ps = c->d_u_ps; u.n_raw = n_raw; u.row_words = (uint32_t)row; u.N = c->N; u.nseg = nseg;
    u.sd = c->unify.mean_stddev; u.min_cover = c->unify.min_relative_window_cover; u.output_failed = c->unify.output_failed;
    if (grow(c, 0, u.seg_first, (size_t)nseg)) return c->status;
    if (grow(c, 1, u.seg_last, (size_t)nseg)) return c->status;
    if (grow(c, 2, u.seg_keep, (size_t)nseg + 1)) return c->status;
    if (grow(c, 3, u.order, (size_t)n_raw)) return c->status;
    if (grow(c, 4, u.wc, (size_t)n_raw)) return c->status;
    if (grow(c, 5, u.starts, (size_t)n_raw)) return c->status;
    if (grow(c, 6, u.sizes, (size_t)n_raw)) return c->status;
    if (grow(c, 7, u.inr, (size_t)n_raw)) return c->status;
    if (grow(c, 8, u.gwc, (size_t)n_raw)) return c->status;
    if (grow(c, 9, u.sig, (size_t)n_raw)) return c->status;
    u.total = c->res_count + 1;
    PD_CUDA(c, cudaMemsetAsync(u.seg_first, 0, (size_t)nseg * 4, st));
    PD_CUDA(c, cudaMemsetAsync(u.seg_last, 0, (size_t)nseg * 4, st));
    k_seg_bounds<<<(n_raw + 255) / 256, 256, 0, st>>>(u);
    k_unify_plan<<<nseg, 256, 0, st>>>(u);
    k_unify_offsets<<<1, 1024, 0, st>>>(u);
    *launches += 3;
    PD_CUDA(c, cudaGetLastError());
    PD_CUDA(c, cudaStreamSynchronize(st));
    const uint32_t total = c->res_count[1];
    if (total) {
        if (ensure_results(c, total, row, 0)) return c->status;
        if (total > c->cap_res_sig || !c->res_sig) {
            if (c->res_sig) cudaFreeHost(c->res_sig);
            c->res_sig = nullptr; c->cap_res_sig = 0;
            const size_t want = std::max<size_t>((size_t)total + total / 2, 1024);
            PD_CUDA(c, cudaHostAlloc(&c->res_sig, want * 4, cudaHostAllocMapped));
            c->cap_res_sig = want;
        }
        u.out_calls = c->res_calls; u.out_ps = c->res_ps; u.out_sig = c->res_sig;
        k_unify_emit<<<nseg, 256, 0, st>>>(u);
        *launches += 1;
        PD_CUDA(c, cudaGetLastError());
        PD_CUDA(c, cudaStreamSynchronize(st));
    }
    *n_out = total;
    return 0;
}
