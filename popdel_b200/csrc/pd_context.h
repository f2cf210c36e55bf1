// pd_context.h -- host-side context of the scan library (internal).
#ifndef PD_CONTEXT_H_
#define PD_CONTEXT_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "../../include/popdel_b200.h"
#include "pd_common.h"

struct PdLong { uint32_t s, e, pos_rel; int32_t dev; };      // wide entry of a long read pair (16 B)
// tile table entry: first stream word of the tile and the range of wide-list entries that can be active in it
// reach (filled by the packers): far << 24 | first stream word of the tile (relative to off) whose read pair is active in a
// later tile; far = how many tiles ahead the tile's stream read pairs reach at most. for_tile_batches skips the rest.
struct PdTile { uint32_t off, long_lo, long_hi, reach; };
// interleaved likelihood tables, one entry per histogram index (and one floor entry per read group)
struct PdTab {
    // first 32 bytes = everything the EM loop touches (one sector, two 128-bit loads)
    double val;     // processed histogram value I()
    double ln;      // ln(val)
    double lnp;     // ln(val + min_prob) - ln2_d          (g1 when the other hypothesis sits on the floor)
    double fr;      // min_prob / (min_prob + val)         (weight of the floor hypothesis)
    // final pass only
    double l10;     // log10(val)
    double l10p;    // log10(val + min_prob) - log10(2)_d
    double val2;    // = val: the fused kernel's final pass reads this half row only
    double fr2;     // = fr
};

// Device view handed to the kernels by value.
struct PdDev {
    const uint32_t * words;          // packed read-pair stream, all read groups
    const PdTile *   tiles;          // [R][NT+1] tile table (word offsets are multiples of 4)
    const PdLong *   longs;          // wide list of long read pairs, all read groups, sorted by s per read group
    const PdRgConst * rgc;           // [R]
    const uint32_t * sample_rg;      // [N+1] read groups of sample s = sample_rg[s] .. sample_rg[s+1]
    const PdTab * tab;               // likelihood tables of all read groups; entry hist_off-1+... see PdRgConst
    const uint4 * tseg;              // [NT + 1] per tile {nb, wlA, wlB, wlC} of tile_seg (pd_common.h), built once per contig
    uint32_t NT;                     // tiles per read group
    uint32_t N, R;
    uint32_t window_buffer;
    int32_t  t_min;                  // min over read groups of min_init_del_len
    int32_t  t_mark;                 // the screen marks / counts read pairs above this (t_min, or t_known with the second stage)
    int32_t  t_known;                // second screen stage: samples whose Q3 can exceed this get an exact Q3 first (0: stage off)
    uint32_t w_begin, w_end;         // windows to scan [w_begin, w_end)
};

#define PD_CAP_RING 4096u
// The active-coverage counter of ChromosomeProfile::add (profile_structure_popdel_call.h:1084-1113), exactly: the
// reference counts stored read pairs minus the end entries a cursor has passed (getEndCount :870-928). The cursor walks
// the per-segment end sets in order of last window, moves on to the next set lazily (the call that exhausts a set returns
// before it looks into the next one), and correctConsecutiveSwitch (:722-737) zeroes the counter of a read group that sat
// out a whole segment. Restated with a histogram of last windows per set instead of a sorted vector: within one call
// every entry below the read pair's window is passed, so only the counts per set matter, not the order inside a set.
// (The look-ahead branch of getEndCount cannot pass an entry: entries of the set after the write set end at or behind
// the border of the segment being loaded.)
struct PdCapState {
    struct Set {
        std::vector<uint32_t> ring;      // ring[lw % PD_CAP_RING] = entries with last window lw, kc <= lw < kc + PD_CAP_RING
        std::vector<uint32_t> far;       // min-heap of last windows beyond the ring
        uint64_t total = 0, passed = 0;  // entries added / passed by the cursor
        uint64_t right = 0;              // right border of the set (bp, anchor-relative)
        uint32_t kc = 0;                 // every entry with last window < kc has been passed
    };
    Set set[3];
    int write_set = 0, pos_set = 0;      // writeSet, endPosSet
    uint32_t load = 0;                   // activeLoad
    int64_t seg = -1;                    // segment the write set belongs to

    void clear_set(Set & s, uint64_t right, uint32_t wb)
    {
        if (s.ring.empty()) s.ring.assign(PD_CAP_RING, 0u);
        else if (s.total != s.passed) std::fill(s.ring.begin(), s.ring.end(), 0u);
        s.far.clear(); s.total = s.passed = 0; s.right = right; s.kc = (uint32_t)((right - wb) / PD_WIN);
    }
    void start(uint32_t wb)              // resetTo :1717-1748
    {
        for (int k = 0; k < 3; ++k) clear_set(set[k], (uint64_t)(k + 1) * wb, wb);
        write_set = pos_set = 0; load = 0; seg = 0;
    }
    void next_segment(uint32_t wb)       // switchWriteSet :713-721 + correctConsecutiveSwitch :722-737
    {
        const int nxt = (write_set + 1) % 3, nn = (write_set + 2) % 3;
        clear_set(set[nn], set[nxt].right + wb, wb);
        write_set = nxt; ++seg;
        if (pos_set != (write_set + 2) % 3) { pos_set = write_set; load = 0; }
        else if (set[pos_set].total == 0) pos_set = write_set;
    }
    uint32_t end_count(uint32_t b)       // getEndCount :870-928 for a read pair of 30-bp bucket b
    {
        if (set[pos_set].passed == set[pos_set].total) {
            if (pos_set == write_set) return 0;
            pos_set = (pos_set + 1) % 3;
        }
        Set & s = set[pos_set];
        if (s.total == 0) return 0;
        uint32_t c = 0;
        if (b > s.kc) {
            if (s.passed + s.far.size() != s.total) {
                const uint32_t lim = b - s.kc < PD_CAP_RING ? b : s.kc + PD_CAP_RING;
                for (uint32_t k = s.kc; k != lim; ++k) { uint32_t & n = s.ring[k & (PD_CAP_RING - 1)]; c += n; n = 0; }
            }
            s.kc = b;
            while (!s.far.empty() && s.far.front() < b) { std::pop_heap(s.far.begin(), s.far.end(), std::greater<uint32_t>()); s.far.pop_back(); ++c; }
            s.passed += c;
        }
        if (c && s.passed == s.total && pos_set != write_set) pos_set = (pos_set + 1) % 3;
        return c;
    }
    void insert(uint32_t lw)             // CyclicEndEntryTable::add :789-803
    {
        Set & s = set[(uint64_t)lw * PD_WIN >= set[write_set].right ? (write_set + 1) % 3 : write_set];
        if (lw - s.kc < PD_CAP_RING) ++s.ring[lw & (PD_CAP_RING - 1)];
        else { s.far.push_back(lw); std::push_heap(s.far.begin(), s.far.end(), std::greater<uint32_t>()); }
        ++s.total;
    }
    // true: the read pair (bucket b, last window lw) is stored
    bool admit(uint32_t b, uint32_t lw, uint32_t max_load)
    {
        if (load >= max_load) {
            load -= end_count(b);
            if (load >= max_load) return false;
            insert(lw); ++load;
        } else {
            insert(lw); ++load;
            load -= end_count(b);
        }
        return true;
    }
};

struct PdHostRg {                    // host staging of one read group of the current contig (filled by pd_contig_push)
    uint32_t * words = nullptr;      // packed stream (pinned when the context has a device); tiles padded to 4 words
    size_t n_words = 0, cap_words = 0;
    bool words_pinned = false;
    std::vector<uint32_t> tile_rel;  // tile_rel[t] = first word of tile t (relative to this read group), size = tiles seen + 1
    std::vector<uint32_t> tile_reach;// PdTile::reach of tile t (same indexing)
    std::vector<PdLong> longs;
    uint32_t long_span = 0;
    uint32_t cur_tile = 0;           // tile being appended
    uint64_t n_reads = 0, dropped = 0;
    PdCapState cap;                  // active-coverage cap (ChromosomeProfile::add, profile_structure :1084-1113)
    uint32_t last_pos = 0;
    bool any = false;
    // bookkeeping for the reference's last scanned window: per segment (current, previous)
    int64_t seg = -1;                // segment of the most recent read pair
    int64_t S = -1, E_own = -1, E_spill = -1;         // current segment: max pos, max lw <= wl(seg), max lw > wl(seg)
    int64_t prev_seg = -1, prev_E_spill = -1;         // previous non-empty segment
};

// what the last scanned window of a contig depends on (last_scanned_window): final segment kf of the pushed read
// pairs, max start position S and max last window E of end set kf, max spill-over last window of segment kf's own pairs
struct PdTail { int64_t kf = -1, S = -1, E = -1, E_spill = -1; };
struct PdShard;                      // sample sharding state (pd_shard.cu)

// pd_contig_push_pinned: raw arrays; pd_contig_push_compact: 16-bit position remainders per 65 536-bp block + 24-bit deviations
struct PdRawRg {
    const uint32_t * pos = nullptr; const int32_t * dev = nullptr; uint64_t n = 0;
    const uint16_t * lo = nullptr; const uint8_t * d24 = nullptr; const uint32_t * blk = nullptr; uint32_t nblk = 0;
    const uint32_t * w32 = nullptr;  // pd_contig_push_compact32: dev:24 | pos & 0xFF per read pair, blk = 256-bp blocks
    bool on_device = false;          // pos / dev are device pointers (pd_contig_push_device)
    bool compact() const { return lo != nullptr || w32 != nullptr; }
};

struct pd_ctx {
    pd_params params;
    uint32_t N = 0, R = 0;
    int device = -1;                 // -1: host-only context (packing / validation hooks), scans fail
    std::vector<pd_rg> rgs;          // values pointers point into tables_
    std::vector<std::vector<double>> tables_;
    std::vector<PdRgConst> rgc;
    std::vector<uint32_t> sample_rg;
    int32_t t_min = 0;
    std::string err;
    int status = 0;

    // current contig
    bool contig_open = false, packed = false, uploaded = false;
    PdGrid grid{0, 200000};
    std::vector<PdHostRg> hrg;
    std::vector<PdRawRg> raw;            // caller-owned page-locked arrays (device-side packing)
    std::vector<std::vector<uint32_t>> raw_pos_dec;    // compact input decoded on the host (only for the host-packer fallback)
    std::vector<std::vector<int32_t>> raw_dev_dec;
    bool dev_mode = false, host_mode = false;
    bool pinned_staging = true;          // pd_set_staging: page-locked (default) or pageable staging of pd_contig_push
    // packed host image (offset tables; the words stay in the per-read-group staging vectors)
    std::vector<PdTile> h_tiles;
    std::vector<uint32_t> h_long_off;
    std::vector<uint64_t> h_word_base;   // first word of each read group in the device stream
    uint64_t total_words = 0, total_longs = 0;
    uint32_t NT = 0;
    uint64_t n_windows_total = 0;    // reference's last scanned window + 1
    PdTail tail;                     // inputs of n_windows_total (combined over the ranks of a sample-sharded cohort)
    uint64_t min_windows = 0;        // pd_contig_reserve_windows: the tile tables cover at least this many windows
    uint64_t n_reads = 0;

    // device
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: result compaction + D2H, overlapped with the EM
    cudaEvent_t ev[16] = {};
    cudaEvent_t ev_pack[9] = {};        // device packer: one per copy group (+1)
    uint32_t * d_words = nullptr; size_t cap_words = 0;
    PdTile * d_tiles = nullptr; size_t cap_tiles = 0;
    PdLong * d_longs = nullptr; size_t cap_longs = 0;
    PdRgConst * d_rgc = nullptr;
    uint32_t * d_sample_rg = nullptr;
    PdTab * d_tab = nullptr;
    uint32_t * d_min_init = nullptr; // [R] minInitDelLengths (of the whole cohort when sharded by sample)
    PdShard * shard = nullptr;
    // scan scratch (grown on demand)
    void * d_scratch[80] = {}; size_t cap_scratch[80] = {};
    void * d_pack[16] = {}; size_t cap_pack[16] = {};          // device packer scratch (raw arrays, tile firsts, ...)
    // results
    pd_call * res_calls = nullptr; size_t cap_res_calls = 0;   // page-locked, mapped: written by the device (k_emit_rows)
    uint32_t * res_ps = nullptr; size_t cap_res_ps = 0;        // page-locked, mapped
    uint32_t * res_count = nullptr;                            // page-locked, mapped: [0] = number of calls
    // device-side unifyCalls (pd_unify.cu): window calls stay on the device, merged variants go to res_*
    bool unify_on = false;
    pd_unify_params unify{};
    pd_call * d_u_calls = nullptr; size_t cap_u_calls = 0;
    uint32_t * d_u_ps = nullptr; size_t cap_u_ps = 0;
    void * d_unify[16] = {}; size_t cap_unify[16] = {};
    uint32_t * res_sig = nullptr; size_t cap_res_sig = 0;     // page-locked, mapped: significantWindows per variant
    // word -> tile index of the current upload (k_stream's slow path) and wide-list ranges per read group
    bool index_built = false;
    uint32_t * d_gran_off = nullptr, * d_gran_tile = nullptr, * d_long_off = nullptr; size_t cap_gran = 0;
    uint4 * d_tseg = nullptr; size_t cap_tseg = 0;
    uint32_t max_rg_words = 0;
    size_t pool_cap = 0;                                       // active read-pair pool capacity (persists across scans)
    size_t e2_devt_cap = 0;                                    // lane-interleaved read-pair copies of pd_em2.cu (words)
    int e2_retry = 0;
    bool e2_attr_set = false;
    float ms_h2d = 0;
    uint64_t h2d_bytes = 0;
};

int pd_fail(pd_ctx * c, int status, const std::string & msg);
int pd_pack_contig(pd_ctx * c);                    // finalises the offset tables of the packed image
int pd_pack_on_device(pd_ctx * c);                // pd_pack.cu: 0 ok, 1 = use the host packer, <0 error
int pd_run_scan(pd_ctx * c, uint64_t first_window, uint64_t n_windows, pd_result * out);   // pd_scan.cu
uint64_t pd_tail_windows(const PdTail & t, uint32_t window_buffer);
void pd_shard_release(pd_ctx * c);                // pd_shard.cu
int pd_unify_ensure_raw(pd_ctx * c, size_t need_calls, size_t row, size_t keep_calls);       // pd_unify.cu
void pd_unify_release(pd_ctx * c);
int pd_run_unify(pd_ctx * c, uint32_t n_raw, size_t row, uint32_t nseg, int (*ensure_results)(pd_ctx *, size_t, size_t, size_t),
                 uint64_t * n_out, uint64_t * launches);

#endif
