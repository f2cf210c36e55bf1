// pd_common.h -- packed read-pair layout and the closed-form activity rule shared by host packer and kernels.
//
// Layout in HBM (DESIGN.md "Data layout"): per read group one position-sorted stream of 32-bit words,
//   word = dev:21 (signed, bits 31..11) | long:1 (bit 10) | pit:10 (bits 9..0)
// pit = position inside the read pair's TILE (32 windows = 960 bp, relative to the contig anchor); tile t of read
// group g occupies words [tile_off[g*(NT+1)+t], tile_off[g*(NT+1)+t+1]) and every tile starts at a multiple of 4
// words (padding words = PD_PAD_WORD are ignored by every kernel), so the streams can be read with 128-bit loads.
// Read pairs whose active interval reaches further than PD_LOOKBACK_TILES tiles are flagged `long` in the stream and
// duplicated into a small per-read-group wide list (pos_rel, dev) that every kernel consults separately.
//
// Activity rule (restates the reference's cyclic tables in closed form; SURVEY.md App. D, verified against the
// reference through oracle/):  a read pair (pos, dev) of read group g, bucket b = (pos-anchor)/30, is active at
// grid window w (position anchor+30w) iff  s <= w <= e  with
//   s  = b + (pos%30 != 0)                        nextWindow activates startPos <= currentPos   (profile_structure :1229-1235)
//   lw = (pos-anchor + max(0, dev+median-2*readLen))/30   lastWindow of the end entry            (load_profile :473-477, :789-791)
//   e  = lw + 1                                    removed when lastWin + 30 < currentPos        (profile_structure :1236-1242)
// and the segment artefacts, segment j = floor(30b / windowBuffer), wl(j) = last window before border j+1:
//   s > wl(j)              -> never active  (start set abandoned at the switch, :1887,1907 + :1125)
//   lw <= wl(j)            -> e = min(e, wl(j))     (end set drained at the switch, :1884-1885,1904-1905)
//   lw >  wl(j)            -> e = min(e, wl(j+1))   (the spill-over end set is drained one switch later)
#ifndef PD_COMMON_H_
#define PD_COMMON_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PD_HD __host__ __device__ __forceinline__
#else
#define PD_HD inline
#endif

#define PD_WIN 30u
#define PD_TILE_WINDOWS 32u
#define PD_TILE_BP (PD_WIN * PD_TILE_WINDOWS)       /* 960 */
#define PD_DEV_BITS 21
#define PD_DEV_MAX ((1 << (PD_DEV_BITS - 1)) - 1)
#define PD_DEV_MIN (-(1 << (PD_DEV_BITS - 1)))
#define PD_LONG_BIT 0x400u
#define PD_PAD_WORD 0x80000400u                     /* dev = PD_DEV_MIN, long flag set, pit = 0 */
#define PD_MAX_LOOKBACK_TILES 8

struct PdRgConst {                  // per read group, device-resident
    int32_t  inner_off;             // median - 2*readLength
    uint32_t max_load;
    uint32_t sample;
    uint32_t lookback_tiles;        // stream read pairs of tiles [t-lookback, t] can be active in tile t
    int32_t  median;
    int32_t  hist_base;             // median - offset: table index = dev + hist_base
    uint32_t hist_len;
    uint32_t hist_off;              // index of this read group's FLOOR entry in the table array; values follow at hist_off+1+i
    double   min_prob, ln_min_prob, l10_min_prob;
    double   stddev;
    int32_t  lower_q, upper_q;
    uint32_t min_init;              // minInitDelLengths[rg]
    uint32_t pad_;
};

struct PdGrid {                     // window grid of the current contig
    uint32_t anchor;                // position of window 0 (multiple of 30)
    uint32_t window_buffer;         // segment length in bp
};

PD_HD uint32_t pd_pack(int32_t dev, uint32_t pit, bool is_long)
{
    return ((uint32_t)dev << 11) | (is_long ? PD_LONG_BIT : 0u) | pit;
}
PD_HD int32_t pd_word_dev(uint32_t w) { return (int32_t)w >> 11; }
PD_HD uint32_t pd_word_pit(uint32_t w) { return w & 0x3FFu; }
PD_HD bool pd_word_long(uint32_t w) { return (w & PD_LONG_BIT) != 0; }

// last window index of segment j (largest w with 30*w < (j+1)*window_buffer)
PD_HD uint64_t pd_seg_last_window(uint64_t j, uint32_t window_buffer)
{
    return ((j + 1) * (uint64_t)window_buffer - 1) / PD_WIN;
}

// Active interval [s, e] (window indices) of a read pair at anchor-relative position pos_rel.
// Returns false when the read pair is never active.
PD_HD bool pd_interval(uint64_t pos_rel, int32_t dev, int32_t inner_off, uint32_t window_buffer,
                       int64_t & s, int64_t & e)
{
    uint64_t b = pos_rel / PD_WIN;
    int64_t inner = (int64_t)dev + inner_off;
    if (inner < 0) inner = 0;
    uint64_t lw = (pos_rel + (uint64_t)inner) / PD_WIN;
    uint64_t j = (b * PD_WIN) / window_buffer;
    uint64_t wl = pd_seg_last_window(j, window_buffer);
    s = (int64_t)b + ((pos_rel - b * PD_WIN) != 0 ? 1 : 0);
    e = (int64_t)lw + 1;
    if ((uint64_t)s > wl) return false;
    if (lw <= wl) { if ((uint64_t)e > wl) e = (int64_t)wl; }
    else { uint64_t wl2 = pd_seg_last_window(j + 1, window_buffer); if ((uint64_t)e > wl2) e = (int64_t)wl2; }
    return true;
}

// Same rule with the per-tile segment constants hoisted (what the kernels evaluate per stream word).
struct TileSeg { uint32_t base_bp; uint32_t nb; int32_t wlA, wlB, wlC; };

PD_HD TileSeg tile_seg(uint32_t tile, uint32_t wb)
{
    // positions are 32-bit, so tile * 960 fits 32 bits and the only division by a run-time value is a 32-bit one
    // (this runs once per (warp, read group, tile) in the gather / count kernels)
    TileSeg t;
    uint32_t base = tile * PD_TILE_BP;
    uint32_t j0 = base / wb;
    uint64_t nb = ((uint64_t)j0 + 1) * wb;
    t.base_bp = base;
    t.nb = (uint32_t)nb;
    t.wlA = (int32_t)((nb - 1) / PD_WIN);
    t.wlB = (int32_t)((nb + wb - 1) / PD_WIN);
    t.wlC = (int32_t)((nb + 2ull * wb - 1) / PD_WIN);
    return t;
}

// returns false for pads / long read pairs / never-active read pairs
PD_HD bool word_interval(uint32_t w, const TileSeg & ts, int32_t inner_off, int32_t & s, int32_t & e,
                                              int32_t & dev, uint32_t & pos_rel)
{
    if (w & PD_LONG_BIT) return false;
    dev = (int32_t)w >> 11;
    pos_rel = ts.base_bp + (w & 0x3FFu);
    uint32_t b = pos_rel / PD_WIN;
    uint32_t bp = b * PD_WIN;
    int32_t inner = dev + inner_off;
    inner = inner < 0 ? 0 : inner;
    int32_t lw = (int32_t)((pos_rel + (uint32_t)inner) / PD_WIN);
    bool next = bp >= ts.nb;
    int32_t wl = next ? ts.wlB : ts.wlA;
    int32_t wl2 = next ? ts.wlC : ts.wlB;
    s = (int32_t)b + (pos_rel != bp ? 1 : 0);
    e = lw + 1;
    int32_t cap = lw <= wl ? wl : wl2;
    e = e < cap ? e : cap;
    return s <= wl;
}

// Smallest number of values above T among n sorted values such that the upper-half median (Q3) can exceed T
// (genotype_deletion_popdel_call.h:15-27): n<4 -> the maximum; else position (3n+2+n%2)/4-1 = l + r,
// Q3 = (1-r)v[l] + r v[l+1] <= v[l+1] (r>0) or = v[l] (r==0).
PD_HD uint32_t pd_q3_need(uint32_t n)
{
    if (n < 4) return 1;
    uint32_t q = 3 * n + 2 + (n & 1);
    uint32_t l = q / 4 - 1;
    return (q % 4 == 0) ? n - l : n - l - 1;
}

#endif
