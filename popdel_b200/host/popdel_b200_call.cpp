// popdel_b200_call -- host shell of the B200 scan: a drop-in for `popdel call` (reference workflow_popdel.h:256-371).
//
// Host side (this file, plain C++17, all cores): command line, profile decoding into a flat structure-of-arrays image
// per file, histogram preprocessing and parameter calculation, regions of interest, the segment loader (which read pairs
// the reference loads in which segment), the VCF writer. Device side, through the C ABI of libpopdel_b200.so
// (include/popdel_b200.h): every window of the scan and the segment-level merge (unifyCalls, pd_set_unify). There is
// no CPU path for the scan: without a CUDA device the program stops with an error.
//
// Usage: popdel_b200_call [options] PROFILE-LIST-FILE | PROFILE1 PROFILE2 [...]
//   -o FILE  output VCF (popdel.vcf)        -n  window-wise output, no merging       -F  also write failed calls
//   -l NUM   min initial deletion length    -m NUM  min deletion length              -a NUM  active-coverage cap (100)
//   -b NUM   segment length in bp (200000)  -t NUM  EM iterations (15)               -p NUM  prior (1e-4)
//   -s NUM   min sample fraction (0.1)      -c NUM  min relative window cover (0.5)  -f NUM  pseudo-count fraction (500)
//   -u       unsmoothed histograms          -x  uncompressed profiles                -C  cell-population priors
//   -r CHR[:BEGIN[-END]]  region of interest (1-based, closed; may be repeated)        -R FILE  one region per line
//   -A FILE  per-read-group active-coverage caps ("ReadGroup maxCov" lines)        -e  rename conflicting read-group IDs per file
//   -g NUM   CUDA device (0); -1 = dry run of the host path (no scan, header-only VCF; for profiling the loader)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>
#include <zlib.h>

#include "../../include/popdel_b200.h"

namespace {

[[noreturn]] void die(const std::string & msg) { std::cerr << "[popdel_b200] " << msg << std::endl; exit(1); }
inline int rnd(double d) { return (int)std::floor(d + 0.5); }

// fn(i) for i in [0, n) on all host cores (dynamic schedule)
template <typename F> void parallelFor(size_t n, F fn)
{
    static const unsigned cores = getenv("PD_THREADS") ? (unsigned)std::max(1, atoi(getenv("PD_THREADS"))) : std::thread::hardware_concurrency();
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(n, cores));
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (auto & th : pool) th.join();
}

struct StageTimer {                                          // PD_TIMING=1: wall time per stage on stderr
    bool on = getenv("PD_TIMING") != nullptr;
    std::map<std::string, double> acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char * stage)
    {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        acc[stage] += std::chrono::duration<double>(t1 - t0).count();
        t0 = t1;
    }
    void report() const { if (on) for (const auto & kv : acc) fprintf(stderr, "[popdel_b200] %-12s %9.3f s\n", kv.first.c_str(), kv.second); }
};

inline void appendUint(std::string & s, uint64_t v)
{
    char buf[24]; int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) s.push_back(buf[--n]);
}
inline void appendG(std::string & s, double v)                // default ostream formatting of a double (= %g)
{
    char buf[40];
    s.append(buf, (size_t)snprintf(buf, sizeof(buf), "%g", v));
}

// ---------------------------------------------------------------------------------------------------------------
// profile files (format: reference insert_histogram_popdel.h:334-525 header, window_podel.h:211-251 records)
// ---------------------------------------------------------------------------------------------------------------
struct RgHeader { std::string name; uint32_t median, readLength; double stddev; int32_t offset; std::vector<double> counts; };
struct Profile {
    std::string path;
    uint32_t indexRegionSize = 10000;
    std::vector<RgHeader> rgs;
    std::vector<std::string> contigNames;
    std::vector<int32_t> contigLengths;
    // 256-bp window records in file order, flat: the read pairs of (window w, read group g) are
    // recPos / recDev[winOff[w * nrg + g] .. winOff[w * nrg + g + 1])
    size_t nrg = 0;
    std::vector<int32_t> winChrom;
    std::vector<uint32_t> winBegin;
    std::vector<uint64_t> winOff;
    std::vector<uint32_t> recPos;
    std::vector<int32_t> recDev;
    std::vector<size_t> contigFirst;                         // first window of each contig (number of windows if none)
    size_t numWins() const { return winBegin.size(); }
    // file layout: byte offset of the body, the index (one entry per 10-kbp region, contig after contig), first entry per contig
    uint64_t bodyStart = 0, fileSize = 0;
    std::vector<uint64_t> index;
    std::vector<size_t> regionBase;
};

template <typename T> T get(const std::vector<unsigned char> & d, size_t & o)
{
    if (o + sizeof(T) > d.size()) die("truncated profile");
    T v; memcpy(&v, d.data() + o, sizeof(T)); o += sizeof(T); return v;
}

// inflates the gzip members in in[off, end) and appends to `out`
void inflateMembers(const std::vector<unsigned char> & in, size_t off, size_t end, std::vector<unsigned char> & out)
{
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 31) != Z_OK) die("zlib init failed");
    // zlib counts its input in 32 bits: feed at most 1 GiB at a time
    size_t fed = off;
    zs.next_in = const_cast<unsigned char *>(in.data() + off);
    zs.avail_in = 0;
    // inflate straight into `out` (grown geometrically, ~3.5x the compressed size is typical): no bounce buffer
    size_t have = out.size();
    out.resize(have + std::max<size_t>((end - off) * 4, 1u << 16));
    int rc = Z_STREAM_END;                                          // (an empty range is complete)
    while (zs.avail_in > 0 || fed < end) {
        if (zs.avail_in == 0) { const size_t n = std::min<size_t>(end - fed, 1u << 30); zs.avail_in = (uInt)n; fed += n; }
        if (have == out.size()) out.resize(out.size() + out.size() / 2);
        zs.next_out = out.data() + have; zs.avail_out = (uInt)std::min<size_t>(out.size() - have, 1u << 30);
        const size_t room = zs.avail_out;
        rc = inflate(&zs, Z_NO_FLUSH);
        have += room - zs.avail_out;
        if (rc == Z_STREAM_END) { if (zs.avail_in == 0 && fed == end) break; inflateReset(&zs); }
        else if (rc != Z_OK && !(rc == Z_BUF_ERROR && zs.avail_out == 0)) die("corrupt gzip block in profile");
    }
    // the last member must have ended: a profile cut off inside a gzip member would otherwise lose read pairs silently
    if (rc != Z_STREAM_END) die("truncated gzip member in profile");
    out.resize(have);
    inflateEnd(&zs);
}

// parses header + index + contig table from the first bytes of a file; false if `d` ends inside the header
bool parseHeader(const std::string & path, const std::vector<unsigned char> & d, Profile & p)
{
    size_t o = 7;
    auto have = [&](size_t n) { return o + n <= d.size(); };
    auto u32 = [&]() { uint32_t v; memcpy(&v, d.data() + o, 4); o += 4; return v; };
    p.path = path; p.rgs.clear(); p.contigNames.clear(); p.contigLengths.clear(); p.index.clear();
    if (!have(8)) return false;
    p.indexRegionSize = u32();
    const uint32_t nRegions = u32();
    if (p.indexRegionSize == 0) die("'" + path + "': index region size 0");
    if (!have(8ull * nRegions + 4)) return false;
    p.index.resize(nRegions);
    if (nRegions) memcpy(p.index.data(), d.data() + o, 8ull * nRegions);
    o += 8ull * nRegions;
    const uint32_t nrg = u32();
    if (nrg == 0) die("profile without read groups");
    for (uint32_t i = 0; i < nrg; ++i) {
        RgHeader h;
        if (!have(4)) return false;
        const uint32_t nl = u32();
        if (nl == 0 || !have((size_t)nl + 24)) return false;
        h.name.assign((const char *)d.data() + o, nl - 1); o += nl;
        h.median = u32(); memcpy(&h.stddev, d.data() + o, 8); o += 8; h.readLength = u32();
        h.offset = (int32_t)u32();
        const uint32_t end = u32();
        if (end < (uint32_t)h.offset || !have(8ull * (end - h.offset))) return false;
        h.counts.resize(end - h.offset);
        if (!h.counts.empty()) memcpy(h.counts.data(), d.data() + o, 8ull * h.counts.size());
        o += 8ull * h.counts.size();
        p.rgs.push_back(h);
    }
    if (!have(4)) { p.rgs.clear(); return false; }
    const uint32_t nc = u32();
    for (uint32_t i = 0; i < nc; ++i) {
        if (!have(4)) { p.rgs.clear(); return false; }
        const uint32_t nl = u32();
        if (nl == 0 || !have((size_t)nl + 4)) { p.rgs.clear(); return false; }
        p.contigNames.emplace_back((const char *)d.data() + o, nl - 1); o += nl;
        int32_t len; memcpy(&len, d.data() + o, 4); o += 4;
        p.contigLengths.push_back(len);
    }
    p.nrg = nrg;
    p.bodyStart = o;
    p.regionBase.assign(nc + 1, 0);
    for (uint32_t i = 0; i < nc; ++i) p.regionBase[i + 1] = p.regionBase[i] + (size_t)(std::max(p.contigLengths[i], 0) / p.indexRegionSize) + 1;
    if (p.regionBase[nc] > p.index.size()) p.regionBase.assign(nc + 1, 0), p.index.clear();      // index does not cover the contigs: no seeking
    return true;
}

// reads `n` bytes at `off`
void readAt(const std::string & path, uint64_t off, size_t n, std::vector<unsigned char> & out)
{
    out.resize(n);
    if (n == 0) return;
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) die("cannot open profile '" + path + "'");
    f.seekg((std::streamoff)off);
    if (!f.read(reinterpret_cast<char *>(out.data()), (std::streamsize)n)) die("cannot read profile '" + path + "'");
}

// header, index and contig table of a profile (insert_histogram_popdel.h:334-525); the body stays on disk
void loadHeader(const std::string & path, Profile & p)
{
    {
        std::ifstream f(path, std::ios::binary);
        if (!f.good()) die("cannot open profile '" + path + "'");
        f.seekg(0, std::ios::end);
        p.fileSize = (uint64_t)std::max<std::streamoff>(f.tellg(), 0);
    }
    std::vector<unsigned char> d;
    readAt(path, 0, (size_t)std::min<uint64_t>(p.fileSize, 15), d);
    if (d.size() < 15 || memcmp(d.data(), "POPDEL\1", 7) != 0) die("'" + path + "' is not a PopDel profile (magic string)");
    uint32_t nRegions; memcpy(&nRegions, d.data() + 11, 4);
    // the header is small next to the body but has no length field: read generously, everything if that was not enough
    uint64_t guess = std::min<uint64_t>(p.fileSize, 15 + 8ull * nRegions + (4u << 20));
    for (;;) {
        readAt(path, 0, (size_t)guess, d);
        if (parseHeader(path, d, p) || guess == p.fileSize) break;
        guess = p.fileSize;
    }
    if (p.rgs.empty()) die("truncated profile '" + path + "'");
}

// one contig region of the body -> the flat image (previous contents are dropped): the windows of contig `c` from the
// index region of `beginPos` to the one of `endPos + 29`; the byte range comes from the index like the reference's
// jumpToRegion (insert_histogram_popdel.h:531-562), so a region query decodes only what it needs
void loadBody(Profile & p, bool uncompressed, int32_t c, uint64_t beginPos, uint64_t endPos, unsigned threads)
{
    const std::string & path = p.path;
    const uint32_t nrg = (uint32_t)p.nrg;
    const size_t nc = p.contigNames.size();
    p.winChrom.clear(); p.winBegin.clear(); p.winOff.clear(); p.recPos.clear(); p.recDev.clear();
    p.contigFirst.assign(nc + 1, 0);
    uint64_t a = p.fileSize, e = p.fileSize;
    if ((size_t)c < nc && p.index.empty()) { a = p.bodyStart; e = p.fileSize; }      // no usable index: the whole body, filtered below
    else if ((size_t)c < nc) {
        const uint64_t irs = p.indexRegionSize;
        const size_t r0 = p.regionBase[c], nr = p.regionBase[c + 1] - r0;
        // (a 256-bp window that begins up to 29 bp beyond the region's end can still hold a 30-bp bucket below it: the
        // bucket grid is anchored at the window's first read pair, window_podel.h:314-418)
        const uint64_t rb = beginPos / irs, re = (endPos + 29) / irs + 1;
        if (rb < nr) a = p.index[r0 + rb];
        const size_t endEntry = r0 + (size_t)std::min<uint64_t>(re, nr);
        if (endEntry < p.index.size()) e = p.index[endEntry];
        a = std::min(std::max(a, p.bodyStart), p.fileSize); e = std::min(std::max(e, a), p.fileSize);
    }
    std::vector<unsigned char> d;
    readAt(path, a, (size_t)(e - a), d);
    const size_t o = 0;
    p.nrg = nrg;
    // ---- body: 256-bp window records, one gzip member per 10-kbp index region (or raw with -x). The index offsets are
    // record boundaries, so runs of members can be inflated and parsed independently: `threads` tasks per file when there
    // are fewer files than cores
    struct Part { std::vector<int32_t> chrom; std::vector<uint32_t> begin, pos; std::vector<uint64_t> off; std::vector<int32_t> dev; };
    auto parse = [&](const unsigned char * body, size_t size, Part & q) {
        q.pos.reserve(size / 5); q.dev.reserve(size / 5);
        size_t b = 0;
        auto u32 = [&]() { uint32_t v; if (b + 4 > size) die("truncated window record in '" + path + "'"); memcpy(&v, body + b, 4); b += 4; return v; };
        while (b + 8 <= size) {
            q.chrom.push_back((int32_t)u32()); q.begin.push_back(u32());
            const uint32_t begin = q.begin.back();
            for (uint32_t g = 0; g < nrg; ++g) {
                const uint32_t n = u32();
                if (b + 5ull * n > size) die("truncated window record in '" + path + "'");
                q.off.push_back(q.pos.size());
                for (uint32_t i = 0; i < n; ++i) {
                    const uint8_t offc = body[b]; b += 1;
                    int32_t dv; memcpy(&dv, body + b, 4); b += 4;
                    q.pos.push_back(begin + offc); q.dev.push_back(dv);
                }
            }
        }
    };
    std::vector<size_t> cuts;                                        // distinct member starts in [a, e), relative to d, ascending
    for (uint64_t v : p.index) if (v >= a && v < e) cuts.push_back((size_t)(v - a));
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    size_t nTasks = std::max<unsigned>(threads, 1);
    if (cuts.empty() || cuts[0] != o || nTasks == 1) { cuts.assign(1, o); nTasks = 1; }      // (index does not start at the body: one task)
    nTasks = std::min(nTasks, cuts.size());
    std::vector<size_t> taskBegin(nTasks + 1, d.size());
    for (size_t t = 0; t < nTasks; ++t) {                            // balanced by compressed bytes, cut at member starts
        const size_t want = o + (d.size() - o) * t / nTasks;
        const auto it = std::lower_bound(cuts.begin(), cuts.end(), want);
        taskBegin[t] = it == cuts.end() ? d.size() : *it;                // (no member starts at or after `want`: an empty task)
    }
    taskBegin[0] = o;
    std::vector<Part> parts(nTasks);
    auto runTask = [&](size_t t) {
        const size_t a = taskBegin[t], e = taskBegin[t + 1];
        if (a >= e) return;
        if (uncompressed) parse(d.data() + a, e - a, parts[t]);
        else { std::vector<unsigned char> body; inflateMembers(d, a, e, body); parse(body.data(), body.size(), parts[t]); }
    };
    if (nTasks == 1) runTask(0);
    else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < nTasks; ++t) pool.emplace_back(runTask, t);
        for (auto & th : pool) th.join();
    }
    if (nTasks == 1) {
        p.winChrom.swap(parts[0].chrom); p.winBegin.swap(parts[0].begin); p.winOff.swap(parts[0].off);
        p.recPos.swap(parts[0].pos); p.recDev.swap(parts[0].dev);
    } else {
        size_t nWin = 0, nRec = 0;
        for (const Part & q : parts) { nWin += q.begin.size(); nRec += q.pos.size(); }
        p.winChrom.reserve(nWin); p.winBegin.reserve(nWin); p.winOff.reserve(nWin * nrg + 1); p.recPos.reserve(nRec); p.recDev.reserve(nRec);
        for (Part & q : parts) {
            const uint64_t base = p.recPos.size();
            p.winChrom.insert(p.winChrom.end(), q.chrom.begin(), q.chrom.end());
            p.winBegin.insert(p.winBegin.end(), q.begin.begin(), q.begin.end());
            for (uint64_t v : q.off) p.winOff.push_back(base + v);
            p.recPos.insert(p.recPos.end(), q.pos.begin(), q.pos.end());
            p.recDev.insert(p.recDev.end(), q.dev.begin(), q.dev.end());
            q = Part();
        }
    }
    // only contig c's windows are visible (a back-filled index entry may have led into the next contig)
    size_t lo = 0;
    while (lo < p.winBegin.size() && p.winChrom[lo] != c) ++lo;
    if (lo > 0) {                                                     // (only without an index: windows of earlier contigs)
        const uint64_t drop = lo < p.winBegin.size() ? p.winOff[lo * nrg] : (uint64_t)p.recPos.size();
        p.winChrom.erase(p.winChrom.begin(), p.winChrom.begin() + lo); p.winBegin.erase(p.winBegin.begin(), p.winBegin.begin() + lo);
        p.winOff.erase(p.winOff.begin(), p.winOff.begin() + lo * nrg);
        for (uint64_t & v : p.winOff) v -= drop;
        p.recPos.erase(p.recPos.begin(), p.recPos.begin() + drop); p.recDev.erase(p.recDev.begin(), p.recDev.begin() + drop);
    }
    size_t keep = 0;
    while (keep < p.winBegin.size() && p.winChrom[keep] == c) ++keep;
    const uint64_t endOff = keep < p.winBegin.size() ? p.winOff[keep * nrg] : (uint64_t)p.recPos.size();
    if (keep < p.winBegin.size()) {
        p.winChrom.resize(keep); p.winBegin.resize(keep); p.winOff.resize(keep * nrg);
        p.recPos.resize(endOff); p.recDev.resize(endOff);
    }
    p.winOff.push_back(endOff);
    for (size_t k = 0; k <= nc; ++k) p.contigFirst[k] = k <= (size_t)c ? 0 : keep;
}

// first window (file order) whose (contig, index region) is >= (c, region): what the index seek lands on
// (reference insert_histogram_popdel.h:531-562; empty index entries are back-filled with the next offset :298-328)
size_t indexSeek(const Profile & p, int32_t c, uint32_t pos)
{
    const uint32_t region = pos / p.indexRegionSize;
    size_t lo = p.contigFirst[c], hi = p.contigFirst[c + 1];
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (p.winBegin[mid] / p.indexRegionSize < region) lo = mid + 1; else hi = mid; }
    return lo;
}

// 30-bp buckets of a 256-bp window (reference window_podel.h:314-418): the grid is anchored at the window's first read
// pair over all read groups; returns that anchor (0xFFFFFFFF for a window without read pairs)
uint32_t bucketBase(const Profile & p, size_t w)
{
    uint32_t first = 0xFFFFFFFFu;
    for (size_t g = 0; g < p.nrg; ++g) {
        const uint64_t a = p.winOff[w * p.nrg + g], b = p.winOff[w * p.nrg + g + 1];
        if (a < b) first = std::min(first, p.recPos[a]);
    }
    return first == 0xFFFFFFFFu ? first : (first / 30) * 30;
}

// ---------------------------------------------------------------------------------------------------------------
// calls and VCF (vcfout_popdel_call.h); the segment-level merge (unifyCalls, utils_popdel.h:237-654) runs on the
// device behind pd_set_unify
// ---------------------------------------------------------------------------------------------------------------
// LR -> QUAL (reference QuantileMap, parameter_parsing_popdel_call.h:14-136, always built with prior 1e-4):
// keys qchisq(1-10^(-i/10), df=1)/2 - ln(prior/(1-prior)), i = 1..100; QUAL = i of the largest key <= LR.
struct QualMap {
    std::vector<double> keys;
    QualMap()
    {
        const double p = std::log(0.0001 / (1 - 0.0001));
        for (int i = 1; i <= 100; ++i) {
            const long double q = std::pow(10.0L, -i / 10.0L);           // upper tail of chi^2_1 = erfc(sqrt(x/2))
            long double z = std::sqrt(-2.0L * std::log(q)), prev = 0;     // solve erfc(z/sqrt2) = q by Newton
            for (int k = 0; k < 100 && std::fabs((double)(z - prev)) > 1e-17L * (double)z; ++k) {
                prev = z;
                const long double fz = erfcl(z / sqrtl(2.0L)) - q;
                const long double dfz = -sqrtl(2.0L / 3.14159265358979323846264338327950288L) * expl(-z * z / 2);
                z -= fz / dfz;
            }
            keys.push_back((double)(z * z) / 2 - p);
        }
    }
    int qual(double lr) const
    {
        if (lr < 0) return 0;
        int i = (int)(std::upper_bound(keys.begin(), keys.end(), lr) - keys.begin());     // number of keys <= lr
        return i >= 100 ? 100 : i;                                 // beyond the last key the reference prints 100 as well
    }
};

std::string sampleName(const std::string & path)                  // reference utils_popdel.h:1469-1489
{
    size_t lastDelim = 0, lastDot = path.size();
    for (size_t i = 0; i < path.size(); ++i) { if (path[i] == '/' || path[i] == '\\') lastDelim = i; else if (path[i] == '.') lastDot = i; }
    if (lastDelim == 0) return path.substr(0, lastDot);
    if (lastDot < lastDelim) return path.substr(lastDelim + 1);
    return path.substr(lastDelim + 1, lastDot - lastDelim - 1);
}

void writeHeader(std::ostream & out, const Profile & first, const std::vector<std::string> & files)
{
    time_t now = time(nullptr); char buf[32];
    strftime(buf, sizeof(buf), "[%Y%m%d] ", localtime(&now));
    out << "##fileformat=VCFv4.3\n##fileDate=" << buf << "\n##source=PopDel-V1.5.0\n";
    for (size_t i = 0; i < first.contigNames.size(); ++i)
        out << "##contig=<ID=" << first.contigNames[i] << ",length=" << first.contigLengths[i] << ">\n";
    out << "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency\">\n"
           "##INFO=<ID=IMPRECISE,Number=0,Type=Flag,Description=\"Imprecise structural variation\">\n"
           "##INFO=<ID=SVLEN,Number=.,Type=Integer,Description=\"Difference in length between REF and ALT alleles\">\n"
           "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n"
           "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the structural variant\">\n"
           "##INFO=<ID=SVMETHOD,Number=1,Type=String,Description=\"Approach used to detect the structural variant\">\n"
           "##INFO=<ID=LR,Number=1,Type=String,Description=\"Log-Likelihood ratio that the test is correct\">\n"
           "##INFO=<ID=YIELD,Number=1,Type=Float,Description=\"Fraction of genotyped samples\">\n"
           "##INFO=<ID=SWIN,Number=1,Type=Integer,Description=\"Number of significant windows merged into this variant\">\n"
           "##FILTER=<ID=LowLR,Description=\"Likelihood ratio below threshold\">\n"
           "##FILTER=<ID=missingSamples,Description=\"Too many samples not genotyped\">\n"
           "##FILTER=<ID=allRefGT,Description=\"All samples genotyped as homozygous reference\">\n"
           "##FILTER=<ID=CSWin,Description=\"Low fraction of significant windows\">\n"
           "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
           "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Phred-scaled genotype likelihoods rounded to the closest integer\">\n"
           "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype quality. Difference of the best and second-best PL\">\n"
           "##FORMAT=<ID=LAD,Number=3,Type=Integer,Description=\"Likelihood derived allelic depth: Count of read-pairs supporting REF, ambiguous, ALT\">\n"
           "##FORMAT=<ID=DAD,Number=5,Type=Integer,Description=\"Distribution derived allelic depth: Count of read-pairs supporting REF only, REF and ALT, neither(between the histograms), ALT only, bigger than ALT\">\n"
           "##FORMAT=<ID=FL,Number=2,Type=Integer,Description=\"Window of the first and last read active in this window\">\n"
           "##FORMAT=<ID=FLD,Number=1,Type=Integer,Description=\"Distance between first and last window\">\n";
    out << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    for (const std::string & f : files) out << "\t" << sampleName(f);
    out << "\n";
}

// one VCF record (vcfout_popdel_call.h:61-205) appended to `out`; per-sample columns are formatted without streams
// (N samples x ~40 bytes per record)
void formatRecord(std::string & out, const std::string & chrom, const pd_call & c, const uint32_t * ps, size_t N, uint32_t sigWin,
                  const QualMap & qm)
{
    unsigned genotyped = (unsigned)N;
    for (size_t s = 0; s < N; ++s) if (ps[13 * s] + ps[13 * s + 1] + ps[13 * s + 2] == 0) --genotyped;
    std::string filter;
    if ((c.filter & 31u) == 0) filter = "PASS";
    else {
        auto add = [&](const char * t) { if (!filter.empty()) filter += ';'; filter += t; };
        if (c.filter & 1) add("lowLR");
        if (c.filter & 2) add("highCov");
        if (c.filter & 4) add("missingSamples");
        if (c.filter & 8) add("allRefGT");
        if (c.filter & 16) add("CSWin");
    }
    const uint32_t pos = c.position > 1 ? c.position - 1 : c.position;           // record.beginPos (0-based)
    out.reserve(out.size() + 160 + 48 * N);
    out += chrom; out += '\t'; appendUint(out, (uint64_t)pos + 1); out += "\t.\tN\t<DEL>\t"; appendUint(out, (uint64_t)qm.qual(c.lr));
    out += '\t'; out += filter;
    out += "\tIMPRECISE;SVLEN=";                                // default stream precision (6), like the reference
    { const int sv = -static_cast<int>(c.deletion_length); if (sv < 0) { out += '-'; appendUint(out, (uint64_t)(-(int64_t)sv)); } else appendUint(out, (uint64_t)sv); }
    out += ";END="; appendUint(out, (uint32_t)(c.position + c.deletion_length));
    out += ";SVTYPE=DEL;AF="; appendG(out, c.frequency);
    out += ";LR="; appendG(out, c.lr);
    out += ";SVMETHOD=PopDelv1.5.0;YIELD="; appendG(out, (double)genotyped / N);
    out += ";SWIN="; appendUint(out, sigWin);
    out += "\tGT:PL:GQ:LAD:DAD:FL:FLD";
    for (size_t s = 0; s < N; ++s) {
        const uint32_t * o = &ps[13 * s];
        unsigned gq = 0;
        if (o[0] + o[1] + o[2] != 0) gq = o[0] == 0 ? std::min(o[1], o[2]) : (o[1] == 0 ? std::min(o[0], o[2]) : std::min(o[0], o[1]));
        gq = std::min(gq, 255u);
        out += '\t';
        if (gq == 0) out += "./.:0,0,0";
        else {
            out += (o[2] == 0 ? '1' : '0'); out += '/'; out += (o[0] != 0 ? '1' : '0'); out += ':';
            appendUint(out, std::min(o[0], 255u)); out += ','; appendUint(out, std::min(o[1], 255u)); out += ','; appendUint(out, std::min(o[2], 255u));
        }
        out += ':'; appendUint(out, gq);
        out += ':'; appendUint(out, o[3]); out += ','; appendUint(out, o[4]); out += ','; appendUint(out, o[5]);
        out += ':'; appendUint(out, o[6]); out += ','; appendUint(out, o[7]); out += ','; appendUint(out, o[8]); out += ','; appendUint(out, o[9]);
        out += ','; appendUint(out, o[10]);
        out += ':'; appendUint(out, o[11]); out += ','; appendUint(out, o[12]); out += ':'; appendUint(out, (uint32_t)(o[12] - o[11]));
    }
    out += '\n';
}

// region of interest, 0-based half-open like the reference's GenomicRegion (parseGenomicRegion: "chr", "chr:begin",
// "chr:begin-end", 1-based closed, thousands separators allowed)
struct Roi { int32_t contig; std::string name; int64_t begin, end; };
Roi parseRegion(const std::string & text)
{
    Roi r{-1, "", 0, 0x7FFFFFFF};
    const size_t colon = text.rfind(':');
    r.name = text.substr(0, colon);
    if (colon == std::string::npos) return r;
    std::string rest;
    for (char ch : text.substr(colon + 1)) if (ch != ',') rest += ch;
    const size_t dash = rest.find('-');
    const std::string b = rest.substr(0, dash), e = dash == std::string::npos ? "" : rest.substr(dash + 1);
    auto number = [&](const std::string & t) -> int64_t {
        if (t.empty() || t.find_first_not_of("0123456789") != std::string::npos) die("Error while parsing genomic region '" + text + "'");
        return atoll(t.c_str());
    };
    if (!b.empty()) r.begin = std::max<int64_t>(number(b) - 1, 0);
    if (!e.empty()) r.end = number(e);
    return r;
}

struct Options {
    std::vector<std::string> files;
    std::string out = "popdel.vcf";
    bool windowWise = false, outputFailed = false, smoothing = true, uncompressed = false, somatic = false, perSampleRgid = false;
    std::string maxLoadFile, roiFile;
    std::vector<std::string> regions;
    long minInit = -1, minLen = -1;
    unsigned maxLoad = 100, buffer = 200000, iterations = 15, pseudo = 500;
    double prior = 0.0001, minSampleFraction = 0.1, minCover = 0.5;
    int device = 0;
    std::vector<int> devices;            // -g 0,1,...: window-range sharding, one scan context (and host thread) per entry
};

}  // namespace

int main(int argc, char ** argv)
{
    Options opt;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) die("missing value for " + a); return argv[++i]; };
        if (a == "-o" || a == "--out") opt.out = val();
        else if (a == "-n" || a == "--no-regenotyping") opt.windowWise = true;
        else if (a == "-F" || a == "--output-failed") opt.outputFailed = true;
        else if (a == "-u" || a == "--unsmoothed") opt.smoothing = false;
        else if (a == "-x" || a == "--uncompressed-in") opt.uncompressed = true;
        else if (a == "-C" || a == "--cell-population") opt.somatic = true;
        else if (a == "-l" || a == "--min-init-length") opt.minInit = atol(val().c_str());
        else if (a == "-m" || a == "--min-length") opt.minLen = atol(val().c_str());
        else if (a == "-a" || a == "--active-coverage") opt.maxLoad = (unsigned)atol(val().c_str());
        else if (a == "-b" || a == "--buffer-size") opt.buffer = (unsigned)atol(val().c_str());
        else if (a == "-t" || a == "--iterations") opt.iterations = (unsigned)atol(val().c_str());
        else if (a == "-f" || a == "--pseudocount-fraction") opt.pseudo = (unsigned)atol(val().c_str());
        else if (a == "-p" || a == "--prior-probability") opt.prior = atof(val().c_str());
        else if (a == "-s" || a == "--min-sample-fraction") opt.minSampleFraction = atof(val().c_str());
        else if (a == "-c" || a == "--min-relative-window-cover") opt.minCover = atof(val().c_str());
        else if (a == "-g" || a == "--gpu") {                        // one device, or a comma-separated list (an entry may repeat)
            std::string v = val(), tok;
            opt.devices.clear();
            std::stringstream ss(v);
            while (std::getline(ss, tok, ',')) if (!tok.empty()) opt.devices.push_back(atoi(tok.c_str()));
            if (opt.devices.empty()) die("-g needs a device number or a comma-separated list");
            opt.device = opt.devices[0];
        }
        else if (a == "-A" || a == "--active-coverage-file") opt.maxLoadFile = val();
        else if (a == "-e" || a == "--per-sample-rgid") opt.perSampleRgid = true;
        else if (a == "-d" || a == "--max-deletion-size") val();           // parsed by the reference's call parser too, used by `popdel profile` only
        else if (a == "-r" || a == "--region-of-interest") opt.regions.push_back(val());
        else if (a == "-R" || a == "--ROI-file") opt.roiFile = val();
        else if (!a.empty() && a[0] == '-') die("unknown option " + a);
        else opt.files.push_back(a);
    }
    if (opt.files.empty()) die("usage: popdel_b200_call [options] PROFILE-LIST-FILE | PROFILE1 PROFILE2 ...");
    if (opt.files.size() == 1) {                                     // a list of profile paths, one per line
        std::ifstream lf(opt.files[0]);
        if (!lf.good()) die("cannot open '" + opt.files[0] + "'");
        std::string first7(7, '\0');
        lf.read(&first7[0], 7);
        if (first7 != std::string("POPDEL\1", 7)) {
            lf.clear(); lf.seekg(0);
            std::vector<std::string> listed; std::string name;        // whitespace-separated, duplicates ignored (utils_popdel.h:879-912)
            std::set<std::string> seenFiles;
            while (lf >> name) {
                if (seenFiles.insert(name).second) listed.push_back(name);
                else std::cout << "WARNING: duplicate file " << name << ". Ignoring additional occurences." << std::endl;
            }
            opt.files = listed;
        }
    }
    if (opt.files.empty()) die("no profiles given (the list file is empty)");
    const size_t N = opt.files.size();
    std::vector<Profile> & profiles = *new std::vector<Profile>(N);  // never destroyed: the process exits right after the output is written
    StageTimer tm;
    const bool dryRun = opt.device < 0;                               // -g -1: host path only (decode, segments, packing), no scan, no records
    std::thread warm([&] { if (!dryRun && !getenv("PD_NO_WARM")) pd_device_warmup(opt.device); });      // CUDA context creation overlaps the profile decoding
    // decode the profiles with all host cores (one file per task; SURVEY.md 8f rank 2: the reference re-opens and
    // inflates every file per 200-kbp segment, single-threaded)
    const unsigned hostCores = getenv("PD_THREADS") ? (unsigned)std::max(1, atoi(getenv("PD_THREADS"))) : std::thread::hardware_concurrency();
    const unsigned perFile = (unsigned)std::max<size_t>(1, hostCores / std::max<size_t>(N, 1));      // fewer files than cores: split the files
    parallelFor(N, [&](size_t i) { loadHeader(opt.files[i], profiles[i]); });
    tm.lap("headers");
    auto debugDecode = [&]() {
    if (getenv("PD_DEBUG_DECODE"))                                    // checksums of the decoded images (decode paths must agree)
        for (size_t i = 0; i < N; ++i) {
            const Profile & p = profiles[i];
            uint64_t a = 0, b = 0, c2 = 0, e = 0;
            for (uint32_t v : p.recPos) a += v;
            for (int32_t v : p.recDev) b += (uint64_t)(int64_t)v;
            for (uint64_t v : p.winOff) c2 += v;
            for (size_t w = 0; w < p.numWins(); ++w) e += p.winBegin[w] * 31ull + (uint64_t)p.winChrom[w];
            fprintf(stderr, "[popdel_b200] decode %zu: %zu windows %zu read pairs sums %llu %llu %llu %llu\n", i, p.numWins(), p.recPos.size(),
                    (unsigned long long)a, (unsigned long long)b, (unsigned long long)c2, (unsigned long long)e);
        }
    };

    // ---- histograms and parameters (reference parameter_calculation_popdel_call.h:160-204)
    std::vector<pd_rg> rgs;
    std::vector<std::vector<double>> tables;
    std::vector<std::vector<uint32_t>> sampleRgs(N);
    std::set<std::string> seen;
    std::vector<std::string> rgNames;
    for (size_t i = 0; i < N; ++i) {
        std::vector<std::string> names;
        for (const RgHeader & h : profiles[i].rgs) names.push_back(h.name);
        for (size_t k = 0; k < names.size(); ++k) {
            if (!seen.insert(names[k]).second) {
                // -e (insert_histogram_popdel.h:1041-1052): on the first conflict every read group of this file is renamed
                if (!opt.perSampleRgid) die("duplicate read group '" + names[k] + "' (use unique read-group IDs or -e / --per-sample-rgid)");
                for (std::string & nm : names) {
                    const std::string renamed = sampleName(opt.files[i]) + ":" + nm;
                    std::cout << "WARNING: Internally replacing read group ID '" << nm << "' of file '" << opt.files[i] << "' with '" << renamed << "'" << std::endl;
                    nm = renamed;
                }
                if (!seen.insert(names[k]).second) die("could not de-duplicate read group '" + names[k] + "' of '" + opt.files[i] + "'");
            }
            const RgHeader & h = profiles[i].rgs[k];
            rgNames.push_back(names[k]);
            tables.push_back(h.counts);
            pd_rg r; memset(&r, 0, sizeof(r));
            r.sample = (uint32_t)i; r.median = h.median; r.read_length = h.readLength; r.stddev = h.stddev; r.offset = h.offset;
            r.len = (uint32_t)h.counts.size();
            r.min_prob = pd_process_histogram(tables.back().data(), r.len, r.offset, r.median, r.read_length, opt.smoothing, opt.pseudo,
                                              &r.lower_quantile_dist, &r.upper_quantile_dist);
            sampleRgs[i].push_back((uint32_t)rgs.size());
            rgs.push_back(r);
        }
    }
    const size_t R = rgs.size();
    double meanStddev = 0;
    for (size_t g = 0; g < R; ++g) { rgs[g].values = tables[g].data(); meanStddev += rgs[g].stddev; }
    meanStddev /= R;
    std::vector<unsigned> minInit(R);
    for (size_t g = 0; g < R; ++g) {
        minInit[g] = opt.minInit >= 0 ? (unsigned)opt.minInit : (unsigned)rnd(4 * rgs[g].stddev);
        rgs[g].min_init_del_len = minInit[g];
        rgs[g].max_load = opt.maxLoad == 0 ? 0xFFFFFFFFu : opt.maxLoad;
    }
    if (!opt.maxLoadFile.empty()) {                                 // -A: "ReadGroup maxCov" lines (loadMaxLoad, parameter_calculation :83-155)
        std::ifstream lf(opt.maxLoadFile);
        if (!lf.is_open()) die("Could not open coverage file '" + opt.maxLoadFile + "' for reading.");
        std::map<std::string, unsigned long> loads;
        std::string name, load;
        while (lf >> name) {
            if (!(lf >> load)) die("Could not read maximum coverage for read group '" + name + "'.");
            loads[name] = std::stoul(load);
        }
        for (size_t g = 0; g < R; ++g) {
            auto it = loads.find(rgNames[g]);
            if (it != loads.end()) rgs[g].max_load = it->second == 0 ? 0xFFFFFFFFu : (uint32_t)it->second;
        }
    }
    pd_params prm; memset(&prm, 0, sizeof(prm));
    {
        std::vector<unsigned> v = minInit;
        std::sort(v.begin(), v.end());
        const unsigned n = (unsigned)v.size();
        const unsigned j = (unsigned)std::floor(n * 0.95);
        const unsigned pct = (j != n * 0.95) ? v[j] : v[j - 1];
        prm.min_len = opt.minLen >= 0 ? (uint32_t)opt.minLen : (uint32_t)rnd(1.0 * pct);
    }
    prm.iterations = opt.iterations;
    prm.min_lr = (6.6349 / 2.0) - std::log(opt.prior / (1 - opt.prior));
    prm.min_sample_fraction = opt.minSampleFraction;
    prm.window_size = 30; prm.window_buffer = opt.buffer;
    prm.somatic = opt.somatic; prm.window_wise = opt.windowWise;

    tm.lap("histograms");
    warm.join();
    tm.lap("cuda-init");
    // Window-range sharding (-g 0,1,...; SURVEY.md 8e, workflow_popdel.h:297-366): every context gets the read pairs of the
    // contig and scans a contiguous range of whole segments (cuts at the reference's segment borders anchor + k * buffer, so
    // every processSegment() group of calls -- and its unifyCalls -- lives on one device); the records are written in
    // range order, which is the reference's order. No collective.
    if (opt.devices.empty()) opt.devices.push_back(opt.device);
    const size_t K = dryRun ? 1 : opt.devices.size();
    std::vector<pd_ctx *> ctxs(K, nullptr);
    for (size_t k = 0; k < K; ++k) {
        ctxs[k] = pd_create(&prm, (uint32_t)N, (uint32_t)R, rgs.data(), dryRun ? opt.device : opt.devices[k]);
        if (!ctxs[k]) die(std::string("cannot create the scan context: ") + pd_create_error());
    }
    pd_ctx * ctx = ctxs[0];
    auto checkc = [&](pd_ctx * cx, int rc) { if (rc != 0) die(std::string("scan library: ") + pd_last_error(cx)); };
    auto check = [&](int rc) { checkc(ctx, rc); };
    for (size_t k = 0; k < K; ++k) {
        checkc(ctxs[k], pd_set_staging(ctxs[k], 0));                 // one-shot process: pageable staging (see include/popdel_b200.h)
        if (!opt.windowWise && !dryRun) {                            // unifyCalls per segment, on the device
            pd_unify_params up; memset(&up, 0, sizeof(up));
            up.mean_stddev = meanStddev; up.min_relative_window_cover = opt.minCover; up.output_failed = opt.outputFailed;
            checkc(ctxs[k], pd_set_unify(ctxs[k], &up));
        }
    }

    std::ofstream out(opt.out);
    if (!out.good()) die("cannot open output '" + opt.out + "'");
    writeHeader(out, profiles[0], opt.files);
    const QualMap qm;
    const uint32_t WB = opt.buffer;
    uint64_t totalWindows = 0, totalCalls = 0;

    // ---- regions of interest (initializeRois, load_profile_popdel_call.h:108-157): -r / -R regions per contig, sorted,
    // overlapping ones merged, in the contig order of the first profile; default = every contig as a whole
    std::vector<Roi> rois;
    {
        std::vector<std::string> texts = opt.regions;
        if (!opt.roiFile.empty()) {
            std::ifstream rf(opt.roiFile);
            if (!rf.is_open()) die("Could not open ROI file '" + opt.roiFile + "' for reading.");
            std::string line;
            while (rf >> line) texts.push_back(line);
        }
        const std::vector<std::string> & names = profiles[0].contigNames;
        if (texts.empty()) for (size_t c = 0; c < names.size(); ++c) rois.push_back(Roi{(int32_t)c, names[c], 0, 0x7FFFFFFF});
        for (const std::string & t : texts) {
            Roi r = parseRegion(t);
            for (size_t c = 0; c < names.size(); ++c) if (names[c] == r.name) r.contig = (int32_t)c;
            if (r.contig < 0) die("Invalid chromosome name '" + r.name + "' in region of interest.");
            rois.push_back(r);
        }
        std::stable_sort(rois.begin(), rois.end(), [](const Roi & a, const Roi & b) {
            if (a.contig != b.contig) return a.contig < b.contig;
            if (a.begin != b.begin) return a.begin < b.begin;
            return a.end < b.end;
        });
        std::vector<Roi> merged;
        for (const Roi & r : rois) {
            if (!merged.empty() && merged.back().contig == r.contig && merged.back().end >= r.begin) merged.back().end = std::max(merged.back().end, r.end);
            else merged.push_back(r);
        }
        rois.swap(merged);
    }
    for (const Roi & roi : rois) {
        const int32_t c = roi.contig;
        // first 30-bp window over all samples (getFirstWindowCoordinate / getFirstWindowOnNextROI, load_profile :291-412): every
        // sample jumps through its index to the region's begin and contributes the first bucket of the 256-bp window it
        // lands on -- which may lie BEFORE the region (index granularity); the window grid and the segment borders of
        // this region are anchored there
        // decode this region of every profile (one pass per file and region on all cores; the reference
        // re-opens, seeks and inflates every file once per 200-kbp segment, single-threaded)
        parallelFor(N, [&](size_t i) { loadBody(profiles[i], opt.uncompressed, c, (uint64_t)roi.begin, (uint64_t)roi.end, perFile); });
        tm.lap("decode");
        debugDecode();
        bool found = false; uint32_t anchor = 0xFFFFFFFFu;
        for (size_t i = 0; i < N; ++i) {
            const Profile & p = profiles[i];
            if ((size_t)c >= p.contigNames.size()) continue;
            size_t w = indexSeek(p, c, (uint32_t)roi.begin);
            if (w >= p.contigFirst[c + 1]) continue;
            const uint32_t base = bucketBase(p, w);
            if (base == 0xFFFFFFFFu) continue;
            found = true; anchor = std::min(anchor, base);
        }
        if (found && (int64_t)anchor >= roi.end) found = false;       // nothing inside the region (adaptRegions :218-236)
        if (!found) continue;
        for (size_t k = 0; k < K; ++k) checkc(ctxs[k], pd_contig_begin(ctxs[k], anchor));
        tm.lap("create");

        // segment loader: which read pairs does the reference load in which iteration of its segment loop
        // (workflow_popdel.h:297-366, readSegment load_profile :526-587). Iteration t accepts buckets below
        // R = anchor + (t+1)*buffer; every sample re-enters through the index at the smallest bucket any sample
        // stopped at, so read pairs between that position and the next 10-kbp index boundary of a straddling
        // 256-bp window are never loaded.
        // per read group: the read pairs the reference loads, in order. Raw arrays sized once (all read pairs of the read
        // group on this contig); the fill counts sit on cache lines of their own (threads fill neighbouring read groups)
        struct alignas(64) RgOut { uint32_t * pos = nullptr; int32_t * dev = nullptr; uint64_t n = 0; };
        std::vector<RgOut> rgOut(R);
        parallelFor(N, [&](size_t i) {                               // one allocation per read group: its read pairs on this contig
            const Profile & p = profiles[i];
            if ((size_t)c >= p.contigNames.size()) return;
            for (size_t r = 0; r < p.nrg; ++r) {
                uint64_t n = 0;
                for (size_t w = p.contigFirst[c]; w < p.contigFirst[c + 1]; ++w) n += p.winOff[w * p.nrg + r + 1] - p.winOff[w * p.nrg + r];
                RgOut & o = rgOut[sampleRgs[i][r]];
                o.pos = (uint32_t *)malloc(std::max<uint64_t>(n, 1) * 4); o.dev = (int32_t *)malloc(std::max<uint64_t>(n, 1) * 4);
                if (!o.pos || !o.dev) die("out of memory");
            }
        });
        std::vector<uint32_t> cand(N, anchor);
        std::vector<char> fin(N, 0);
        // adaptRegions (:218-236): the region's left end follows the current window once that window reaches it
        uint32_t rb = (int64_t)anchor + 29 > roi.begin ? anchor : (uint32_t)roi.begin;
        const uint64_t roiEnd = (uint64_t)roi.end;
        uint64_t Rt = (uint64_t)anchor + WB;
        while (true) {
            std::atomic<uint32_t> nextRead{0xFFFFFFFFu};
            std::atomic<bool> allFin{true};
            parallelFor(N, [&](size_t i) {                           // samples are independent within one iteration
                if (fin[i]) return;
                if ((uint64_t)cand[i] >= Rt) { allFin = false; return; }                // sits out this iteration
                const Profile & p = profiles[i];
                if ((size_t)c >= p.contigNames.size()) { fin[i] = 1; return; }
                size_t w = indexSeek(p, c, rb);
                bool stopped = false;
                for (; !stopped; ++w) {
                    if (w >= p.contigFirst[c + 1]) { fin[i] = 1; break; }            // next contig or end of file
                    if ((uint64_t)p.winBegin[w] + 255 < rb) continue;
                    // the window's 30-bp buckets in ascending order: those below rb are skipped, the first one at or beyond
                    // Rt ends this sample's iteration; per read group the read pairs are sorted, so both are range scans
                    const uint32_t base = bucketBase(p, w);
                    if (base == 0xFFFFFFFFu) continue;
                    // (a bucket at or beyond the region's end finishes the sample, readSegment :553-559)
                    const uint64_t lim = std::min<uint64_t>(Rt, roiEnd);
                    uint32_t stopAt = 0xFFFFFFFFu;
                    for (size_t r = 0; r < p.nrg; ++r) {
                        RgOut & o = rgOut[sampleRgs[i][r]];
                        uint32_t * dp = o.pos; int32_t * dd = o.dev;
                        uint64_t m = o.n;
                        for (uint64_t k = p.winOff[w * p.nrg + r], e = p.winOff[w * p.nrg + r + 1]; k < e; ++k) {
                            const uint32_t bb = base + ((p.recPos[k] - base) / 30) * 30;
                            if ((uint64_t)bb >= lim) { stopAt = std::min(stopAt, bb); break; }
                            if (bb < rb) continue;
                            dp[m] = p.recPos[k]; dd[m] = p.recDev[k]; ++m;
                        }
                        o.n = m;
                    }
                    if (stopAt != 0xFFFFFFFFu) {
                        if ((uint64_t)stopAt >= roiEnd) { fin[i] = 1; break; }
                        cand[i] = stopAt;
                        uint32_t seen = nextRead.load();
                        while (stopAt < seen && !nextRead.compare_exchange_weak(seen, stopAt)) {}
                        stopped = true;
                    }
                }
                if (!fin[i]) allFin = false;
            });
            if (allFin) break;
            if (nextRead != 0xFFFFFFFFu) rb = nextRead;
            Rt += WB;
        }
        tm.lap("segments");
        if (getenv("PD_DEBUG_DECODE"))
            for (size_t g = 0; g < R; ++g) {
                uint64_t a = 0, b = 0;
                for (uint64_t k = 0; k < rgOut[g].n; ++k) { a += rgOut[g].pos[k]; b += (uint64_t)(int64_t)rgOut[g].dev[k]; }
                fprintf(stderr, "[popdel_b200] pushed rg %zu anchor %u: %llu read pairs sums %llu %llu\n", g, anchor, (unsigned long long)rgOut[g].n,
                        (unsigned long long)a, (unsigned long long)b);
            }
        {   // the add() loop of every read group (active-coverage cap + packing), one read group per task
            std::atomic<int> bad{0};
            parallelFor(R * K, [&](size_t t) {
                const size_t g = t % R, k = t / R;
                if (pd_contig_push(ctxs[k], (uint32_t)g, rgOut[g].n, rgOut[g].pos, rgOut[g].dev) != 0) bad = 1;
            });
            for (size_t g = 0; g < R; ++g) { free(rgOut[g].pos); free(rgOut[g].dev); }
            if (bad) {
                for (size_t k = 0; k < K; ++k) if (pd_last_error(ctxs[k])[0]) checkc(ctxs[k], -1);
                die("scan library: pd_contig_push failed");
            }
        }
        tm.lap("push");

        std::vector<pd_result> results(K);
        for (auto & r : results) memset(&r, 0, sizeof(r));
        if (!dryRun) {
            if (K == 1) check(pd_contig_scan(ctx, 0, 0, &results[0]));
            else {
                uint64_t nWin = 0;
                check(pd_contig_window_count(ctx, &nWin));
                // ranges of whole segments: segment j = windows with anchor-relative position in [j * buffer, (j + 1) * buffer)
                const uint64_t nSeg = nWin ? (30 * (nWin - 1)) / WB + 1 : 0;
                auto segFirst = [&](uint64_t j) -> uint64_t { return j == 0 ? 0 : ((uint64_t)j * WB - 1) / 30 + 1; };
                std::vector<std::thread> th;
                std::vector<int> rcs(K, 0);
                for (size_t k = 0; k < K; ++k) {
                    const uint64_t s0 = nSeg * k / K, s1 = nSeg * (k + 1) / K;
                    const uint64_t w0 = std::min(segFirst(s0), nWin), w1 = s1 >= nSeg ? nWin : std::min(segFirst(s1), nWin);
                    if (s1 <= s0 || w1 <= w0) continue;
                    th.emplace_back([&, k, w0, w1] { rcs[k] = pd_contig_scan(ctxs[k], w0, w1 - w0, &results[k]); });
                }
                for (auto & t : th) t.join();
                for (size_t k = 0; k < K; ++k) checkc(ctxs[k], rcs[k]);
            }
        }
        for (const auto & r : results) totalWindows += r.n_windows;
        tm.lap("scan");
        // window calls (-n) or the merged variants of every segment, already in output order; records are formatted on
        // all cores and written in order
        const std::string & chrom = profiles[0].contigNames[c];
        for (size_t rk = 0; rk < K; ++rk) {
        const pd_result & res = results[rk];
        std::vector<size_t> keep;
        for (size_t k = 0; k < res.n_calls; ++k) {
            const pd_call & pc = res.calls[k];
            if (pc.iterations != 0 && (opt.outputFailed || (pc.filter & 31u) == 0)) keep.push_back(k);
        }
        const size_t CHUNK = 64;
        std::vector<std::string> text((keep.size() + CHUNK - 1) / CHUNK);
        parallelFor(text.size(), [&](size_t t) {
            for (size_t q = t * CHUNK; q < std::min(keep.size(), (t + 1) * CHUNK); ++q) {
                const size_t k = keep[q];
                formatRecord(text[t], chrom, res.calls[k], res.per_sample + k * 13ull * N, N,
                             res.significant_windows ? res.significant_windows[k] : 0, qm);
            }
        });
        for (const std::string & t : text) out.write(t.data(), (std::streamsize)t.size());
        totalCalls += keep.size();
        }
        out.flush();
        tm.lap("vcf");
    }
    out.close();
    tm.report();
    std::cout << "[popdel_b200] scanned " << totalWindows << " windows x " << N << " samples, wrote " << totalCalls
              << " records to '" << opt.out << "'" << std::endl;
    // the output is complete: leave without tearing down the CUDA context and the decoded profiles one allocation at a time
    fflush(nullptr);
    _exit(0);
}
