// oracle/popdel_oracle.cpp -- TEST INFRASTRUCTURE ONLY. See popdel_oracle.h for the rules of use.
//
// CPU restatement of the reference `popdel call` scan. Floating point follows the reference's types
// (long double where the reference uses it) so that results agree with oracle/_ref to the last printed digit.
// All "ref:" citations are relative to /root/reference.
#include "popdel_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <vector>
#include <zlib.h>

namespace {

const uint32_t U32MAX = 0xFFFFFFFFu;

inline int rnd(double d) { return (int)std::floor(d + 0.5); }       // ref: utils_popdel.h:1410-1413

// ------------------------------------------------------------------------------------------------
// Histogram preprocessing (ref: insert_histogram_popdel.h:974-986 and the helpers it calls)
// ------------------------------------------------------------------------------------------------
struct Hist {
    std::vector<double> values;
    double min_prob = 0, stddev = 0;
    int offset = 0;
    unsigned median = 0, readLength = 0, lowerQ = 0, upperQ = 0;
};

// ref: insert_histogram_popdel.h:759-778 -- weights exp(-j*j/40), j in [-20,20]; the denominator sums all
// 41 weights even where the window leaves the histogram.
void smooth(std::vector<double> & v)
{
    int len = (int)v.size();
    std::vector<double> out(len);
    for (int i = 0; i < len; ++i) {
        double ws = 0, s = 0;
        for (int j = -20; j <= 20; ++j) {
            double k = std::exp(-j * j / 40.0);
            if (i + j >= 0 && i + j < len) ws += k * v[i + j];
            s += k;
        }
        out[i] = ws / s;
    }
    v.swap(out);
}

// ref: insert_histogram_popdel.h:609-646 -- `unsigned totalCount += double` truncates at every step; the
// element that satisfies the lower quantile is added a second time when the upper loop starts.
void quantiles(Hist & h)
{
    size_t n = h.values.size();
    unsigned total = 0;
    for (size_t i = 1; i + 1 < n; ++i) total += h.values[i];       // truncating accumulation (quirk)
    double lower = total * 0.01, upper = total * 0.99, sum = 0;
    size_t i = 1;
    for (; i + 1 < n; ++i) {
        sum += h.values[i];
        if (sum >= lower) { h.lowerQ = std::abs((int)i + h.offset - (int)h.median); break; }
    }
    for (; i + 1 < n; ++i) {
        sum += h.values[i];                                        // element i counted twice (quirk)
        if (sum >= upper) { h.upperQ = std::abs((int)i + h.offset - (int)h.median); break; }
    }
}

void process_histogram(Hist & h, bool smoothing, unsigned pseudoCountFraction)
{
    const unsigned windowSize = 256;                               // ref: insert_histogram_popdel.h:1038
    if (smoothing) smooth(h.values);
    quantiles(h);
    {   // densityScale, ref :884-891 -- unsigned accumulator truncates every addition, all entries included
        unsigned total = 0;
        for (double v : h.values) total += v;
        for (double & v : h.values) v /= total;
    }
    {   // normalizeValues, ref :718-745 -- starts at index 1 with insertSize = offset (off by one)
        unsigned insertSize = h.offset;
        for (size_t i = 1; i + 1 < h.values.size(); ++i, ++insertSize) {
            int inner = insertSize - 2 * h.readLength;
            if (inner < 1) inner = 1;
            h.values[i] *= static_cast<double>(windowSize + inner - 1) / windowSize;
        }
    }
    {   // setMinimumProbability, ref :793-808
        double mx = 0;
        for (double v : h.values) if (v > mx) mx = v;
        h.min_prob = mx / pseudoCountFraction;
        for (double & v : h.values) if (v < h.min_prob) v = h.min_prob;
    }
}

inline double I(const Hist & h, int deviation)                     // ref: insert_histogram_popdel.h:1157-1163
{
    int i = deviation + (int)h.median - h.offset;
    if (i <= 0 || i + 1 >= (int)h.values.size()) return h.min_prob;
    return h.values[i];
}

// ------------------------------------------------------------------------------------------------
// Active-set state: a restatement of ChromosomeProfile with its 3-slot cyclic start/end tables
// (ref: popdel_call/profile_structure_popdel_call.h:173-450, 671-948, 1035-1785, 1878-1927).
// Reads of a read group live in one growing array indexed by insertion id; a start "set" is an id range.
// ------------------------------------------------------------------------------------------------
struct Rd { uint32_t pos; int32_t dev; };
struct EndE { uint32_t id; uint32_t lastWin; };

struct RgTab {
    std::vector<Rd> all;                  // id -> read (id = insertion count since the last full reset)
    // start table
    uint32_t sOff[3] = {0, 0, 0}, sCnt[3] = {0, 0, 0}, sRight[3] = {0, 0, 0};
    int sWrite = 0, sRead = 2;
    uint32_t sNext = 0;                   // cursor inside the read set
    // end table
    std::vector<EndE> eSet[3];
    uint32_t eRight[3] = {0, 0, 0};
    int eWrite = 0, eRead = 2, ePosSet = 0;
    uint32_t eNext = 0, dNext = 0, dNextOffset = 0;
    // active ids (reference: std::unordered_set<uint32>); kept in activation order
    std::vector<uint32_t> active;
    uint32_t activeLoad = 0, maxLoad = 100;

    bool startAtEnd() const { return sNext >= sCnt[sRead]; }
    bool endAtEnd() const { return eNext >= eSet[eRead].size(); }
    const Rd & startEntry() const { return all[sOff[sRead] + sNext]; }
    void activate(uint32_t id) { if (std::find(active.begin(), active.end(), id) == active.end()) active.push_back(id); }
    void deactivate(uint32_t id) { auto it = std::find(active.begin(), active.end(), id); if (it != active.end()) active.erase(it); }
};

struct Profile {
    std::vector<RgTab> rg;
    uint32_t numWindows = 200000;        // windowBuffer
    uint32_t currentPos = U32MAX, currentRightBorder = 0;
    bool profilesAtEnd = false;

    // ref :1717-1748
    void resetTo(uint32_t pos)
    {
        for (RgTab & t : rg) {
            for (int k = 0; k < 3; ++k) {
                t.sOff[k] = (uint32_t)t.all.size(); t.sCnt[k] = 0; t.sRight[k] = pos + (k + 1) * numWindows;
                t.eSet[k].clear(); t.eRight[k] = pos + (k + 1) * numWindows;
            }
            t.sWrite = 0; t.sRead = 2; t.sNext = 0;
            t.eWrite = 0; t.eRead = 2; t.eNext = 0; t.ePosSet = 0; t.dNext = 0; t.dNextOffset = 0;
            t.activeLoad = 0;
        }
        currentRightBorder = rg.empty() ? 0 : rg[0].sRight[2];
    }
    // ref :1753-1764 (active sets are not cleared by the reference either)
    void fullReset()
    {
        // NOTE: the reference leaves stale ids in activeReads here (they would index freed tables); we drop them.
        for (RgTab & t : rg) { t.all.clear(); t.active.clear(); }
        resetTo(0);
        currentPos = U32MAX; profilesAtEnd = false;
        currentRightBorder = rg.empty() ? 0 : rg[0].sRight[2];
    }
    // --- end table helpers
    static void endSwitchWrite(RgTab & t, uint32_t nw)                    // ref :713-721
    {
        int nxt = (t.eWrite + 1) % 3, nn = (t.eWrite + 2) % 3;
        t.eSet[nn].clear(); t.eRight[nn] = t.eRight[nxt] + nw;
        t.eWrite = nxt;
    }
    static void endCorrectConsecutive(RgTab & t)                           // ref :722-737
    {
        if (t.ePosSet != (t.eWrite + 2) % 3) { t.ePosSet = t.eWrite; t.dNext = 0; t.activeLoad = 0; }
        else if (t.eSet[t.ePosSet].empty()) t.ePosSet = t.eWrite;
    }
    static void endSwitchRead(RgTab & t) { t.eRead = (t.eRead + 1) % 3; t.eNext = 0; }   // ref :743-752
    static void endAdd(RgTab & t, uint32_t id, uint32_t endPos)            // ref :789-803
    {
        uint32_t lastWin = (endPos / 30) * 30;
        int which = t.eWrite;
        if (lastWin >= t.eRight[t.eWrite]) which = (which + 1) % 3;
        std::vector<EndE> & s = t.eSet[which];
        auto at = std::upper_bound(s.begin(), s.end(), lastWin,
                                   [](uint32_t v, const EndE & e) { return v < e.lastWin; });
        s.insert(at, EndE{id, lastWin});
    }
    static uint32_t endCount(RgTab & t, uint32_t pos)                      // ref :870-928 (getEndCount)
    {
        bool lookAhead = false;
        uint32_t i = t.dNext, c = 0;
        int cSet = t.ePosSet;
        uint32_t n = (uint32_t)t.eSet[cSet].size();
        if (i == n) {
            if (t.ePosSet != t.eWrite) {
                t.ePosSet = (t.ePosSet + 1) % 3; t.dNext = t.dNextOffset; t.dNextOffset = 0;
                i = t.dNext; cSet = t.ePosSet; n = (uint32_t)t.eSet[cSet].size();
            } else return 0;
        }
        if (n == 0) return 0;
        uint32_t win = (pos / 30) * 30;
        while (t.eSet[cSet].at(i).lastWin < win) {
            if (lookAhead) ++t.dNextOffset;
            ++c; ++i;
            if (i == n) {
                if (t.ePosSet != t.eWrite) {
                    t.ePosSet = (t.ePosSet + 1) % 3; t.dNext = t.dNextOffset; t.dNextOffset = 0;
                    return c;
                } else {
                    cSet = (cSet + 1) % 3; n = (uint32_t)t.eSet[cSet].size();
                    if (n == 0) break;
                    i = t.dNextOffset; lookAhead = true;
                }
            }
        }
        t.dNext += c - t.dNextOffset;
        return c;
    }
    // --- start table helpers
    void startSwitchBoth(RgTab & t)                                        // ref :261-269, 275-284, 312-318
    {
        int nxt = (t.sWrite + 1) % 3;
        t.sOff[nxt] = (uint32_t)t.all.size(); t.sCnt[nxt] = 0; t.sRight[nxt] = t.sRight[t.sWrite] + numWindows;
        t.sWrite = nxt;
        t.sRead = (t.sRead + 1) % 3; t.sNext = 0;
        currentRightBorder = t.sRight[t.sRead];
    }
    bool needsSwitch(const RgTab & t, uint32_t w) const { return w >= t.sRight[t.sWrite]; }
    bool tooBigForNext(const RgTab & t, uint32_t w) const { return w >= t.sRight[t.sWrite] + numWindows; }

    // ref :1084-1113 (ChromosomeProfile::add)
    void add(uint32_t g, uint32_t startPos, uint32_t endPos, int32_t dev)
    {
        RgTab & t = rg[g];
        uint32_t id = (uint32_t)t.all.size();
        if (t.activeLoad >= t.maxLoad) {
            uint32_t closing = endCount(t, startPos);
            t.activeLoad -= closing;
            if (t.activeLoad < t.maxLoad) {
                t.all.push_back(Rd{startPos, dev}); ++t.sCnt[t.sWrite];
                endAdd(t, id, endPos); ++t.activeLoad;
            }
        } else {
            t.all.push_back(Rd{startPos, dev}); ++t.sCnt[t.sWrite];
            endAdd(t, id, endPos); ++t.activeLoad;
            uint32_t closing = endCount(t, startPos);
            t.activeLoad -= closing;
        }
    }
    // ref :1188-1206
    bool nextEndWindow(RgTab & t)
    {
        if (t.endAtEnd()) return false;
        uint32_t w = t.eSet[t.eRead][t.eNext].lastWin;
        while (w == t.eSet[t.eRead][t.eNext].lastWin) {
            t.deactivate(t.eSet[t.eRead][t.eNext].id);
            ++t.eNext;
            if (t.endAtEnd()) return false;
        }
        return true;
    }
    // ref :1878-1892 / :1898-1910
    void performSwitches(const std::vector<uint32_t> & gs, bool partial)
    {
        for (uint32_t g : gs) {
            RgTab & t = rg[g];
            while (!t.endAtEnd()) nextEndWindow(t);
            startSwitchBoth(t);
            if (!partial) { endSwitchWrite(t, numWindows); endCorrectConsecutive(t); }
            endSwitchRead(t);
        }
    }
    bool checkAllEmpty() const                                             // ref :1915-1927
    {
        for (const RgTab & t : rg)
            if (t.sCnt[t.sWrite] != 0 || !t.eSet[t.eWrite].empty() || !t.active.empty()) return false;
        return true;
    }
    // ref: load_profile_popdel_call.h:419-458
    bool checkAndSwitch(const std::vector<uint32_t> & gs, uint32_t beginPos)
    {
        RgTab & t0 = rg[gs[0]];
        if (tooBigForNext(t0, beginPos) || needsSwitch(t0, beginPos)) { performSwitches(gs, false); return false; }
        if (t0.eRead == t0.eWrite)
            for (uint32_t g : gs) { endSwitchWrite(rg[g], numWindows); endCorrectConsecutive(rg[g]); }
        return true;
    }
    // ref :1120-1163
    void initializeActiveReads()
    {
        for (RgTab & t : rg) {
            t.sNext = 0;
            if (t.startAtEnd()) continue;
            if (t.startEntry().pos < currentPos) currentPos = t.startEntry().pos;
        }
        for (RgTab & t : rg) {
            if (!t.startAtEnd() && t.active.empty()) {
                uint32_t pos = t.startEntry().pos;
                if (pos != currentPos) continue;
                while (!t.startAtEnd() && t.startEntry().pos == pos) { t.activate(t.sOff[t.sRead] + t.sNext); ++t.sNext; }
            }
        }
    }
    // ref :1213-1247
    bool nextWindow(uint32_t shift)
    {
        bool good = false;
        currentPos += shift;
        if (currentPos >= currentRightBorder) { currentPos -= shift; return false; }
        for (RgTab & t : rg) {
            if (!(t.endAtEnd() && t.startAtEnd())) good = true;
            while (!t.startAtEnd() && t.startEntry().pos <= currentPos) {   // nextStartWindow :1169-1182
                uint32_t p = t.startEntry().pos;
                while (!t.startAtEnd() && t.startEntry().pos == p) { t.activate(t.sOff[t.sRead] + t.sNext); ++t.sNext; }
            }
            while (!t.endAtEnd() && t.eSet[t.eRead][t.eNext].lastWin + shift < currentPos) nextEndWindow(t);
        }
        if (!good) currentPos -= shift;
        return good;
    }
    bool isHighCov(uint32_t g) const { return rg[g].active.size() >= rg[g].maxLoad; }   // ref :1072-1075
};

// ------------------------------------------------------------------------------------------------
// Per-window genotyping (ref: popdel_call/genotype_deletion_popdel_call.h)
// ------------------------------------------------------------------------------------------------
struct T3 { long double a, b, c; };
struct D3 { double a, b, c; };

struct CallRec {
    orc_call c;
    std::vector<uint32_t> ps;   // 13 per sample: PL[3] LAD[3] DAD[5] FL[2]
    uint32_t significantWindows = 0;
};

struct Ctx {
    orc_params p;
    std::vector<Hist> hists;
    std::vector<std::vector<uint32_t>> rgs;      // sample -> read groups
    std::vector<uint32_t> minInit;               // per RG
};

int upperHalfMedian(std::vector<int> & v)                               // ref :15-27
{
    std::sort(v.begin(), v.end());
    unsigned n = (unsigned)v.size();
    if (n == 0) return 0;
    if (n < 4) return rnd(v[n - 1]);
    double pos = (3.0 * n + 2.0 + (n % 2)) / 4.0 - 1.0;
    unsigned l = (unsigned)pos;
    double r = pos - l;
    return rnd((1 - r) * v[l] + r * v[l + 1]);
}

std::set<int> initLengths(const Ctx & cx, const Profile & pr, std::vector<bool> & lowCov)   // ref :33-87
{
    std::vector<int> devs;
    for (size_t s = 0; s < cx.rgs.size(); ++s) {
        unsigned cov = 0, high = 0;                                      // ref profile_structure :1443-1478
        for (uint32_t g : cx.rgs[s]) { unsigned c = (unsigned)pr.rg[g].active.size(); cov += c; if (c >= pr.rg[g].maxLoad) high += c; }
        if (cov < 2u) { lowCov[s] = true; continue; }
        std::vector<int> vals;
        for (uint32_t g : cx.rgs[s]) {
            if (pr.isHighCov(g)) continue;
            for (uint32_t id : pr.rg[g].active) vals.push_back(pr.rg[g].all[id].dev);
        }
        if (vals.empty()) continue;
        devs.push_back(upperHalfMedian(vals));
    }
    std::set<int> out;
    if (devs.empty()) return out;
    std::sort(devs.begin(), devs.end());
    int sum = devs[0], n = 1;
    unsigned thr = cx.minInit[0];
    for (unsigned i = 1; i < devs.size(); ++i) {
        if (devs[i - 1] + 50 > devs[i]) { sum += devs[i]; ++n; thr = std::min(thr, cx.minInit[i]); }   // rank-indexed (quirk)
        else { if (sum / n > (int)thr) out.insert(sum / n); sum = devs[i]; n = 1; thr = cx.minInit[i]; }
    }
    if (sum / n > (int)thr) out.insert(sum / n);
    return out;
}

double initFreq(const Ctx & cx, const Profile & pr, unsigned L)          // ref :93-133
{
    unsigned gc = 0, gt = 0;
    for (size_t s = 0; s < cx.rgs.size(); ++s)
        for (uint32_t g : cx.rgs[s]) {
            unsigned n = (unsigned)pr.rg[g].active.size();
            if (n == 0 || n >= pr.rg[g].maxLoad) continue;
            gt += n;
            int wb = std::max((int)L / 2, rnd(L - 2 * cx.hists[g].stddev));
            int we = L + 2 * cx.hists[g].stddev;
            for (uint32_t id : pr.rg[g].active) { int d = pr.rg[g].all[id].dev; if (d > wb && d < we) ++gc; }
        }
    if (gt == 0u) return 0.0;
    return (double)gc / gt;
}

// EM overload, ref :179-253
T3 dataLik(std::vector<T3> & rgWise, const Ctx & cx, const Profile & pr, const std::vector<uint32_t> & sample,
           unsigned L, const std::vector<int> & shifts)
{
    T3 ll = {0, 0, 0};
    for (uint32_t g : sample) {
        T3 & w = rgWise[g];
        w = T3{0, 0, 0};
        if (pr.isHighCov(g)) continue;
        int refShift = shifts[g];
        const Hist & h = cx.hists[g];
        for (uint32_t id : pr.rg[g].active) {
            int d = pr.rg[g].all[id].dev;
            long double ref = I(h, d - refShift);
            long double del = I(h, d - (int)L);
            long double g0 = std::log(ref);
            long double g1 = std::log(ref + del) - std::log(2.0);
            long double g2 = std::log(del);
            w.a += g0; w.b += g1; w.c += g2;
            ll.a += g0; ll.b += g1; ll.c += g2;
        }
        long double m = std::max(std::max(w.a, w.b), w.c);
        w.a = std::exp(w.a - m); w.b = std::exp(w.b - m); w.c = std::exp(w.c - m);
    }
    long double m = std::max(std::max(ll.a, ll.b), ll.c);
    T3 res = {std::exp(ll.a - m), std::exp(ll.b - m), std::exp(ll.c - m)};
    if (res.a == 0 || res.b == 0 || res.c == 0) res = T3{1, 0.0000000001, 0.0000000001};
    if (res.a == res.b && res.a == res.c) res = T3{1, 0.0000000001, 0.0000000001};
    return res;
}

D3 gtPrior(double f, bool somatic)                                       // ref :343-380
{
    const double ps = 0.0000000001;
    D3 g;
    if (!somatic) {
        g.a = std::max((1 - f) * (1 - f), ps); g.b = std::max(2 * f * (1 - f), ps); g.c = std::max(f * f, ps);
    } else if (f <= 0.4) {
        g.a = std::max(1 - 2 * f + ps, ps); g.b = std::max(2 * f - 2 * ps, ps); g.c = ps;
    } else if (f < 0.75) { g.a = ps; g.b = 1.; g.c = ps; }
    else { g.a = ps; g.b = ps; g.c = 1.; }
    return g;
}

// ref :388-462. rgWise[0] drives every read group's reference weights (iterator never advanced, quirk).
unsigned updateLength(const Ctx & cx, const Profile & pr, const std::vector<T3> & dl, const std::vector<T3> & rgWise,
                      const D3 & gt, int L, std::vector<int> & shifts)
{
    double sumDel = 0, wDel = 0;
    const T3 & r0 = rgWise[0];
    for (size_t s = 0; s < cx.rgs.size(); ++s) {
        double aSum = dl[s].a * gt.a + dl[s].b * gt.b + dl[s].c * gt.c;
        double a1 = std::log(dl[s].b) + std::log(gt.b) - std::log(aSum);
        double a2 = std::log(dl[s].c) + std::log(gt.c) - std::log(aSum);
        for (uint32_t g : cx.rgs[s]) {
            if (pr.isHighCov(g)) continue;
            double sumRef = 0, wRef = 0;
            double aSumRg = r0.a * gt.a + r0.b * gt.b + r0.c * gt.c;
            long double l = std::log(r0.a);
            long double gg = std::log(gt.a);
            long double a = std::log(aSumRg);
            long double a0Rg = l + gg - a;
            long double a1Rg = std::log(r0.b) + std::log(gt.b) - std::log(aSumRg);
            const Hist & h = cx.hists[g];
            for (uint32_t id : pr.rg[g].active) {
                int d = pr.rg[g].all[id].dev;
                double del = I(h, d - L);
                double no_del = I(h, d - shifts[g]);
                double pd = std::exp(a1) * del / (del + no_del) + std::exp(a2);
                double prf = std::exp(a1Rg) * no_del / (del + no_del) + std::exp(a0Rg);
                sumDel += pd; sumRef += prf;
                wDel += pd * d; wRef += prf * d;
            }
            double q = wRef / sumRef;
            int sh = (q != q) ? std::numeric_limits<int>::min() : (int)q;   // x86: NaN -> INT_MIN, reset below
            shifts[g] = sh;
            if (shifts[g] > h.stddev || shifts[g] < -1 * h.stddev) shifts[g] = 0;
        }
    }
    if (getenv("ORC_DEBUG_POS") && (uint32_t)atoi(getenv("ORC_DEBUG_POS")) == pr.currentPos)
        fprintf(stderr, "ORC pos %u L %d sumDel %.17g wDel %.17g rgw %.6Lg %.6Lg %.6Lg shifts %d %d %d gt %.6g %.6g %.6g dl0 %.6Lg %.6Lg %.6Lg dl1 %.6Lg %.6Lg %.6Lg dl2 %.6Lg %.6Lg %.6Lg\n", pr.currentPos, L, sumDel, wDel,
                logl(r0.a), logl(r0.b), logl(r0.c), shifts[0], shifts[1], shifts[2], gt.a, gt.b, gt.c,
                logl(dl[0].a), logl(dl[0].b), logl(dl[0].c), logl(dl[1].a), logl(dl[1].b), logl(dl[1].c), logl(dl[2].a), logl(dl[2].b), logl(dl[2].c));
    if (sumDel == 0) return 0;
    double len = wDel / sumDel;
    if (len < 0) return 0;
    return (unsigned)std::round(len);
}

double updateFreq(const std::vector<T3> & dl, const D3 & gt)             // ref :467-485
{
    double sum = 0;
    for (const T3 & d : dl) {
        double p0 = d.a * gt.a, p1 = d.b * gt.b, p2 = d.c * gt.c;
        double pAll = p0 + p1 + p2;
        sum += (p1 + 2 * p2) / pAll;
    }
    return sum / 2.0 / dl.size();
}

double likelihoodRatio(const std::vector<T3> & dl, const D3 & gt)        // ref :490-508
{
    double del = 0, no_del = 0;
    for (const T3 & d : dl) {
        double p0 = d.a * gt.a, p1 = d.b * gt.b, p2 = d.c * gt.c;
        double pAll = p0 + p1 + p2;
        double a0 = p0 / pAll, a1 = p1 / pAll, a2 = p2 / pAll;
        del += std::log(a0 * d.a + a1 * d.b + a2 * d.c);
        no_del += std::log(d.a);
    }
    return del - no_del;
}

struct Dad { unsigned ref = 0, both = 0, between = 0, alt = 0, right = 0; };

// final overload, ref :255-337 (+ assignDad :137-172, profile_structure :1520-1586)
T3 dataLikFinal(T3 & gtLogs, unsigned lad[3], Dad & dad, uint32_t fl[2], std::vector<uint32_t> & suppF,
                std::vector<uint32_t> & suppL, const Ctx & cx, const Profile & pr,
                const std::vector<uint32_t> & sample, unsigned L, const std::vector<int> & shifts)
{
    T3 ll = {0, 0, 0};
    gtLogs = ll;
    int delLower = std::numeric_limits<int>::max(), delUpper = 0;
    for (uint32_t g : sample) {
        if (pr.isHighCov(g)) continue;
        const Hist & h = cx.hists[g];
        int refShift = shifts[g];
        delLower = L - h.lowerQ;
        delUpper = L + h.upperQ;
        for (uint32_t id : pr.rg[g].active) {
            int d = pr.rg[g].all[id].dev;
            int refUpper = h.upperQ;
            if (d > refUpper) { if (d < delLower) ++dad.between; else if (d <= delUpper) ++dad.alt; else ++dad.right; }
            else { if (d < delUpper) ++dad.ref; else ++dad.both; }
            long double ref = I(h, d - refShift);
            long double del = I(h, d - (int)L);
            if (ref >= 2 * del) ++lad[0]; else if (del >= 2 * ref) ++lad[2]; else ++lad[1];
            ll.a += std::log(ref);        gtLogs.a += std::log10(ref);
            ll.b += std::log(ref + del) - std::log(2.0);   gtLogs.b += std::log10(ref + del) - std::log10(2.0);
            ll.c += std::log(del);        gtLogs.c += std::log10(del);
        }
    }
    {   // getActiveReadsFirstLast
        uint32_t mn = U32MAX, mx = 0;
        for (uint32_t g : sample) {
            if (pr.isHighCov(g)) continue;
            const Hist & h = cx.hists[g];
            int med = h.median, drl = 2 * h.readLength;
            for (uint32_t id : pr.rg[g].active) {
                const Rd & r = pr.rg[g].all[id];
                int isz = std::max(0, r.dev + med - drl);
                uint32_t last = r.pos + isz;
                if (r.pos < mn) mn = r.pos;
                if (last > mx) mx = last;
            }
        }
        if (mn == U32MAX) mn = 0;
        fl[0] = mn; fl[1] = mx;
    }
    for (uint32_t g : sample) {   // addSupportFirstLast with the borders of the LAST non-high-cov RG (quirk)
        if (pr.isHighCov(g)) continue;
        const Hist & h = cx.hists[g];
        int med = h.median, drl = 2 * h.readLength;
        for (uint32_t id : pr.rg[g].active) {
            const Rd & r = pr.rg[g].all[id];
            if (r.dev >= delLower && r.dev <= delUpper) {
                int isz = std::max(0, r.dev + med - drl);
                suppF.push_back(r.pos); suppL.push_back(r.pos + isz);
            }
        }
    }
    if (gtLogs.a + gtLogs.b + gtLogs.c == 0.0) return T3{1, 0.0000000001, 0.0000000001};
    long double mg = std::max(std::max(gtLogs.a, gtLogs.b), gtLogs.c);
    long double md = std::max(std::max(ll.a, ll.b), ll.c);
    gtLogs.a -= mg; gtLogs.b -= mg; gtLogs.c -= mg;
    if (gtLogs.a == gtLogs.b && gtLogs.a == gtLogs.c) { gtLogs.a = 0; gtLogs.b = -10; gtLogs.c = -10; }
    T3 res = {std::exp(ll.a - md), std::exp(ll.b - md), std::exp(ll.c - md)};
    if (res.a == 0 || res.b == 0 || res.c == 0) res = T3{1, 0.0000000001, 0.0000000001};
    if (res.a == res.b && res.a == res.c) res = T3{1, 0.0000000001, 0.0000000001};
    return res;
}

// ref :536-730
bool genotypeWindow(std::vector<CallRec> & calls, const Ctx & cx, const Profile & pr, uint32_t segment)
{
    const size_t N = cx.rgs.size();
    std::vector<bool> lowCov(N, false);
    std::set<int> lens = initLengths(cx, pr, lowCov);
    if (lens.empty()) return false;
    bool ret = false;
    std::vector<int> shifts(cx.hists.size());
    for (int L0 : lens) {
        std::fill(shifts.begin(), shifts.end(), 0);
        std::map<int, double> visited;
        double freq = initFreq(cx, pr, L0);
        if (freq == 0) continue;
        std::vector<T3> rgWise(cx.hists.size(), T3{0, 0, 0});
        std::vector<T3> dl(N);
        for (size_t s = 0; s < N; ++s) dl[s] = dataLik(rgWise, cx, pr, cx.rgs[s], L0, shifts);
        D3 gt = gtPrior(freq, cx.p.somatic != 0);
        unsigned len = L0;
        double prevFreq = freq; unsigned prevLen = len; D3 prevGt = gt;
        std::vector<int> prevShifts = shifts;
        uint32_t it = 0;
        while (len >= cx.p.min_len && it < cx.p.iterations) {
            ++it;
            visited[len] = freq;
            prevLen = len; prevFreq = freq; prevGt = gt;
            len = updateLength(cx, pr, dl, rgWise, gt, len, shifts);
            for (size_t s = 0; s < N; ++s) dl[s] = dataLik(rgWise, cx, pr, cx.rgs[s], len, shifts);
            freq = updateFreq(dl, gt);
            if (freq == 0) break;
            gt = gtPrior(freq, cx.p.somatic != 0);
            auto v = visited.find(len);
            if (v != visited.end() && std::fabs(v->second - freq) <= 0.0001) {
                double lr = likelihoodRatio(dl, gt);
                for (size_t s = 0; s < N; ++s) dl[s] = dataLik(rgWise, cx, pr, cx.rgs[s], prevLen, prevShifts);
                double plr = likelihoodRatio(dl, prevGt);
                if (plr > lr) { len = prevLen; freq = prevFreq; shifts = prevShifts; }
                break;
            }
        }
        if (freq < 0.0000000001 || len < cx.p.min_len) continue;
        std::vector<T3> gtLogs(N);
        std::vector<uint32_t> ps(13 * N, 0);
        std::vector<uint32_t> suppF, suppL;
        for (size_t s = 0; s < N; ++s) {
            unsigned lad[3] = {0, 0, 0}; Dad dad; uint32_t fl[2];
            dl[s] = dataLikFinal(gtLogs[s], lad, dad, fl, suppF, suppL, cx, pr, cx.rgs[s], len, shifts);
            uint32_t * o = &ps[13 * s];
            o[3] = lad[0]; o[4] = lad[1]; o[5] = lad[2];
            o[6] = dad.ref; o[7] = dad.both; o[8] = dad.between; o[9] = dad.alt; o[10] = dad.right;
            o[11] = fl[0]; o[12] = fl[1];
        }
        uint32_t sF = 0, sL = 0;                                         // getSuppFirstLast :514-529
        if (!suppF.empty()) {
            double p = 0.8;
            unsigned l = (unsigned)std::round(static_cast<double>(suppF.size() - 1) * p);
            unsigned r = (unsigned)std::round(static_cast<double>(suppL.size() - 1) * (1 - p));
            std::sort(suppF.begin(), suppF.end()); std::sort(suppL.begin(), suppL.end());
            sF = suppF[l]; sL = suppL[r];
        }
        if (sF == 0 && sL == 0) continue;
        double lr = likelihoodRatio(dl, gt);
        if (lr >= cx.p.min_lr) {
            CallRec cr;
            cr.c.initial_length = L0; cr.c.iterations = it; cr.c.deletion_length = len; cr.c.lr = lr; cr.c.frequency = freq;
            cr.c.window_position = pr.currentPos - 1;
            cr.c.position = cx.p.window_wise ? pr.currentPos - 1 : sF;
            cr.c.end_position = cx.p.window_wise ? 0 : sL;
            cr.c.segment = segment;
            for (size_t s = 0; s < N; ++s) {                             // calculatePhredGL utils :1511-1528
                long double g0 = gtLogs[s].a, g1 = gtLogs[s].b, g2 = gtLogs[s].c;
                const long double gTot = std::log10(std::exp(g0) + std::exp(g1) + std::exp(g2));
                g0 = -10 * (g0 - gTot); g1 = -10 * (g1 - gTot); g2 = -10 * (g2 - gTot);
                const long double mn = std::min(std::min(g0, g1), g2);
                ps[13 * s + 0] = (uint32_t)std::round(g0 - mn);
                ps[13 * s + 1] = (uint32_t)std::round(g1 - mn);
                ps[13 * s + 2] = (uint32_t)std::round(g2 - mn);
            }
            uint32_t filter = 0;
            unsigned data = (unsigned)N;
            for (size_t s = 0; s < N; ++s) if (lowCov[s]) { ps[13 * s] = ps[13 * s + 1] = ps[13 * s + 2] = 0; --data; }
            if (!((double)data / N >= cx.p.min_sample_fraction)) filter |= 4;   // utils :707-717
            cr.c.filter = filter;
            cr.ps.swap(ps);
            calls.push_back(std::move(cr));
            ret = true;
        }
    }
    return ret;
}


// ------------------------------------------------------------------------------------------------
// Segment-level merge (ref: utils_popdel.h:237-654)
// ------------------------------------------------------------------------------------------------
inline bool allPass(const CallRec & c) { return (c.c.filter & 31u) == 0; }          // ref :207-214
inline void markInvalid(CallRec & c) { c.c.filter = 255; }                            // ref :227-230

bool delSizeSimilar(unsigned a, unsigned b, double sd, double f = 0.5)               // ref :237-257
{
    unsigned l = std::min(a, b), r = std::max(a, b);
    if (l + 2 * sd >= r) return true;
    return l >= f * r;
}
bool inDelRange(const CallRec & a, const CallRec & b, double sd, double f = 4)       // ref :258-262
{
    return (b.c.position - a.c.position < (std::min(a.c.deletion_length, b.c.deletion_length) + f * sd));
}
bool checkAndExtend(CallRec & a, CallRec & b, double sd)                             // ref :269-280
{
    unsigned aSpan = a.c.end_position - a.c.position, bSpan = b.c.end_position - b.c.position;
    if ((aSpan < a.c.deletion_length || bSpan < b.c.deletion_length) && inDelRange(a, b, sd)) {
        a.c.end_position = b.c.end_position; return true;
    }
    return false;
}
bool enoughOverlap(const CallRec & a, const CallRec & b, double sd, double f = 0.25) // ref :285-299
{
    unsigned aSpan = a.c.end_position - a.c.position, bSpan = b.c.end_position - b.c.position;
    unsigned minLen = std::min(aSpan, bSpan);
    unsigned left = std::max(a.c.position, b.c.position);
    unsigned right = std::min(a.c.position + aSpan, b.c.position + bSpan);
    int overlap = (right - left);
    if (overlap >= f * minLen) return true;
    if (overlap + 2 * sd >= minLen) return true;
    return false;
}
bool similar(CallRec & a, CallRec & b, double sd)                                    // ref :306-320
{
    if (delSizeSimilar(a.c.deletion_length, b.c.deletion_length, sd)) {
        if (enoughOverlap(a, b, sd)) return true;
        return checkAndExtend(a, b, sd);
    }
    return false;
}
bool lowerCall(const CallRec & l, const CallRec & r)                                 // ref :327-343
{
    if (l.c.position != r.c.position) return l.c.position < r.c.position;
    if (l.c.deletion_length != r.c.deletion_length) return l.c.deletion_length < r.c.deletion_length;
    return l.c.lr > r.c.lr;
}

struct MergeAcc {
    std::vector<std::vector<uint32_t>> vals;      // per sample x 8 (LAD3 + DAD5) lists
    std::vector<uint32_t> gsum;                   // per sample x 3
    std::vector<unsigned> starts, sizes;
};

// ref :441-503 (setGenotypes) + :344-366 (setFreqFromGTs) + :512-559 (mergeWindowRange)
void mergeRange(std::vector<CallRec> & calls, size_t start, size_t last, MergeAcc & m, long double & lr,
                unsigned & callCount, unsigned & winCount, unsigned & sigWin, double r)
{
    CallRec & st = calls[start];
    std::sort(m.starts.begin(), m.starts.end()); st.c.position = m.starts[m.starts.size() / 2]; m.starts.clear();
    std::sort(m.sizes.begin(), m.sizes.end()); st.c.deletion_length = m.sizes[m.sizes.size() / 2]; m.sizes.clear();
    st.c.lr = lr / winCount;
    size_t N = st.ps.size() / 13;
    unsigned gwc = 0;
    for (size_t k = start; k < last; ++k) {
        const CallRec & g = calls[k];
        if (g.c.window_position > st.c.position && g.c.window_position - 30 < st.c.position + st.c.deletion_length) {
            for (size_t s = 0; s < N; ++s) {
                for (int j = 0; j < 3; ++j) m.gsum[3 * s + j] += g.ps[13 * s + j];
                for (int j = 0; j < 8; ++j) m.vals[8 * s + j].push_back(g.ps[13 * s + 3 + j]);
            }
            ++gwc;
        }
    }
    if (gwc == 0) { markInvalid(st); return; }
    unsigned alleles = 0;
    for (size_t s = 0; s < N; ++s) {
        uint32_t * g = &m.gsum[3 * s];
        double mn = std::min(std::min(g[0], g[1]), g[2]);
        double ref = static_cast<double>(g[0] - mn) / gwc, het = static_cast<double>(g[1] - mn) / gwc,
               hom = static_cast<double>(g[2] - mn) / gwc;
        g[0] = g[1] = g[2] = 0;
        uint32_t * o = &st.ps[13 * s];
        o[0] = (uint32_t)std::round(ref); o[1] = (uint32_t)std::round(het); o[2] = (uint32_t)std::round(hom);
        for (int j = 0; j < 8; ++j) {
            std::vector<uint32_t> & v = m.vals[8 * s + j];
            std::sort(v.begin(), v.end());
            o[3 + j] = v[v.size() / 2];
            v.clear();
        }
        if (o[1] == o[2]) { if (het > hom) ++o[1]; else ++o[2]; }
        else if (o[0] == o[1]) { if (ref > het) ++o[0]; else ++o[1]; }
        if (o[0] == 0) continue;                 // setFreqFromGTs
        else if (o[1] == 0) alleles += 1;
        else alleles += 2;
    }
    st.c.frequency = static_cast<double>(alleles) / (N * 2);
    st.significantWindows = sigWin;
    if (30.0 * sigWin / st.c.deletion_length < r) st.c.filter |= 16;
    winCount = 1; sigWin = 1; lr = 0.0; ++callCount;
}

// ref :567-654
bool unify(std::vector<CallRec> & calls, double sd, double r, bool outputFailed)
{
    if (calls.size() <= 1u) return false;
    std::stable_sort(calls.begin(), calls.end(), lowerCall);   // NOTE reference uses std::sort (ties: identical keys)
    size_t cur = 0, last = calls.size() - 1;
    if (!outputFailed) {
        while (!allPass(calls[cur])) { if (cur == last) return false; ++cur; }
        if (cur == last) return false;
    }
    size_t first = cur;
    size_t N = calls[cur].ps.size() / 13;
    MergeAcc m;
    m.vals.resize(8 * N); m.gsum.assign(3 * N, 0);
    m.starts.push_back(calls[cur].c.position); m.sizes.push_back(calls[cur].c.deletion_length);
    unsigned callCount = 1, winCount = 1, sigWin = 1;
    long double lr = calls[cur].c.lr;
    size_t it = first + 1;
    while (true) {
        if (similar(calls[cur], calls[it], sd)) {
            if (allPass(calls[it])) { m.starts.push_back(calls[it].c.position); m.sizes.push_back(calls[it].c.deletion_length); ++sigWin; }
            ++winCount;
            lr += calls[it].c.lr;
            markInvalid(calls[it]);
            if (it == last) {
                if (!m.starts.empty()) mergeRange(calls, cur, it, m, lr, callCount, winCount, sigWin, r);
                break;
            }
        } else {
            if (winCount != 1 && !m.starts.empty()) mergeRange(calls, cur, it, m, lr, callCount, winCount, sigWin, r);
            else markInvalid(calls[cur]);
            cur = it;
        }
        if (it != last) ++it;
        else { if (winCount == 1) { --callCount; markInvalid(calls[cur]); } break; }
    }
    std::vector<CallRec> tmp;
    for (size_t k = first; k <= last; ++k) if (calls[k].c.filter != 255) tmp.push_back(std::move(calls[k]));
    calls.swap(tmp);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Read-pair sources: a profile file (256-bp windows behind a 10 kbp index) or in-memory arrays
// ------------------------------------------------------------------------------------------------
struct Win30 { int32_t chrom; uint32_t beginPos; std::vector<std::vector<Rd>> rec; };    // rec per RG of the sample
struct Win256 { int32_t chrom; uint32_t beginPos; std::vector<std::vector<Rd>> rec; };

// ref: popdel_profile/window_podel.h:314-418
bool convertWindow(const Win256 & o, std::vector<Win30> & c)
{
    c.clear();
    uint32_t beginPos = U32MAX;
    for (const auto & r : o.rec) if (!r.empty() && r[0].pos < beginPos) beginPos = r[0].pos;
    beginPos = (beginPos / 30) * 30;
    std::map<uint32_t, size_t> idx;      // window index -> slot (windows with content, ascending)
    for (const auto & r : o.rec) for (const Rd & x : r) idx[(x.pos - beginPos) / 30] = 0;
    if (idx.empty()) return false;
    size_t k = 0;
    for (auto & kv : idx) { kv.second = k++; }
    c.resize(idx.size());
    for (auto & kv : idx) { Win30 & w = c[kv.second]; w.chrom = o.chrom; w.beginPos = beginPos + kv.first * 30; w.rec.resize(o.rec.size()); }
    for (size_t g = 0; g < o.rec.size(); ++g)
        for (const Rd & x : o.rec[g]) c[idx[(x.pos - beginPos) / 30]].rec[g].push_back(x);
    return true;
}

struct Source {
    std::vector<Win256> wins;            // file order
    uint32_t indexRegionSize = 10000;
    size_t cursor = 0;
    // index jump, ref insert_histogram_popdel.h:531-562 + :298-328 (empty regions back-filled with the next offset)
    void jump(int32_t chrom, uint32_t beginPos)
    {
        uint32_t region = beginPos / indexRegionSize;
        cursor = wins.size();
        for (size_t i = 0; i < wins.size(); ++i) {
            const Win256 & w = wins[i];
            if (w.chrom > chrom || (w.chrom == chrom && w.beginPos / indexRegionSize >= region)) { cursor = i; break; }
        }
    }
    const Win256 * next() { return cursor < wins.size() ? &wins[cursor++] : nullptr; }
};

struct FileMeta {
    std::vector<std::string> rgNames;
    std::vector<Hist> hists;             // raw header histograms
    std::vector<std::string> contigNames;
    std::vector<int32_t> contigLengths;
};

bool inflateAll(const std::vector<unsigned char> & in, size_t off, std::vector<unsigned char> & out)
{
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 31) != Z_OK) return false;
    zs.next_in = const_cast<unsigned char *>(in.data() + off);
    zs.avail_in = (uInt)(in.size() - off);
    std::vector<unsigned char> buf(1 << 20);
    while (zs.avail_in > 0) {
        zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size();
        int rc = inflate(&zs, Z_NO_FLUSH);
        out.insert(out.end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
        if (rc == Z_STREAM_END) { if (zs.avail_in == 0) break; inflateReset(&zs); }
        else if (rc != Z_OK) { inflateEnd(&zs); return false; }
    }
    inflateEnd(&zs);
    return true;
}

template <typename T> T rd(const std::vector<unsigned char> & d, size_t & o) { T v; memcpy(&v, d.data() + o, sizeof(T)); o += sizeof(T); return v; }

// ref: insert_histogram_popdel.h:334-525 (header), window_podel.h:211-251 (records)
bool loadProfile(const char * path, bool uncompressed, FileMeta & fm, Source & src)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) return false;
    std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (d.size() < 15 || memcmp(d.data(), "POPDEL\1", 7) != 0) return false;
    size_t o = 7;
    src.indexRegionSize = rd<uint32_t>(d, o);
    uint32_t numRegions = rd<uint32_t>(d, o);
    o += 8ull * numRegions;
    uint32_t nrg = rd<uint32_t>(d, o);
    for (uint32_t i = 0; i < nrg; ++i) {
        uint32_t nl = rd<uint32_t>(d, o);
        fm.rgNames.emplace_back((const char *)d.data() + o, nl - 1); o += nl;
        Hist h;
        h.median = rd<uint32_t>(d, o); h.stddev = rd<double>(d, o); h.readLength = rd<uint32_t>(d, o);
        h.offset = (int)rd<uint32_t>(d, o);
        uint32_t histEnd = rd<uint32_t>(d, o);
        h.values.resize(histEnd - h.offset);
        for (double & v : h.values) v = rd<double>(d, o);
        fm.hists.push_back(h);
    }
    uint32_t nc = rd<uint32_t>(d, o);
    for (uint32_t i = 0; i < nc; ++i) {
        uint32_t nl = rd<uint32_t>(d, o);
        fm.contigNames.emplace_back((const char *)d.data() + o, nl - 1); o += nl;
        fm.contigLengths.push_back(rd<int32_t>(d, o));
    }
    std::vector<unsigned char> body;
    if (uncompressed) body.assign(d.begin() + o, d.end());
    else if (o < d.size() && !inflateAll(d, o, body)) return false;
    size_t b = 0;
    while (b + 8 <= body.size()) {
        Win256 w; w.chrom = (int32_t)rd<uint32_t>(body, b); w.beginPos = rd<uint32_t>(body, b);
        w.rec.resize(nrg);
        for (uint32_t g = 0; g < nrg; ++g) {
            uint32_t n = rd<uint32_t>(body, b);
            w.rec[g].resize(n);
            for (uint32_t i = 0; i < n; ++i) {
                unsigned char offc = body[b]; b += 1;
                int32_t dv = rd<int32_t>(body, b);
                w.rec[g][i] = Rd{w.beginPos + offc, dv};
            }
        }
        src.wins.push_back(std::move(w));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Driver: emulation of popdel_call()'s segment loop (ref: workflow_popdel.h:256-371) over whole contigs
// ------------------------------------------------------------------------------------------------
struct Driver {
    Ctx cx;
    Profile pr;
    std::vector<Source> src;
    std::vector<std::string> contigNames;      // representative (first sample), ref parameter_parsing :168
    double meanStddev = 0, minRelWinCover = 0.5;
    bool outputFailed = false;
    FILE * dump = nullptr;
    bool dumpWindows = false;
    int64_t windowsScanned = 0;
    uint32_t segment = 0;
    std::vector<CallRec> * rawSink = nullptr;   // in-memory API: collects raw window calls
    std::vector<int64_t> * winSink = nullptr;   // in-memory API: per window {pos, ncalls, 0} + per RG {count, sum_dev, sum_pos}

    void dumpCalls(const std::vector<CallRec> & calls, bool merged)
    {
        if (!dump) return;
        for (const CallRec & c : calls) {
            size_t N = c.ps.size() / 13;
            fprintf(dump, "C %u %u %u %a %a %u %u %u %u %u", c.c.initial_length, c.c.iterations, c.c.deletion_length,
                    c.c.lr, c.c.frequency, c.c.window_position, c.c.position, c.c.end_position, c.c.filter, (unsigned)N);
            if (merged) fprintf(dump, " %u", c.significantWindows);
            fprintf(dump, "\n");
            for (size_t s = 0; s < N; ++s) {
                const uint32_t * o = &c.ps[13 * s];
                fprintf(dump, "G %u %u %u %u %u %u %u %u %u %u %u %u %u\n", o[0], o[1], o[2], o[3], o[4], o[5], o[6],
                        o[7], o[8], o[9], o[10], o[11], o[12]);
            }
        }
    }
    void genotypeAndLog(std::vector<CallRec> & calls)
    {
        genotypeWindow(calls, cx, pr, segment);
        ++windowsScanned;
        if (winSink) {
            winSink->push_back(pr.currentPos); winSink->push_back((int64_t)calls.size()); winSink->push_back(0);
            for (const RgTab & t : pr.rg) {
                long long sd = 0; long long sp = 0;
                for (uint32_t id : t.active) { sd += t.all[id].dev; sp += t.all[id].pos; }
                winSink->push_back((int64_t)t.active.size()); winSink->push_back(sd); winSink->push_back(sp);
            }
        }
        if (dump && dumpWindows) {
            fprintf(dump, "W %u %u", pr.currentPos, (unsigned)calls.size());
            for (const RgTab & t : pr.rg) {
                long long sd = 0; unsigned long long sp = 0;
                for (uint32_t id : t.active) { sd += t.all[id].dev; sp += t.all[id].pos; }
                fprintf(dump, " %u %lld %llu", (unsigned)t.active.size(), sd, sp);
            }
            fprintf(dump, "\n");
        }
    }
    // ref: workflow_popdel.h:28-54 / :58-86
    void processSegment(std::vector<CallRec> & calls)
    {
        if (pr.profilesAtEnd) pr.profilesAtEnd = !pr.nextWindow(30);
        pr.initializeActiveReads();
        while (!pr.profilesAtEnd) {
            genotypeAndLog(calls);
            pr.profilesAtEnd = !pr.nextWindow(30);
        }
        if (rawSink) { for (CallRec & c : calls) rawSink->push_back(c); }
        if (!cx.p.window_wise) {
            if (dump) fprintf(dump, "S %u %u\n", segment, (unsigned)calls.size());
            dumpCalls(calls, false);
            bool res = unify(calls, meanStddev, minRelWinCover, outputFailed);
            if (dump) { if (res) { fprintf(dump, "M %u\n", (unsigned)calls.size()); dumpCalls(calls, true); } else fprintf(dump, "M 0\n"); }
        } else {
            if (dump) fprintf(dump, "S %u %u\n", segment, (unsigned)calls.size());
            dumpCalls(calls, false);
        }
        ++segment;
    }
    // ref: load_profile_popdel_call.h:463-480
    void addRecords(const std::vector<uint32_t> & gs, const Win30 & w)
    {
        for (size_t r = 0; r < w.rec.size(); ++r)
            for (const Rd & x : w.rec[r]) {
                const Hist & h = cx.hists[gs[r]];
                int inner = x.dev + (int)h.median - 2 * (int)h.readLength;
                uint32_t endPos = x.pos;
                if (inner > 0) endPos += inner;
                pr.add(gs[r], x.pos, endPos, x.dev);
            }
    }
    // ref: load_profile_popdel_call.h:488-587 (readTillRoi + readSegment). cand = nextCandidateWindows[i].
    // Returns 0 segment full, 2 end of contig/ROI, 3 EOF.
    unsigned readSegment(size_t i, int32_t & candChrom, uint32_t & candPos, int32_t roiChrom, uint32_t roiBegin, uint32_t roiEnd)
    {
        const std::vector<uint32_t> & gs = cx.rgs[i];
        std::vector<Win30> conv;
        while (true) {
            const Win256 * w;
            do {
                w = src[i].next();
                if (!w) { pr.performSwitches(gs, true); candChrom = -1; candPos = U32MAX; return 3; }
                if (w->chrom != candChrom) { pr.performSwitches(gs, true); candChrom = w->chrom; candPos = w->beginPos; return 2; }
            } while (w->beginPos + 255 < roiBegin);
            if (!convertWindow(*w, conv)) return 3;   // reference throws here
            for (const Win30 & it : conv) {
                if (it.chrom != roiChrom || it.beginPos >= roiEnd) {
                    pr.performSwitches(gs, true); candChrom = it.chrom; candPos = it.beginPos; return 2;
                }
                if (it.beginPos < roiBegin) continue;
                RgTab & t0 = pr.rg[gs[0]];
                if (pr.tooBigForNext(t0, it.beginPos)) {
                    if (pr.checkAllEmpty()) pr.resetTo(it.beginPos);
                    else { pr.performSwitches(gs, true); candChrom = it.chrom; candPos = it.beginPos; return 0; }
                } else if (pr.needsSwitch(t0, it.beginPos)) {
                    pr.performSwitches(gs, true); candChrom = it.chrom; candPos = it.beginPos; return 0;
                }
                addRecords(gs, it);
            }
        }
    }
    // ref: load_profile_popdel_call.h:291-412 -- smallest first 30-bp window over all samples for ROI = contig c
    bool firstWindow(int32_t c, uint32_t & pos)
    {
        bool found = false; int32_t bestChrom = 0x7fffffff; uint32_t bestPos = U32MAX;
        std::vector<Win30> conv;
        for (size_t i = 0; i < src.size(); ++i) {
            src[i].jump(c, 0);
            const Win256 * w = src[i].next();
            if (!w) continue;
            if (!convertWindow(*w, conv)) continue;
            int32_t ch = conv[0].chrom; uint32_t p = conv[0].beginPos;
            // lowerCoord(min, cur): keep min when it is lower or equal (rank by contig order, then position)
            if (!found || !(bestChrom < ch || (bestChrom == ch && bestPos <= p))) { bestChrom = ch; bestPos = p; found = true; }
        }
        if (!found || bestChrom != c) return false;
        pos = bestPos;
        return true;
    }
    void run()
    {
        size_t n = src.size();
        std::vector<bool> finishedFiles(n, false), finishedROIs(n, false);
        unsigned fileCount = (unsigned)n;
        int32_t roi = 0; const int32_t nRoi = (int32_t)contigNames.size();
        uint32_t anchor = 0;
        while (roi < nRoi && !firstWindow(roi, anchor)) ++roi;
        if (roi >= nRoi) return;
        uint32_t curWin = anchor, roiBegin = 0; const uint32_t roiEnd = 0x7fffffffu;
        std::vector<int32_t> candChrom(n, roi); std::vector<uint32_t> candPos(n, anchor);
        uint32_t nextReadPos = U32MAX;
        pr.currentPos = anchor; pr.resetTo(anchor);
        std::vector<CallRec> calls;
        while (fileCount != 0) {
            // goNextRegion/adaptRegions, ref load_profile :220-286 (current window is on the ROI's contig here)
            if ((int)curWin >= (int)roiEnd) break;
            if ((int)(curWin + 29) == (int)roiBegin) {}
            else if ((int)(curWin + 29) > (int)roiBegin) roiBegin = curWin;
            else curWin += ((roiBegin - curWin) / 30) * 30;
            for (size_t i = 0; i < n; ++i) {
                if (finishedFiles[i] || finishedROIs[i]) continue;
                src[i].jump(roi, roiBegin);
                if (pr.checkAndSwitch(cx.rgs[i], candPos[i])) {
                    unsigned code = readSegment(i, candChrom[i], candPos[i], roi, roiBegin, roiEnd);
                    if (code == 0) { if (candPos[i] < nextReadPos) nextReadPos = candPos[i]; }
                    else if (code < 3) finishedROIs[i] = true;
                    else { finishedROIs[i] = true; finishedFiles[i] = true; --fileCount; }
                }
            }
            calls.clear();
            processSegment(calls);
            bool all = true; for (bool b : finishedROIs) all = all && b;
            if (all) {
                ++roi;                                            // finalizeRoi, ref workflow :110-138
                if (roi >= nRoi) return;
                pr.fullReset();
                nextReadPos = U32MAX;
                std::fill(finishedROIs.begin(), finishedROIs.end(), false);
                while (roi < nRoi && !firstWindow(roi, anchor)) ++roi;
                if (roi >= nRoi) return;
                curWin = anchor; roiBegin = 0;
                pr.currentPos = anchor; pr.resetTo(anchor);
                for (size_t i = 0; i < n; ++i) { candChrom[i] = roi; candPos[i] = anchor; }
            } else {
                curWin = nextReadPos; nextReadPos = U32MAX;
            }
        }
    }
};

unsigned percentile95(std::vector<unsigned> v)                                        // ref utils :1494-1505
{
    std::sort(v.begin(), v.end());
    unsigned n = (unsigned)v.size();
    double p = 0.95;
    unsigned j = (unsigned)std::floor(n * p);
    bool remainder = j != n * p;
    return remainder ? v[j] : v[j - 1];
}

}  // namespace

extern "C" double orc_process_histogram(double * values, uint32_t len, int32_t offset, uint32_t median,
                                        uint32_t read_length, int smoothing, uint32_t pseudo_count_fraction,
                                        uint32_t * lower_q, uint32_t * upper_q)
{
    Hist h; h.values.assign(values, values + len); h.offset = offset; h.median = median; h.readLength = read_length;
    process_histogram(h, smoothing != 0, pseudo_count_fraction);
    std::copy(h.values.begin(), h.values.end(), values);
    if (lower_q) *lower_q = h.lowerQ;
    if (upper_q) *upper_q = h.upperQ;
    return h.min_prob;
}

// Test hook: the stored / skipped decision of ChromosomeProfile::add (ref :1084-1113) for one read group's position-
// sorted read pairs (start, end = start + max(0, inner distance), both relative to the position the tables were reset
// to), with the end-table switches performSwitches (ref :1880-1893) applies at every segment border up to the read
// pair's segment. stored[i] = 1 when read pair i entered the tables.
extern "C" int orc_cap_replay(uint32_t window_buffer, uint32_t max_load, uint64_t n, const uint32_t * start, const uint32_t * end,
                              uint8_t * stored)
{
    Profile pr;
    pr.rg.resize(1);
    pr.numWindows = window_buffer;
    pr.rg[0].maxLoad = max_load;
    pr.resetTo(0);
    RgTab & t = pr.rg[0];
    uint64_t seg = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t j = (uint64_t)(start[i] / 30 * 30) / window_buffer;
        for (; seg < j; ++seg) { Profile::endSwitchWrite(t, window_buffer); Profile::endCorrectConsecutive(t); }
        const size_t before = t.all.size();
        pr.add(0, start[i], end[i], 0);
        stored[i] = t.all.size() != before;
    }
    return 0;
}

extern "C" int64_t orc_scan_contig(const orc_params * p, uint32_t n_samples, uint32_t n_rg, const orc_rg * rgs,
                                   const uint64_t * rg_off, const uint32_t * pos, const int32_t * dev,
                                   uint32_t anchor, uint32_t last_pos,
                                   orc_call * calls, uint32_t * per_sample, int64_t max_calls,
                                   int64_t * win_dump, int64_t win_dump_cap, int64_t * n_windows_scanned)
{
    (void)last_pos; (void)anchor;
    Driver d;
    std::vector<int64_t> wins;
    if (win_dump) d.winSink = &wins;
    d.cx.p = *p;
    d.cx.rgs.resize(n_samples);
    d.cx.hists.resize(n_rg);
    d.cx.minInit.resize(n_rg);
    d.pr.rg.resize(n_rg);
    d.pr.numWindows = p->window_buffer;
    for (uint32_t g = 0; g < n_rg; ++g) {
        Hist & h = d.cx.hists[g];
        h.values.assign(rgs[g].values, rgs[g].values + rgs[g].len);
        h.min_prob = rgs[g].min_prob; h.stddev = rgs[g].stddev; h.offset = rgs[g].offset; h.median = rgs[g].median;
        h.readLength = rgs[g].read_length; h.lowerQ = rgs[g].lower_quantile_dist; h.upperQ = rgs[g].upper_quantile_dist;
        d.cx.rgs[rgs[g].sample].push_back(g);
        d.cx.minInit[g] = rgs[g].min_init_del_len;
        d.pr.rg[g].maxLoad = rgs[g].max_load;
    }
    // in-memory source per sample: one pseudo 256-window per 256-bp block (exact index: region size 256)
    d.src.resize(n_samples);
    d.contigNames.push_back("contig");
    for (uint32_t s = 0; s < n_samples; ++s) {
        Source & src = d.src[s];
        src.indexRegionSize = 256;
        const std::vector<uint32_t> & gs = d.cx.rgs[s];
        std::map<uint32_t, size_t> byWin;
        for (size_t r = 0; r < gs.size(); ++r)
            for (uint64_t k = rg_off[gs[r]]; k < rg_off[gs[r] + 1]; ++k) byWin[pos[k] / 256] = 0;
        size_t idx = 0;
        for (auto & kv : byWin) kv.second = idx++;
        src.wins.resize(byWin.size());
        for (auto & kv : byWin) { Win256 & w = src.wins[kv.second]; w.chrom = 0; w.beginPos = kv.first * 256; w.rec.resize(gs.size()); }
        for (size_t r = 0; r < gs.size(); ++r)
            for (uint64_t k = rg_off[gs[r]]; k < rg_off[gs[r] + 1]; ++k)
                src.wins[byWin[pos[k] / 256]].rec[r].push_back(Rd{pos[k], dev[k]});
    }
    std::vector<CallRec> raw;
    d.rawSink = &raw;
    d.cx.p.window_wise = 1;                   // no merge for the in-memory API: raw window calls are the product
    const bool ww = p->window_wise != 0;
    // positions follow the requested mode, merging is skipped either way
    d.cx.p.window_wise = ww ? 1 : 0;
    d.outputFailed = true;
    {
        // run() merges when !window_wise; the raw sink is filled before the merge
        d.run();
    }
    if (n_windows_scanned) *n_windows_scanned = d.windowsScanned;
    if (win_dump) {
        if ((int64_t)wins.size() > 3 * win_dump_cap) return -1;
        std::copy(wins.begin(), wins.end(), win_dump);
    }
    if ((int64_t)raw.size() > max_calls) return -(int64_t)raw.size();
    for (size_t k = 0; k < raw.size(); ++k) {
        calls[k] = raw[k].c;
        memcpy(per_sample + 13ull * n_samples * k, raw[k].ps.data(), sizeof(uint32_t) * 13ull * n_samples);
    }
    return (int64_t)raw.size();
}

// Segment-level merge of window calls given in memory (test hook for the device-side unify of popdel_b200): the calls of
// one processSegment() = a run of equal `segment` tags; every run goes through unify() (ref utils_popdel.h:567-654) and
// the surviving variants are appended to the output in run order. Returns the number of variants.
extern "C" int64_t orc_unify_segments(const orc_call * calls, const uint32_t * per_sample, int64_t n_calls, uint32_t n_samples,
                                      double mean_stddev, double min_cover, int output_failed,
                                      orc_call * out_calls, uint32_t * out_ps, uint32_t * out_sig)
{
    int64_t n_out = 0;
    const size_t row = 13ull * n_samples;
    for (int64_t k = 0; k < n_calls;) {
        std::vector<CallRec> seg;
        const uint32_t sidx = calls[k].segment;
        for (; k < n_calls && calls[k].segment == sidx; ++k) {
            CallRec r;
            r.c = calls[k];
            r.ps.assign(per_sample + row * k, per_sample + row * (k + 1));
            seg.push_back(std::move(r));
        }
        if (!unify(seg, mean_stddev, min_cover, output_failed != 0)) continue;
        for (const CallRec & r : seg) {
            out_calls[n_out] = r.c;
            memcpy(out_ps + row * n_out, r.ps.data(), row * sizeof(uint32_t));
            out_sig[n_out] = r.significantWindows;
            ++n_out;
        }
    }
    return n_out;
}

extern "C" int orc_call_files(const char * const * files, uint32_t n_files, const char * dump_path,
                              int window_wise, int dump_windows, int uncompressed, uint32_t min_init_len,
                              uint32_t max_load, int64_t * n_windows_scanned)
{
    Driver d;
    orc_params & p = d.cx.p;
    p.iterations = 15; p.min_sample_fraction = 0.1; p.window_size = 30; p.window_buffer = 200000;
    p.somatic = 0; p.window_wise = window_wise;
    d.src.resize(n_files);
    d.cx.rgs.resize(n_files);
    std::set<std::string> seen;
    for (uint32_t i = 0; i < n_files; ++i) {
        FileMeta fm;
        if (!loadProfile(files[i], uncompressed != 0, fm, d.src[i])) { fprintf(stderr, "oracle: cannot load %s\n", files[i]); return 1; }
        if (i == 0) d.contigNames = fm.contigNames;
        for (size_t r = 0; r < fm.hists.size(); ++r) {
            if (!seen.insert(fm.rgNames[r]).second) { fprintf(stderr, "oracle: duplicate read group %s\n", fm.rgNames[r].c_str()); return 1; }
            process_histogram(fm.hists[r], true, 500);
            d.cx.rgs[i].push_back((uint32_t)d.cx.hists.size());
            d.cx.hists.push_back(fm.hists[r]);
        }
    }
    size_t R = d.cx.hists.size();
    double sum = 0;
    for (const Hist & h : d.cx.hists) sum += h.stddev;
    d.meanStddev = sum / R;                                           // ref parameter_calculation :71-78
    d.cx.minInit.resize(R);
    for (size_t g = 0; g < R; ++g) d.cx.minInit[g] = min_init_len ? min_init_len : (unsigned)rnd(4 * d.cx.hists[g].stddev);
    p.min_len = (unsigned)rnd(1.0 * percentile95(d.cx.minInit));      // ref :59-67
    p.min_lr = (6.6349 / 2.0) - std::log(0.0001 / (1 - 0.0001));      // ref :16-23
    d.pr.rg.resize(R);
    d.pr.numWindows = p.window_buffer;
    for (size_t g = 0; g < R; ++g) d.pr.rg[g].maxLoad = max_load ? max_load : U32MAX;
    d.dump = fopen(dump_path, "w");
    if (!d.dump) return 2;
    d.dumpWindows = dump_windows != 0;
    d.run();
    fclose(d.dump);
    if (n_windows_scanned) *n_windows_scanned = d.windowsScanned;
    return 0;
}

#ifdef ORACLE_MAIN
int main(int argc, char ** argv)
{
    // usage: popdel_oracle <dump> <window_wise> <dump_windows> file1 file2 ...
    if (argc < 5) { fprintf(stderr, "usage: %s dump window_wise dump_windows files...\n", argv[0]); return 1; }
    int64_t nw = 0;
    int rc = orc_call_files(argv + 4, argc - 4, argv[1], atoi(argv[2]), atoi(argv[3]), 0, 0, 100, &nw);
    fprintf(stderr, "windows scanned: %lld\n", (long long)nw);
    return rc;
}
#endif
