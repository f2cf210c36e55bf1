// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Builds the UNMODIFIED reference `popdel call` (headers included where they lie under
// /root/reference) with two call sites interposed by the preprocessor so that the intermediate
// products of the hot path can be dumped:
//   * genotype_deletion_window (reference popdel_call/genotype_deletion_popdel_call.h:536-730), called
//     once per 30-bp window from processSegment (workflow_popdel.h:42-47): we log the window position,
//     the active-set size / checksums per read group and the number of calls appended;
//   * unifyCalls (utils_popdel.h:567-654), called once per segment (workflow_popdel.h:48): we log the
//     raw window calls of the segment (the observable product of the hot path) before
//     the merge, and the merged calls after it.
// No reference source text is copied: the reference functions are called as they are.
//
// Usage: POPDEL_HARNESS_OUT=dump.txt [POPDEL_HARNESS_WINDOWS=1] popdel_ref_harness <popdel call args...>
// Output lines:
//   W <pos> <calls_after> { <rg_active> <sum_dev> <sum_startpos> }*     (only with POPDEL_HARNESS_WINDOWS=1)
//   S <segment_index> <n_raw_calls>
//   C <initLen> <iterations> <delLen> <LR %a> <freq %a> <windowPos> <pos> <endPos> <filter> <nSamples>
//   G <pl0> <pl1> <pl2> <lad0> <lad1> <lad2> <dad ref both between alt right> <fl0> <fl1>   (one per sample)
//   M <n_merged_calls>  followed by C/G lines (plus a trailing "significantWindows" on the C line)
#include <cstdio>
#include <cstdlib>

#include <seqan/arg_parse.h>
#include <seqan/bam_io.h>

#include "parse_popdel.h"
#include "utils_popdel.h"
#include "popdel_profile/profile_parameter_parsing_popdel.h"
#include "popdel_profile/parameter_estimation_popdel.h"
#include "popdel_profile/bam_window_iterator_popdel.h"
#include "popdel_call/parameter_calculation_popdel_call.h"
#include "popdel_call/genotype_deletion_popdel_call.h"
#include "popdel_call/load_profile_popdel_call.h"
#include "popdel_call/vcfout_popdel_call.h"

static FILE * g_out = NULL;
static bool g_windows = false;
static unsigned g_segment = 0;

static void dumpCalls(const String<Call> & calls, bool merged)
{
    for (unsigned i = 0; i < length(calls); ++i)
    {
        const Call & c = calls[i];
        fprintf(g_out, "C %u %u %u %a %a %u %u %u %u %u", c.initialLength, c.iterations, c.deletionLength,
                c.likelihoodRatio, c.frequency, c.windowPosition, c.position, c.endPosition,
                (unsigned)c.filter, (unsigned)length(c.gtLikelihoods));
        if (merged)
            fprintf(g_out, " %u", c.significantWindows);
        fprintf(g_out, "\n");
        for (unsigned s = 0; s < length(c.gtLikelihoods); ++s)
        {
            fprintf(g_out, "G %u %u %u %u %u %u %u %u %u %u %u %u %u\n",
                    c.gtLikelihoods[s].i1, c.gtLikelihoods[s].i2, c.gtLikelihoods[s].i3,
                    c.lads[s].i1, c.lads[s].i2, c.lads[s].i3,
                    c.dads[s].ref, c.dads[s].both, c.dads[s].between, c.dads[s].alt, c.dads[s].right,
                    c.firstLast[s].i1, c.firstLast[s].i2);
        }
    }
}

inline bool harness_unifyCalls(String<Call> & calls, const double & stddev, const double & r,
                               const bool outputFailed = false)
{
    fprintf(g_out, "S %u %u\n", g_segment, (unsigned)length(calls));
    dumpCalls(calls, false);
    bool res = unifyCalls(calls, stddev, r, outputFailed);
    if (res)
    {
        fprintf(g_out, "M %u\n", (unsigned)length(calls));
        dumpCalls(calls, true);
    }
    else
    {
        fprintf(g_out, "M 0\n");
    }
    ++g_segment;
    return res;
}

inline bool harness_genotype_deletion_window(String<Call> & calls,
                                             const ChromosomeProfile & chromosomeProfiles,
                                             const TRGs & rgs,
                                             PopDelCallParameters & params)
{
    bool res = genotype_deletion_window(calls, chromosomeProfiles, rgs, params);
    if (g_windows)
    {
        fprintf(g_out, "W %u %u", chromosomeProfiles.currentPos, (unsigned)length(calls));
        for (unsigned rg = 0; rg < length(chromosomeProfiles.activeReads); ++rg)
        {
            long long sumDev = 0;
            unsigned long long sumPos = 0;
            for (ChromosomeProfile::TActiveSet::const_iterator it = chromosomeProfiles.activeReads[rg].begin();
                 it != chromosomeProfiles.activeReads[rg].end(); ++it)
            {
                sumDev += chromosomeProfiles.startProfiles[rg].getDeviationAt(*it);
                sumPos += chromosomeProfiles.startProfiles[rg].getStartPosAt(*it);
            }
            fprintf(g_out, " %u %lld %llu", (unsigned)chromosomeProfiles.activeReads[rg].size(), sumDev, sumPos);
        }
        fprintf(g_out, "\n");
    }
    return res;
}

#define unifyCalls harness_unifyCalls
#define genotype_deletion_window harness_genotype_deletion_window
#include "workflow_popdel.h"
#undef unifyCalls
#undef genotype_deletion_window

// Replay mode (POPDEL_HARNESS_CAP_REPLAY=<input file>): drives the reference's own ChromosomeProfile::add for ONE read
// group with the segment switches of performSwitches, and writes one byte per read pair (1 = stored) to the harness
// output. Input: u32 windowBuffer, u32 maxLoad, u64 n, then n u32 start and n u32 end positions (sorted by start).
static int capReplay(const char * inName, const char * outName)
{
    FILE * in = fopen(inName, "rb");
    if (!in) return 2;
    uint32_t hdr[2]; uint64_t n = 0;
    if (fread(hdr, 4, 2, in) != 2 || fread(&n, 8, 1, in) != 1) return 2;
    std::vector<uint32_t> start(n), end(n);
    if (n && (fread(start.data(), 4, n, in) != n || fread(end.data(), 4, n, in) != n)) return 2;
    fclose(in);
    String<unsigned> maxLoad;
    appendValue(maxLoad, hdr[1]);
    ChromosomeProfile profile(1, maxLoad, hdr[0]);
    profile.resetTo(0);
    TReadGroupIndices rg;
    appendValue(rg, 0u);
    std::vector<unsigned char> stored(n);
    uint64_t seg = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
        const uint64_t j = (uint64_t)(start[i] / 30 * 30) / hdr[0];
        for (; seg < j; ++seg)
            performSwitches(profile, rg);
        const unsigned before = profile.startProfiles[0].insertionCount;
        profile.add(0, start[i], end[i], 0);
        stored[i] = profile.startProfiles[0].insertionCount != before;
    }
    FILE * out = fopen(outName, "wb");
    if (!out) return 2;
    fwrite(stored.data(), 1, n, out);
    fclose(out);
    return 0;
}

int main(int argc, char const ** argv)
{
    if (getenv("POPDEL_HARNESS_CAP_REPLAY"))
        return capReplay(getenv("POPDEL_HARNESS_CAP_REPLAY"), getenv("POPDEL_HARNESS_OUT") ? getenv("POPDEL_HARNESS_OUT") : "cap_replay.out");
    const char * outName = getenv("POPDEL_HARNESS_OUT");
    g_out = fopen(outName ? outName : "popdel_harness_dump.txt", "w");
    if (!g_out)
    {
        fprintf(stderr, "cannot open harness output\n");
        return 2;
    }
    g_windows = getenv("POPDEL_HARNESS_WINDOWS") != NULL;
    argv[0] = "popdel call";
    int res = popdel_call(argc, argv);
    fclose(g_out);
    return res;
}
