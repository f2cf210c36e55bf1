// integration/popdel_call_gpu.cpp -- the UNMODIFIED reference `popdel call` with its window loop replaced by libpopdel_b200.
//
// Proof that the C ABI (include/popdel_b200.h) is a drop-in for the reference's hot path ("reference-side binding
// actually compiled"): the reference's own translation unit -- command line, parameter calculation, profile reader and
// segment loader, unifyCalls, VCF writer, all included where they lie under $(REF) -- is built with three call sites
// interposed by the preprocessor (the technique of oracle/ref_harness.cpp; no reference text is copied):
//   * ChromosomeProfile::add / resetTo  (profile_structure_popdel_call.h:1084-1113, :1717-1748) through a derived class that the
//     loader and the workflow see instead: every read pair the loader hands to the scan state is also recorded per read
//     group (= pd_contig_push), every resetTo opens a new contig batch (= pd_contig_begin);
//   * genotype_deletion_window  (genotype_deletion_popdel_call.h:536-730): NOT called any more -- the hook appends the window
//     calls the GPU scan produced for this window (integration/gpu_scan_popdel_call.h) to processSegment()'s call buffer;
//     unifyCalls / writeRegenotypedCalls / the window-wise writer then run unchanged.
// The scan needs a whole contig per call, the reference genotypes while it loads: popdel_call() therefore runs twice --
// pass 1 collects the read pairs (its window loop genotypes nothing), then every contig batch is scanned on the GPU, pass 2
// injects the calls window by window and lets the reference merge and write them. Slow by construction (the profiles are
// read twice); it is a proof, not a product.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <seqan/arg_parse.h>
#include <seqan/bam_io.h>

#include "parse_popdel.h"
#include "utils_popdel.h"
#include "popdel_profile/profile_parameter_parsing_popdel.h"
#include "popdel_profile/parameter_estimation_popdel.h"
#include "popdel_profile/bam_window_iterator_popdel.h"
#include "popdel_call/parameter_parsing_popdel_call.h"
#include "popdel_call/profile_structure_popdel_call.h"
#include "popdel_call/genotype_deletion_popdel_call.h"

#include "gpu_scan_popdel_call.h"

namespace gpubind {

struct ContigBatch {
    unsigned anchor;
    std::vector<std::vector<uint32_t> > pos;      // per read group
    std::vector<std::vector<int32_t> > dev;
    std::vector<pd_call> calls;                   // window calls of the GPU scan
    std::vector<uint32_t> rows;                   // n_calls x nSamples x 13
};
int pass = 1;
unsigned nRg = 0, nSamples = 0;
std::vector<ContigBatch> batches;
size_t batchNo = 0, cursor = 0;                   // pass 2: current contig batch (counted by resetTo) and next call to inject
PopDelCallParameters * params = NULL;
String<unsigned> maxLoadCopy;

}  // namespace gpubind

// the scan state the loader and the workflow see: records what enters it
struct GpuChromosomeProfile : public ChromosomeProfile
{
    GpuChromosomeProfile(unsigned numReadGroups, String<unsigned> & maximumLoad, unsigned bufferSize = 100) :
        ChromosomeProfile(numReadGroups, maximumLoad, bufferSize) {}

    inline void resetTo(unsigned pos)
    {
        if (gpubind::pass == 1)
        {
            gpubind::ContigBatch b;
            b.anchor = pos; b.pos.resize(gpubind::nRg); b.dev.resize(gpubind::nRg);
            gpubind::batches.push_back(b);
        }
        else { ++gpubind::batchNo; gpubind::cursor = 0; }
        ChromosomeProfile::resetTo(pos);
    }
    inline void add(unsigned readGroup, unsigned startPos, unsigned endPos, int deviation)
    {
        if (gpubind::pass == 1)
        {
            gpubind::ContigBatch & b = gpubind::batches.back();
            b.pos[readGroup].push_back(startPos); b.dev[readGroup].push_back(deviation);
        }
        ChromosomeProfile::add(readGroup, startPos, endPos, deviation);
    }
};

#define ChromosomeProfile GpuChromosomeProfile
#include "popdel_call/load_profile_popdel_call.h"
#include "popdel_call/parameter_calculation_popdel_call.h"
#undef ChromosomeProfile
#include "popdel_call/vcfout_popdel_call.h"

inline void gpubind_loadAndCalculateParameters(PopDelCallParameters & p)
{
    loadAndCalculateParameters(p);
    if (!gpubind::params) gpubind::params = new PopDelCallParameters(p);     // (popdel_call()'s own object dies with the call)
    gpubind::nRg = length(p.readGroups); gpubind::nSamples = length(p.rgs);
    gpubind::maxLoadCopy = p.maxLoad;              // the ChromosomeProfile constructor moves params.maxLoad away
}

// the window loop no longer genotypes: pass 1 walks only, pass 2 appends the GPU's calls of the window
inline bool gpubind_genotype_deletion_window(String<Call> & calls, const ChromosomeProfile & profile, const TRGs &, PopDelCallParameters &)
{
    if (gpubind::pass == 1 || gpubind::batchNo == 0)
        return false;
    const gpubind::ContigBatch & b = gpubind::batches[gpubind::batchNo - 1];
    bool any = false;
    size_t & k = gpubind::cursor;                                  // calls come in window order, windows are walked in order
    while (k < b.calls.size() && b.calls[k].window_position + 1 < profile.currentPos) ++k;
    for (; k < b.calls.size() && b.calls[k].window_position + 1 == profile.currentPos; ++k, any = true)
        appendValue(calls, gpuScanToCall(b.calls[k], &b.rows[k * 13ull * gpubind::nSamples], gpubind::nSamples));
    return any;
}

#define ChromosomeProfile GpuChromosomeProfile
#define loadAndCalculateParameters gpubind_loadAndCalculateParameters
#define genotype_deletion_window gpubind_genotype_deletion_window
#include "workflow_popdel.h"
#undef genotype_deletion_window
#undef loadAndCalculateParameters
#undef ChromosomeProfile

int main(int argc, char const ** argv)
{
    argv[0] = "popdel call";
    gpubind::pass = 1;
    int res = popdel_call(argc, argv);
    if (res != 0 || !gpubind::params)
        return res;
    // ---- the scan: one pd_contig_* round per contig batch
    PopDelCallParameters & p = *gpubind::params;
    p.maxLoad = gpubind::maxLoadCopy;
    const char * dev = getenv("POPDEL_GPU");
    pd_ctx * ctx = gpuScanCreate(p, dev ? atoi(dev) : 0);
    for (size_t b = 0; b < gpubind::batches.size(); ++b)
    {
        gpubind::ContigBatch & cb = gpubind::batches[b];
        bool any = false;
        for (unsigned g = 0; g < gpubind::nRg; ++g) any = any || !cb.pos[g].empty();
        if (!any) continue;
        int rc = pd_contig_begin(ctx, cb.anchor);
        for (unsigned g = 0; g < gpubind::nRg && rc == 0; ++g)
            rc = pd_contig_push(ctx, g, cb.pos[g].size(), cb.pos[g].data(), cb.dev[g].data());
        pd_result r;
        if (rc == 0) rc = pd_contig_scan(ctx, 0, 0, &r);
        if (rc != 0) { fprintf(stderr, "scan library: %s\n", pd_last_error(ctx)); return 3; }
        cb.calls.assign(r.calls, r.calls + r.n_calls);
        cb.rows.assign(r.per_sample, r.per_sample + r.n_calls * 13ull * gpubind::nSamples);
        std::vector<std::vector<uint32_t> >().swap(cb.pos); std::vector<std::vector<int32_t> >().swap(cb.dev);
    }
    pd_destroy(ctx);
    // ---- pass 2: the reference loads again, merges the injected calls per segment and writes the VCF
    gpubind::pass = 2; gpubind::batchNo = 0; gpubind::cursor = 0;
    return popdel_call(argc, argv);
}
