// popdel_call/gpu_scan_popdel_call.h -- the reference-side binding of libpopdel_b200: what a maintainer of
// kehrlab/PopDel adds next to popdel_call/*.h. Needs the reference's headers (PopDelCallParameters, Histogram, Call, Dad)
// to be included first and include/popdel_b200.h on the include path. Compiled against the unmodified reference by
// integration/Makefile (integration/popdel_call_gpu.cpp) and run on the golden cohorts by tests/test_integration_gpu.py.
#ifndef GPU_SCAN_POPDEL_CALL_H_
#define GPU_SCAN_POPDEL_CALL_H_

#include <vector>

#include "popdel_b200.h"

// after loadAndCalculateParameters(params): one context per GPU (parameter_parsing_popdel_call.h:141-210)
inline pd_ctx * gpuScanCreate(const PopDelCallParameters & params, int device)
{
    pd_params p;
    p.iterations = params.iterations; p.min_len = params.minLen; p.min_lr = params.minimumLikelihoodRatio;
    p.min_sample_fraction = params.minSampleFraction; p.window_size = params.windowSize; p.window_buffer = params.windowBuffer;
    p.somatic = params.somatic; p.window_wise = params.windowWiseOutput;
    std::vector<pd_rg> rgs(length(params.histograms));
    for (unsigned s = 0; s < length(params.rgs); ++s)
        for (unsigned j = 0; j < length(params.rgs[s]); ++j)
        {
            const unsigned g = params.rgs[s][j];
            const Histogram & h = params.histograms[g];          // already processed by processHistogram()
            pd_rg r;
            r.sample = s; r.median = h.median; r.read_length = h.readLength; r.stddev = h.stddev; r.offset = h.offset;
            r.len = (uint32_t)length(h.values); r.values = &h.values[0]; r.min_prob = h.min_prob;
            r.lower_quantile_dist = h.lowerQuantileDist; r.upper_quantile_dist = h.upperQuantileDist;
            r.max_load = params.maxLoad[g] == 0 ? 0xFFFFFFFFu : params.maxLoad[g];
            r.min_init_del_len = params.minInitDelLengths[g];
            rgs[g] = r;
        }
    pd_ctx * ctx = pd_create(&p, (uint32_t)length(params.rgs), (uint32_t)rgs.size(), rgs.data(), device);
    if (!ctx)
        SEQAN_THROW(IOError(pd_create_error()));
    return ctx;
}

// one window call of pd_contig_scan -> the reference's Call (utils_popdel.h:57-124)
inline Call gpuScanToCall(const pd_call & c, const uint32_t * ps, unsigned nSamples)
{
    Call call(c.initial_length, c.iterations, c.deletion_length, c.lr, c.frequency,
              c.window_position, c.position, c.end_position);
    call.filter = (unsigned char)c.filter;
    resize(call.gtLikelihoods, nSamples); resize(call.lads, nSamples);
    resize(call.dads, nSamples); resize(call.firstLast, nSamples);
    for (unsigned s = 0; s < nSamples; ++s, ps += 13)
    {
        call.gtLikelihoods[s] = Triple<unsigned>(ps[0], ps[1], ps[2]);
        call.lads[s] = Triple<unsigned>(ps[3], ps[4], ps[5]);
        call.dads[s] = Dad(ps[6], ps[7], ps[8], ps[9], ps[10]);
        call.firstLast[s] = Pair<unsigned>(ps[11], ps[12]);
    }
    return call;
}

#endif
