// pd_synth.cpp -- host generator of synthetic read pairs, counter-based so that any read group of any
// cohort size can be generated independently and in parallel. Not part of the scan library (libpdsynth.so).
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/pdsynth.h"

#define PD_WIN 30u

namespace {

inline uint64_t mix64(uint64_t x)      // splitmix64 finaliser
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct Rng {
    uint64_t key, ctr;
    uint64_t next() { return mix64(key + (ctr++) * 0xD1342543DE82EF95ull); }
    double uniform() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
};

}  // namespace

extern "C" int64_t pd_synth_read_group(uint64_t seed, uint32_t rg_index, double mu, double sigma, uint32_t read_length,
                                       double pairs_per_bp, uint32_t first_pos, uint32_t end_pos,
                                       uint32_t n_dels, const uint32_t * del_start, const uint32_t * del_len,
                                       const uint8_t * del_genotype, uint32_t * pos, int32_t * isize, uint64_t capacity)
{
    if (!pos || !isize || end_pos <= first_pos || sigma <= 0 || (n_dels && (!del_start || !del_len || !del_genotype)))
        return -1;
    const double lambda = pairs_per_bp * PD_WIN / 2.0;          // per bucket and haplotype
    // Poisson inverse-CDF table
    std::vector<double> cdf;
    { double p = std::exp(-lambda), c = p; cdf.push_back(c); for (int k = 1; k < 64; ++k) { p *= lambda / k; c += p; cdf.push_back(c); } }
    const int lo_clip = 2 * (int)read_length + 1, hi_clip = 19999;
    uint64_t n = 0;
    struct P { uint32_t pos; int32_t isz; };
    std::vector<P> bucket;
    const uint32_t b0 = first_pos / PD_WIN, b1 = (end_pos + PD_WIN - 1) / PD_WIN;
    for (uint32_t b = b0; b < b1; ++b) {
        bucket.clear();
        for (uint32_t hap = 0; hap < 2; ++hap) {
            Rng r{mix64(seed ^ mix64(((uint64_t)rg_index << 34) ^ ((uint64_t)b << 1) ^ hap)), 0};
            const double u = r.uniform();
            int k = 0;
            while (k < 63 && u > cdf[k]) ++k;
            for (int i = 0; i < k; ++i) {
                uint32_t p = b * PD_WIN + (uint32_t)(r.uniform() * PD_WIN);
                // Box-Muller
                const double u1 = r.uniform(), u2 = r.uniform();
                const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
                int isz = (int)std::lrint(mu + sigma * z);
                isz = std::min(std::max(isz, lo_clip), hi_clip);
                if (p < first_pos || p >= end_pos) continue;
                bool drop = false;
                const int64_t inner = isz - 2 * (int)read_length;   // forward-read end -> reverse-read start
                for (uint32_t d = 0; d < n_dels; ++d) {
                    if (del_genotype[d] <= hap) continue;
                    const int64_t s = del_start[d], e = (int64_t)del_start[d] + del_len[d];
                    if ((int64_t)p < s) {
                        const int64_t rs = (int64_t)p + inner;              // reverse-read start on the donor haplotype
                        if (rs >= s) isz += (int)del_len[d];                // pair spans the junction
                        else if (rs + (int64_t)read_length > s) { drop = true; break; }   // junction inside the reverse read
                    } else if ((int64_t)p - (int64_t)read_length + 1 < e) { drop = true; break; }   // forward read in the deleted segment
                }
                if (drop) continue;
                bucket.push_back(P{p, isz});
            }
        }
        std::sort(bucket.begin(), bucket.end(), [](const P & x, const P & y) { return x.pos != y.pos ? x.pos < y.pos : x.isz < y.isz; });
        if (n + bucket.size() > capacity) return -4;
        for (const P & q : bucket) { pos[n] = q.pos; isize[n] = q.isz; ++n; }
    }
    return (int64_t)n;
}
