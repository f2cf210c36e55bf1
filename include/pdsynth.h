/* pdsynth.h -- synthetic read-pair generator (test / benchmark infrastructure, NOT part of the scan library:
 * built into its own libpdsynth.so so that the reference arm of bench.py maps nothing from the product). */
#ifndef PDSYNTH_H_
#define PDSYNTH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Synthetic read pairs of ONE read group, generated on the host with a counter-based RNG: per 30-bp
 * bucket and haplotype Poisson(pairs_per_bp*30/2) read pairs, insert size round(N(mu, sigma^2)) clipped to
 * (2*read_length, 20000); planted deletions are applied per haplotype (genotype 0/1/2): pairs whose forward read lies
 * in the deleted segment do not exist, pairs spanning the breakpoint get isize += length. Output sorted by
 * (pos, isize) like a profile. Returns the number of read pairs written (<= capacity) or a negative value (-1 invalid argument, -4 capacity).
 * Thread-safe (no context); used by bench.py and the tests to build cohorts of BASELINE.json's sizes quickly. */
int64_t pd_synth_read_group(uint64_t seed, uint32_t rg_index, double mu, double sigma, uint32_t read_length,
                            double pairs_per_bp, uint32_t first_pos, uint32_t end_pos,
                            uint32_t n_dels, const uint32_t * del_start, const uint32_t * del_len,
                            const uint8_t * del_genotype, uint32_t * pos, int32_t * isize, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* PDSYNTH_H_ */
