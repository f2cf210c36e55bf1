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

/* The same generator on the GPU (libpdsynth_cuda.so, SURVEY.md 8d: cohorts whose read pairs fit neither the host nor a
 * PCIe budget): all read groups of a cohort for [first_pos, end_pos) at once, written to device arrays owned by the handle
 * (valid until the next call): *d_pos / *d_dev hold the read pairs of read group g at [rg_start[g], rg_start[g+1])
 * (rg_start: host array of n_rg + 1), positions sorted by (pos, isize), dev = isize - median -- the arrays
 * pd_contig_push_device takes. Same stream of read pairs as pd_synth_read_group(seed, rgs[g].rg_index, ...) with the
 * genotype column of rgs[g].sample. del_start must be ascending; del_genotype is [n_dels][n_samples]. Returns the number of
 * read pairs, or <0: -1 argument, -2 deletions not sorted, -3 CUDA error / out of memory, -5 more than 32 read pairs in
 * one 30-bp bucket. */
typedef struct pdsynth_dev pdsynth_dev;
typedef struct { double mu, sigma, pairs_per_bp; uint32_t rg_index, sample, read_length, median; } pdsynth_rg;
pdsynth_dev * pdsynth_dev_create(int device);
void pdsynth_dev_destroy(pdsynth_dev * h);
int64_t pdsynth_dev_generate(pdsynth_dev * h, uint64_t seed, uint32_t n_rg, const pdsynth_rg * rgs, uint32_t first_pos, uint32_t end_pos,
                             uint32_t n_dels, const uint32_t * del_start, const uint32_t * del_len, const uint8_t * del_genotype,
                             uint32_t n_samples, const uint32_t ** d_pos, const int32_t ** d_dev, uint64_t * rg_start);
/* copies generated read pairs back to the host (tests) */
int pdsynth_dev_copy_to_host(pdsynth_dev * h, void * dst, const void * d_src, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PDSYNTH_H_ */
