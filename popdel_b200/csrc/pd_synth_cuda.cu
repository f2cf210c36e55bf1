// pd_synth_cuda.cu -- device-side version of pd_synth.cpp (the on-device counter-based generator for the
// cohort sizes that do not fit the host): the SAME stream of read pairs as pd_synth_read_group, bucket by bucket, written
// straight into device arrays that pd_contig_push_device packs where they are. Not part of the scan library
// (libpdsynth_cuda.so); test / benchmark infrastructure.
//
// One thread = one 30-bp bucket of one read group (both haplotypes). Pass 1 counts the read pairs of every chunk of 256
// buckets, an exclusive scan over (read group, chunk) gives the output offsets, pass 2 regenerates the buckets (the RNG is
// counter-based, so nothing is stored in between), sorts each bucket by (pos, isize) and writes position and deviation.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/pdsynth.h"

#define PD_WIN 30u
#define SYN_CHUNK 256
#define SYN_MAX_BUCKET 32            // read pairs per bucket the device generator handles (host: 126)

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct Rng {
    uint64_t key, ctr;
    __device__ uint64_t next() { return mix64(key + (ctr++) * 0xD1342543DE82EF95ull); }
    __device__ double uniform() { return __dmul_rn(__dadd_rn((double)(next() >> 11), 0.5), 1.0 / 9007199254740992.0); }
};

struct DevRg { double mu, sigma, lambda; uint32_t rg_index, sample, read_length; int32_t median; uint32_t cdf_off, pad; };

struct GenArgs {
    const DevRg * rgs; const double * cdf;           // per read group a 64-entry Poisson CDF (built on the host like pd_synth.cpp does)
    const uint32_t * del_start, * del_len; const uint8_t * del_gt;      // [n_dels], [n_dels][n_samples]
    uint32_t n_dels, n_samples, max_del_len;
    uint64_t seed;
    uint32_t first_pos, end_pos, b0, n_buckets, n_chunks;
    uint32_t * overflow;
};

// the read pairs of bucket b of read group g, unsorted; returns their number (or SYN_MAX_BUCKET + 1 on overflow).
// COUNT_ONLY: nothing is stored, and a bucket that no planted deletion can touch and that lies inside the range is just
// its two Poisson draws.
template <bool COUNT_ONLY>
__device__ __forceinline__ int gen_bucket(const GenArgs & a, const DevRg & rg, uint32_t b, uint32_t * P, int32_t * Z)
{
    int n = 0;
    const double * cdf = a.cdf + rg.cdf_off;
    const int lo_clip = 2 * (int)rg.read_length + 1, hi_clip = 19999;
    // deletions are sorted by start: only those with start in (p - max_len - read_length, p + 20000 + read_length) can touch a
    // read pair at p; one search per bucket (p in [30 b, 30 b + 29])
    const long long p_lo = (long long)b * PD_WIN, p_hi = p_lo + PD_WIN - 1;
    const long long lo_s = p_lo - (long long)a.max_del_len - (long long)rg.read_length - 1, hi_s = p_hi + 20000 + (long long)rg.read_length;
    uint32_t d0 = 0, d1 = a.n_dels;
    while (d0 < d1) { const uint32_t m = (d0 + d1) >> 1; if ((long long)a.del_start[m] < lo_s) d0 = m + 1; else d1 = m; }
    d1 = d0;
    while (d1 < a.n_dels && (long long)a.del_start[d1] <= hi_s) ++d1;
    const bool plain = d0 == d1 && p_lo >= (long long)a.first_pos && p_hi < (long long)a.end_pos;
    for (uint32_t hap = 0; hap < 2; ++hap) {
        Rng r{mix64(a.seed ^ mix64(((uint64_t)rg.rg_index << 34) ^ ((uint64_t)b << 1) ^ hap)), 0};
        const double u = r.uniform();
        int k = 0;
        while (k < 63 && u > cdf[k]) ++k;
        if (COUNT_ONLY && plain) { n += k; continue; }
        for (int i = 0; i < k; ++i) {
            const uint32_t p = b * PD_WIN + (uint32_t)(r.uniform() * PD_WIN);
            const double u1 = r.uniform(), u2 = r.uniform();
            if (COUNT_ONLY && d0 == d1) { if (p >= a.first_pos && p < a.end_pos) ++n; continue; }
            const double z = __dmul_rn(sqrt(__dmul_rn(-2.0, log(u1))), cos(__dmul_rn(6.283185307179586, u2)));
            int isz = (int)lrint(__dadd_rn(rg.mu, __dmul_rn(rg.sigma, z)));      // (no fused multiply-add: the host generator has none)
            isz = min(max(isz, lo_clip), hi_clip);
            if (p < a.first_pos || p >= a.end_pos) continue;
            bool drop = false;
            const long long inner = isz - 2 * (int)rg.read_length;
            for (uint32_t d = d0; d < d1; ++d) {
                if (a.del_gt[(size_t)d * a.n_samples + rg.sample] <= hap) continue;
                const long long s = a.del_start[d], e = s + a.del_len[d];
                if ((long long)p < s) {
                    const long long rs = (long long)p + inner;
                    if (rs >= s) isz += (int)a.del_len[d];
                    else if (rs + (long long)rg.read_length > s) { drop = true; break; }
                } else if ((long long)p - (long long)rg.read_length + 1 < e) { drop = true; break; }
            }
            if (drop) continue;
            if (n >= SYN_MAX_BUCKET) return SYN_MAX_BUCKET + 1;
            if (!COUNT_ONLY) { P[n] = p; Z[n] = isz; }
            ++n;
        }
    }
    return n;
}

__device__ __forceinline__ uint32_t block_sum_and_scan(uint32_t v, uint32_t & excl)     // 256 threads
{
    __shared__ uint32_t ws[8];
    __shared__ uint32_t tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < 8; ++i) { const uint32_t t = ws[i]; ws[i] = s; s += t; } tot = s; }
    __syncthreads();
    excl = ws[w] + x - v;
    return tot;
}

__global__ void __launch_bounds__(SYN_CHUNK) k_syn_count(GenArgs a, unsigned long long * __restrict__ chunk_sum)
{
    const uint32_t g = blockIdx.y, b = a.b0 + blockIdx.x * SYN_CHUNK + threadIdx.x;
    int n = 0;
    if (blockIdx.x * SYN_CHUNK + threadIdx.x < a.n_buckets) n = gen_bucket<true>(a, a.rgs[g], b, nullptr, nullptr);
    if (n > SYN_MAX_BUCKET) { atomicExch(a.overflow, 1u); n = 0; }
    uint32_t excl;
    const uint32_t tot = block_sum_and_scan((uint32_t)n, excl);
    if (threadIdx.x == 0) chunk_sum[(size_t)g * a.n_chunks + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SYN_CHUNK) k_syn_fill(GenArgs a, const unsigned long long * __restrict__ chunk_off, uint32_t * __restrict__ pos,
                                                        int32_t * __restrict__ dev)
{
    const uint32_t g = blockIdx.y, b = a.b0 + blockIdx.x * SYN_CHUNK + threadIdx.x;
    const DevRg rg = a.rgs[g];
    uint32_t P[SYN_MAX_BUCKET]; int32_t Z[SYN_MAX_BUCKET];
    int n = 0;
    if (blockIdx.x * SYN_CHUNK + threadIdx.x < a.n_buckets) n = gen_bucket<false>(a, rg, b, P, Z);
    if (n > SYN_MAX_BUCKET) n = 0;
    uint32_t excl;
    block_sum_and_scan((uint32_t)n, excl);
    for (int i = 1; i < n; ++i) {                                    // insertion sort by (pos, isize)
        const uint32_t p = P[i]; const int32_t z = Z[i];
        int j = i - 1;
        while (j >= 0 && (P[j] > p || (P[j] == p && Z[j] > z))) { P[j + 1] = P[j]; Z[j + 1] = Z[j]; --j; }
        P[j + 1] = p; Z[j + 1] = z;
    }
    const unsigned long long o = chunk_off[(size_t)g * a.n_chunks + blockIdx.x] + excl;
    for (int i = 0; i < n; ++i) { pos[o + i] = P[i]; dev[o + i] = Z[i] - rg.median; }
}

__global__ void k_syn_rg_start(const unsigned long long * __restrict__ chunk_off, uint32_t n_rg, uint32_t n_chunks, unsigned long long total,
                               unsigned long long * __restrict__ out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_rg) out[g] = chunk_off[(size_t)g * n_chunks];
    if (g == n_rg) out[g] = total;
}

template <typename T>
bool grow(T *& p, size_t & cap, size_t need)
{
    if (need <= cap && p) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = need + need / 8 + 256;
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) return false;
    cap = want;
    return true;
}

}  // namespace

struct pdsynth_dev {
    int device = 0;
    cudaStream_t st = nullptr;
    uint32_t * d_pos = nullptr; size_t cap_pos = 0;
    int32_t * d_dev = nullptr; size_t cap_dev = 0;
    unsigned long long * d_sum = nullptr; size_t cap_sum = 0;
    unsigned long long * d_off = nullptr; size_t cap_off = 0;
    char * d_tmp = nullptr; size_t cap_tmp = 0;
    char * d_in = nullptr; size_t cap_in = 0;
};

extern "C" pdsynth_dev * pdsynth_dev_create(int device)
{
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    pdsynth_dev * h = new pdsynth_dev();
    h->device = device;
    if (cudaStreamCreate(&h->st) != cudaSuccess) { delete h; return nullptr; }
    return h;
}

extern "C" void pdsynth_dev_destroy(pdsynth_dev * h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_pos); cudaFree(h->d_dev); cudaFree(h->d_sum); cudaFree(h->d_off); cudaFree(h->d_tmp); cudaFree(h->d_in);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int64_t pdsynth_dev_generate(pdsynth_dev * h, uint64_t seed, uint32_t n_rg, const pdsynth_rg * rgs, uint32_t first_pos, uint32_t end_pos,
                                        uint32_t n_dels, const uint32_t * del_start, const uint32_t * del_len, const uint8_t * del_genotype,
                                        uint32_t n_samples, const uint32_t ** d_pos, const int32_t ** d_dev, uint64_t * rg_start)
{
    if (!h || !n_rg || !rgs || end_pos <= first_pos || !d_pos || !d_dev || !rg_start || (n_dels && (!del_start || !del_len || !del_genotype)))
        return -1;
    for (uint32_t d = 1; d < n_dels; ++d) if (del_start[d] < del_start[d - 1]) return -2;       // deletions must be sorted by start
    if (cudaSetDevice(h->device) != cudaSuccess) return -3;
    // host-side tables: the Poisson CDF exactly as pd_synth.cpp builds it (one per distinct density)
    std::vector<DevRg> hr(n_rg);
    std::vector<double> cdf;
    std::vector<double> dens;
    for (uint32_t g = 0; g < n_rg; ++g) {
        if (rgs[g].sigma <= 0 || rgs[g].sample >= std::max(n_samples, 1u)) return -1;
        size_t k = 0;
        while (k < dens.size() && dens[k] != rgs[g].pairs_per_bp) ++k;
        if (k == dens.size()) {
            dens.push_back(rgs[g].pairs_per_bp);
            const double lambda = rgs[g].pairs_per_bp * PD_WIN / 2.0;
            double p = std::exp(-lambda), c = p;
            cdf.push_back(c);
            for (int i = 1; i < 64; ++i) { p *= lambda / i; c += p; cdf.push_back(c); }
        }
        DevRg & r = hr[g];
        r.mu = rgs[g].mu; r.sigma = rgs[g].sigma; r.lambda = 0; r.rg_index = rgs[g].rg_index; r.sample = rgs[g].sample;
        r.read_length = rgs[g].read_length; r.median = (int32_t)rgs[g].median; r.cdf_off = (uint32_t)(k * 64); r.pad = 0;
    }
    uint32_t max_len = 0;
    for (uint32_t d = 0; d < n_dels; ++d) max_len = std::max(max_len, del_len[d]);
    GenArgs a;
    a.b0 = first_pos / PD_WIN;
    a.n_buckets = (end_pos + PD_WIN - 1) / PD_WIN - a.b0;
    a.n_chunks = (a.n_buckets + SYN_CHUNK - 1) / SYN_CHUNK;
    a.n_dels = n_dels; a.n_samples = std::max(n_samples, 1u); a.max_del_len = max_len; a.seed = seed; a.first_pos = first_pos; a.end_pos = end_pos;
    // inputs in one device buffer: rgs | cdf | del_start | del_len | del_gt | overflow flag
    const size_t o_cdf = (hr.size() * sizeof(DevRg) + 15) & ~(size_t)15, o_ds = o_cdf + cdf.size() * 8, o_dl = o_ds + (size_t)n_dels * 4,
                 o_gt = o_dl + (size_t)n_dels * 4, o_fl = (o_gt + (size_t)n_dels * a.n_samples + 15) & ~(size_t)15, in_bytes = o_fl + 16;
    if (!grow(h->d_in, h->cap_in, in_bytes)) return -3;
    cudaStream_t st = h->st;
    cudaMemcpyAsync(h->d_in, hr.data(), hr.size() * sizeof(DevRg), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(h->d_in + o_cdf, cdf.data(), cdf.size() * 8, cudaMemcpyHostToDevice, st);
    if (n_dels) {
        cudaMemcpyAsync(h->d_in + o_ds, del_start, (size_t)n_dels * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(h->d_in + o_dl, del_len, (size_t)n_dels * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(h->d_in + o_gt, del_genotype, (size_t)n_dels * a.n_samples, cudaMemcpyHostToDevice, st);
    }
    cudaMemsetAsync(h->d_in + o_fl, 0, 16, st);
    a.rgs = reinterpret_cast<const DevRg *>(h->d_in); a.cdf = reinterpret_cast<const double *>(h->d_in + o_cdf);
    a.del_start = reinterpret_cast<const uint32_t *>(h->d_in + o_ds); a.del_len = reinterpret_cast<const uint32_t *>(h->d_in + o_dl);
    a.del_gt = reinterpret_cast<const uint8_t *>(h->d_in + o_gt); a.overflow = reinterpret_cast<uint32_t *>(h->d_in + o_fl);
    const size_t n_cs = (size_t)n_rg * a.n_chunks;
    if (!grow(h->d_sum, h->cap_sum, n_cs + 1) || !grow(h->d_off, h->cap_off, n_cs + n_rg + 2)) return -3;
    k_syn_count<<<dim3(a.n_chunks, n_rg), SYN_CHUNK, 0, st>>>(a, h->d_sum);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_sum, h->d_off, (int)n_cs, st);
    if (!grow(h->d_tmp, h->cap_tmp, tmp_bytes)) return -3;
    cub::DeviceScan::ExclusiveSum(h->d_tmp, tmp_bytes, h->d_sum, h->d_off, (int)n_cs, st);
    unsigned long long last_off = 0, last_sum = 0; uint32_t overflow = 0;
    cudaMemcpyAsync(&last_off, h->d_off + n_cs - 1, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last_sum, h->d_sum + n_cs - 1, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&overflow, a.overflow, 4, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return -3;
    if (overflow) return -5;                                         // more than SYN_MAX_BUCKET read pairs in a 30-bp bucket
    const unsigned long long total = last_off + last_sum;
    if (!grow(h->d_pos, h->cap_pos, (size_t)total + 1) || !grow(h->d_dev, h->cap_dev, (size_t)total + 1)) return -3;
    k_syn_fill<<<dim3(a.n_chunks, n_rg), SYN_CHUNK, 0, st>>>(a, h->d_off, h->d_pos, h->d_dev);
    unsigned long long * d_rs = h->d_off + n_cs;
    k_syn_rg_start<<<(n_rg + 256) / 256, 256, 0, st>>>(h->d_off, n_rg, a.n_chunks, total, d_rs);
    cudaMemcpyAsync(rg_start, d_rs, ((size_t)n_rg + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return -3;
    *d_pos = h->d_pos; *d_dev = h->d_dev;
    return (int64_t)total;
}

extern "C" int pdsynth_dev_copy_to_host(pdsynth_dev * h, void * dst, const void * d_src, uint64_t bytes)
{
    if (!h || (bytes && (!dst || !d_src))) return -1;
    if (cudaSetDevice(h->device) != cudaSuccess) return -3;
    return cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
}
