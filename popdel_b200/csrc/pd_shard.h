// pd_shard.h -- sample sharding state of a context (internal; pd_shard.cu, pd_scan.cu)
#ifndef PD_SHARD_H_
#define PD_SHARD_H_

#include "pd_device.cuh"

#define PD_XR_PAIRS 16384u            /* pairs per EM launch the exchange slots hold (= largest EM chunk) */

struct PdGroup;
struct PdShard {
    uint32_t rank = 0, world = 1, n_global = 0, r_global = 0, sample_offset = 0, n_local_max = 0;
    uint32_t part_n[PD_MAX_WORLD] = {};  // samples per rank
    int mode = 0;                        // 1 = in-process group, 2 = NCCL + CUDA IPC (one process per GPU)
    // in-kernel exchange (pd_em_common.cuh)
    XrSlot * xr_mine = nullptr; XrSlot * xr_peer[PD_MAX_WORLD] = {}; uint32_t xr_pairs_cap = 0;
    uint32_t * d_ticket = nullptr, * d_err = nullptr;
    uint64_t epoch = 1;                  // EM launches so far (identical on every rank)
    uint32_t grid_cap = 0;
    void * nccl_comm = nullptr;
    PdGroup * group = nullptr;
    void * d_small = nullptr;            // 4 KB staging of tiny all-gathers
    void * d_send = nullptr, * d_recv = nullptr; size_t cap_send = 0, cap_recv = 0;
};

int pd_shard_allgather(pd_ctx * c, const void * send, void * recv, size_t bytes, cudaStream_t st);
int pd_shard_or_flags(pd_ctx * c, uint32_t * flags, uint32_t n, cudaStream_t st);
int pd_shard_window_total(pd_ctx * c, uint64_t * total);
int pd_shard_prelaunch(pd_ctx * c);

#endif
