// Thin runtime binding to NCCL (dlopen of libnccl.so.2, so that single-GPU use has no NCCL dependency).
// Communicators are cached per process under their unique id (a state created with an id seen before reuses it);
// collectives are enqueued on the state's stream.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace cmfb200 {

class NcclLink {
public:
    NcclLink() {}
    ~NcclLink();
    static int unique_id(void *out128);                        // ncclGetUniqueId; out must hold 128 bytes
    int init(const void *id128, int rank, int world);          // ncclCommInitRank, or the cached communicator of that id
    // every rank contributes `bytes_per_rank` bytes located at base + rank*bytes_per_rank
    int all_gather_inplace(void *base, size_t bytes_per_rank, cudaStream_t stream);
    int all_reduce_sum(void *buf, size_t count, bool is_double, cudaStream_t stream);
    int group_begin();   // collectives issued until group_end() go out as one NCCL launch
    int group_end();
    int rank = 0, world = 1;
private:
    void *comm = nullptr;
};

}  // namespace cmfb200
