// comm.h — the communicator behind ocbnn_comm (comm.cu): NCCL bound at run time, plus the fork/join side stream the one-call
// sampler steps (ocbnn_svgd_step, ocbnn_bbb_step) use to overlap K1 with the select chain.
#pragma once
#include "common.cuh"

namespace ocbnn {

struct Comm {
    void* nccl = nullptr;        // ncclComm_t
    int rank = 0, world = 1;
    bool owned = false;          // created by ocbnn_comm_create (destroyed with the handle) or adopted by ocbnn_comm_wrap
};

enum CommType { kCommF32 = 0, kCommU32 = 1, kCommU64 = 2 };

inline int comm_rank(const Comm* c) { return c ? c->rank : 0; }
inline int comm_world(const Comm* c) { return c ? c->world : 1; }

// contiguous shares: the first n % world ranks hold one more (ocbnn_comm_shard)
inline void shard_of(long long n, int rank, int world, long long* lo, long long* cnt) {
    const long long base = n / world, rem = n % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *cnt = base + (rank < rem ? 1 : 0);
}

// in-place sum over the ranks (no-op for one rank)
int comm_all_sum(Comm* c, void* buf, long long count, CommType type, cudaStream_t st);
// in-place all-gather of the row shares of all[n x row_floats] laid out by shard_of (no-op for one rank)
int comm_all_gather_rows(Comm* c, float* all, long long n, long long row_floats, cudaStream_t st);

// per-device side stream with a fork and a join event (created on first use, never during a capture of the caller's stream:
// the first step of a run is always eager)
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
int side_stream(SideStream** out);

}  // namespace ocbnn
