// comm.cu — ocbnn_comm: the collective plumbing of the two samplers with an exchange step (SVGD all-gathers
// particles and gradients and all-reduces the select histograms; BBB all-reduces its (D,2) gradient).  NCCL is bound with
// dlopen/dlsym at first use: a process that never creates a communicator (HMC, SGLD, one-GPU runs, the CPU-side ABI tests)
// needs no libnccl, and a process that already carries one (PyTorch's) shares it instead of loading a second copy.
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <nccl.h>
#include "comm.h"

namespace ocbnn {

namespace {
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // the copy already mapped into the process first (same SONAME as PyTorch's bundled library), then the system one
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        n.lib = h;
        bool all = true;
        auto sym = [&](const char* name) { void* p = dlsym(h, name); all = all && p; return p; };
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
        n.Broadcast = reinterpret_cast<decltype(n.Broadcast)>(sym("ncclBroadcast"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
        n.ok = all;
    });
    return n;
}

int nccl_fail(ncclResult_t r, const char* what) {
    set_error("NCCL error %d (%s) in %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
    return OCBNN_ECUDA;
}
#define OCBNN_NCCL(call)                                         \
    do {                                                         \
        ncclResult_t _r = (call);                                \
        if (_r != ncclSuccess) return nccl_fail(_r, #call);      \
    } while (0)
}  // namespace

int comm_all_sum(Comm* c, void* buf, long long count, CommType type, cudaStream_t st) {
    if (!c || c->world == 1 || count == 0) return OCBNN_OK;
    const ncclDataType_t t = type == kCommF32 ? ncclFloat32 : type == kCommU32 ? ncclUint32 : ncclUint64;
    OCBNN_NCCL(nccl().AllReduce(buf, buf, (size_t)count, t, ncclSum, (ncclComm_t)c->nccl, st));
    return OCBNN_OK;
}

int comm_all_gather_rows(Comm* c, float* all, long long n, long long row_floats, cudaStream_t st) {
    if (!c || c->world == 1 || n == 0) return OCBNN_OK;
    Nccl& N = nccl();
    ncclComm_t comm = (ncclComm_t)c->nccl;
    if (n % c->world == 0) {
        const size_t cnt = (size_t)(n / c->world * row_floats);
        OCBNN_NCCL(N.AllGather(all + (size_t)c->rank * cnt, all, cnt, ncclFloat32, comm, st));
        return OCBNN_OK;
    }
    // ragged shares: one broadcast per owner, fused into a single NCCL group
    OCBNN_NCCL(N.GroupStart());
    for (int r = 0; r < c->world; ++r) {
        long long lo, cnt;
        shard_of(n, r, c->world, &lo, &cnt);
        if (cnt == 0) continue;
        float* part = all + lo * row_floats;
        ncclResult_t res = N.Broadcast(part, part, (size_t)(cnt * row_floats), ncclFloat32, r, comm, st);
        if (res != ncclSuccess) { N.GroupEnd(); return nccl_fail(res, "ncclBroadcast"); }
    }
    OCBNN_NCCL(N.GroupEnd());
    return OCBNN_OK;
}

int side_stream(SideStream** out) {
    static SideStream per_dev[64];
    static std::mutex mu;
    int dev = 0;
    OCBNN_CUDA(cudaGetDevice(&dev));
    OCBNN_REQUIRE(dev >= 0 && dev < 64, OCBNN_EINVAL, "side_stream: device index %d", dev);
    std::lock_guard<std::mutex> lock(mu);
    SideStream& s = per_dev[dev];
    if (!s.stream) {
        OCBNN_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        OCBNN_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        OCBNN_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    *out = &s;
    return OCBNN_OK;
}

}  // namespace ocbnn

using namespace ocbnn;

static_assert(OCBNN_COMM_ID_BYTES == sizeof(ncclUniqueId), "ocbnn_comm id size");

extern "C" OCBNN_API int ocbnn_comm_unique_id(uint8_t* id) {
    OCBNN_REQUIRE(id, OCBNN_EINVAL, "ocbnn_comm_unique_id: NULL argument");
    OCBNN_REQUIRE(nccl().ok, OCBNN_EUNSUPPORTED, "libnccl.so.2 could not be loaded (%s)", dlerror() ? "dlopen failed" : "missing symbols");
    ncclUniqueId u;
    OCBNN_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_comm_create(const uint8_t* id, int32_t rank, int32_t world, ocbnn_comm** out) {
    OCBNN_REQUIRE(id && out && world >= 1 && rank >= 0 && rank < world, OCBNN_EINVAL, "ocbnn_comm_create: bad argument");
    OCBNN_REQUIRE(nccl().ok, OCBNN_EUNSUPPORTED, "libnccl.so.2 could not be loaded");
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    OCBNN_NCCL(nccl().CommInitRank(&comm, world, u, rank));
    Comm* c = new Comm;
    c->nccl = comm; c->rank = rank; c->world = world; c->owned = true;
    *out = reinterpret_cast<ocbnn_comm*>(c);
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_comm_wrap(void* nccl_comm, int32_t rank, int32_t world, ocbnn_comm** out) {
    OCBNN_REQUIRE(nccl_comm && out && world >= 1 && rank >= 0 && rank < world, OCBNN_EINVAL, "ocbnn_comm_wrap: bad argument");
    OCBNN_REQUIRE(nccl().ok, OCBNN_EUNSUPPORTED, "libnccl.so.2 could not be loaded");
    Comm* c = new Comm;
    c->nccl = nccl_comm; c->rank = rank; c->world = world; c->owned = false;
    *out = reinterpret_cast<ocbnn_comm*>(c);
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_comm_destroy(ocbnn_comm* h) {
    if (!h) return OCBNN_OK;
    Comm* c = reinterpret_cast<Comm*>(h);
    if (c->owned && c->nccl && nccl().ok) nccl().CommDestroy((ncclComm_t)c->nccl);
    delete c;
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_comm_shard(int64_t n, int32_t rank, int32_t world, int64_t* row_begin, int64_t* n_rows) {
    OCBNN_REQUIRE(row_begin && n_rows && n >= 0 && world >= 1 && rank >= 0 && rank < world, OCBNN_EINVAL, "ocbnn_comm_shard: bad argument");
    long long lo, cnt;
    shard_of(n, rank, world, &lo, &cnt);
    *row_begin = lo;
    *n_rows = cnt;
    return OCBNN_OK;
}
