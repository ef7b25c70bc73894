// NCCL all-reduce for row-sharded training without a host-language hop.
//
// The reference is single-process (no collectives).  The row-sharded B200 path sums 2T p-vectors, one
// p x p Gram and one k-vector per image across ranks; this file owns an NCCL communicator for exactly
// that, so that the 41 small all-reduces of an image are ncclAllReduce calls enqueued on the training stream by the
// library itself (round 1 went through a ctypes -> Python -> torch.distributed callback per reduction).
//
// libnccl is resolved at run time (dlopen): the host process normally has it loaded already (PyTorch links
// libnccl.so.2); otherwise NLE_B200_NCCL_LIB names it.  libnle_b200.so itself therefore links no NCCL, and a
// single-GPU user never needs it.  The host language only ferries the 128-byte unique id from rank 0 to the other
// ranks (bench.py / sharding.py: one torch.distributed broadcast at start-up).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../include/nle_b200.h"
#include "common.cuh"

namespace nle {
namespace {

struct NcclId { char internal[128]; };          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef struct ncclComm* ncclComm_t;
constexpr int kNcclFloat64 = 8;                 // ncclDataType_t::ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;                     // ncclRedOp_t::ncclSum

struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        if (const char* e = getenv("NLE_B200_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // already in the process (PyTorch)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.why = "libnccl.so.2 not found (set NLE_B200_NCCL_LIB)"; return; }
        auto sym = [&](const char* name) { return dlsym(h, name); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
        if (!api.ok) api.why = "libnccl is missing one of ncclGetUniqueId/CommInitRank/AllReduce/CommDestroy/GetErrorString";
    });
    return api;
}

int fail(const std::string& msg) {
    set_error(msg);
    return NLE_B200_ERR_CUDA;
}

}  // namespace
}  // namespace nle

using namespace nle;

struct nle_b200_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

extern "C" {

int nle_b200_comm_unique_id(unsigned char id[128]) {
    NcclApi& api = nccl_api();
    if (!api.ok) return fail("NCCL unavailable: " + api.why);
    if (!id) { set_error("null pointer"); return NLE_B200_ERR_INVALID; }
    NcclId u;
    const int rc = api.GetUniqueId(&u);
    if (rc != 0) return fail(std::string("ncclGetUniqueId: ") + api.GetErrorString(rc));
    std::memcpy(id, u.internal, 128);
    return NLE_B200_OK;
}

int nle_b200_comm_create(const unsigned char id[128], int rank, int nranks, nle_b200_comm** out) {
    NcclApi& api = nccl_api();
    if (!api.ok) return fail("NCCL unavailable: " + api.why);
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad communicator arguments"); return NLE_B200_ERR_INVALID; }
    *out = nullptr;
    NcclId u;
    std::memcpy(u.internal, id, 128);
    ncclComm_t c = nullptr;
    const int rc = api.CommInitRank(&c, nranks, u, rank);      // collective: every rank calls it, on its own device
    if (rc != 0) return fail(std::string("ncclCommInitRank: ") + api.GetErrorString(rc));
    auto* h = new nle_b200_comm;
    h->comm = c; h->rank = rank; h->nranks = nranks;
    *out = h;
    return NLE_B200_OK;
}

/* An nle_b200_allreduce_fn: pass it as `allreduce` with `user` = the nle_b200_comm*.  In-place FP64 sum on `cuda_stream`. */
int nle_b200_comm_allreduce(void* dev_buf, size_t count, void* cuda_stream, void* user) {
    auto* h = static_cast<nle_b200_comm*>(user);
    if (!h || !h->comm) return 1;
    if (h->nranks == 1 || count == 0) return 0;
    const int rc = nccl_api().AllReduce(dev_buf, dev_buf, count, kNcclFloat64, kNcclSum, h->comm, static_cast<cudaStream_t>(cuda_stream));
    if (rc != 0) { set_error(std::string("ncclAllReduce: ") + nccl_api().GetErrorString(rc)); return rc; }
    return 0;
}

void nle_b200_comm_destroy(nle_b200_comm* h) {
    if (!h) return;
    if (h->comm) nccl_api().CommDestroy(h->comm);
    delete h;
}

}  // extern "C"
