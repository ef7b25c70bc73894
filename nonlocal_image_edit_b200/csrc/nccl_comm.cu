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
//
// Small messages do not go through NCCL at all.  The p-vectors of the Sinkhorn passes (2T per image, 12.8 KB each) and
// the k-vector of an apply are latency-sized, and each one sits between two dependent kernels: a ncclAllReduce costs
// 45-65 us there (profiles/r2r, r2u: sinkhorn_passes 8.9 -> 10.7 ms on 2 GPUs, -> 11.5 ms on 8).  peer_allreduce_kernel
// does the exchange itself over NVLink peer memory: every rank owns an inbox (cudaMalloc, exported with CUDA IPC, the
// handles all-gathered once at communicator creation), a reduction is ONE launch in which thread i stores element i as a
// flagged 16-byte cell {lo, tag, hi, tag} straight into the inbox of every rank (st.relaxed.sys over NVLink; each 8-byte
// half carries its tag, nothing beyond 8-byte atomicity is assumed -- the LL protocol, as in tridiag_cluster_kernel),
// polls its own inbox until the nranks cells of element i carry the tag of this call and adds them up in rank order, so
// every rank computes bit-identical sums.  No barrier, no second pass, no host involvement; inboxes are double-buffered by
// the parity of the call number (a rank can be at most one call ahead of a peer: it needs that peer's cells to finish).
// Messages above kPeerMax doubles (the p x p Gram) and communicators whose peers cannot map each other's memory
// (no P2P, IPC refused) use ncclAllReduce.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

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
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;     // optional (peer set-up)
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
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
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

// ---- all-reduce of small FP64 vectors over NVLink peer memory ---------------------------------------------------
constexpr int kPeerRanks = 16;                  // inbox slots (ranks of one NVSwitch domain)
constexpr int kPeerMax = 8192;                  // doubles per message; inbox = 2 * kPeerRanks * kPeerMax cells = 4 MB
constexpr int kNcclChar = 0;                    // ncclDataType_t::ncclInt8 / ncclChar

struct PeerInboxes { uint4* p[kPeerRanks]; };

__device__ __forceinline__ void cell_store_sys(uint4* cell, double x, unsigned tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(t | (bits & 0xffffffffull)), "l"(t | (bits >> 32))
                 : "memory");
}
__device__ __forceinline__ void cell_load_sys(const uint4* cell, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(cell) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// buf[i] <- sum over ranks of buf[i], i < count <= kPeerMax.  `tag` = number of this call on the communicator (>= 1,
// the same on every rank), slot (tag & 1, source rank) of every inbox receives the source's vector.
// A peer that never arrives must not hang the device for ever: after kPeerTimeoutNs without progress the launch traps
// (every later CUDA call of the process then reports the error).
constexpr unsigned long long kPeerTimeoutNs = 120ull * 1000000000ull;
template <int R>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(double* __restrict__ buf, int count, PeerInboxes in, int rank, int nranks, unsigned tag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const size_t par = (size_t)(tag & 1u) * kPeerRanks;
    const double x = buf[i];
    const size_t mine = (par + rank) * kPeerMax + i;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r < nranks) cell_store_sys(in.p[r] + mine, x, tag);
    const uint4* box = in.p[rank];
    unsigned long long a[R], b[R];
    unsigned pend = (nranks >= 32) ? 0xffffffffu : ((1u << nranks) - 1u);
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (pend) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (pend & (1u << r)) cell_load_sys(box + (par + r) * kPeerMax + i, a[r], b[r]);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if ((pend & (1u << r)) && (unsigned)(a[r] >> 32) == tag && (unsigned)(b[r] >> 32) == tag) pend &= ~(1u << r);
        if (pend && (++spins & 0x3ffu) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kPeerTimeoutNs) __trap();
        }
    }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r < nranks) {
            const double v = __longlong_as_double((long long)((b[r] << 32) | (a[r] & 0xffffffffull)));
            s = (r == 0) ? v : s + v;
        }
    buf[i] = s;
}

}  // namespace
}  // namespace nle

using namespace nle;

struct nle_b200_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory path
    bool peer_ok = false;
    uint4* inbox = nullptr;                 // this rank's inbox (device memory of the communicator's device)
    PeerInboxes peers = {};                 // peers.p[r] = rank r's inbox mapped into this process (p[rank] = inbox)
    unsigned calls = 0;                     // peer all-reduces issued so far (the tag of the next one is calls + 1)
    unsigned long long n_peer = 0, n_nccl = 0;
    std::string peer_why = "not attempted";
};

namespace {

// Collective over the communicator: allocate and export the inbox, all-gather the IPC handles, map the peers' inboxes,
// agree (all-reduce of a flag) on whether EVERY rank can reach EVERY inbox.  Any local failure only clears the flag;
// the collectives are still executed so that no rank is left waiting.
void peer_setup(nle_b200_comm* h) {
    NcclApi& api = nccl_api();
    if (const char* e = getenv("NLE_B200_PEER_AR"))
        if (std::string(e) == "off") { h->peer_why = "disabled (NLE_B200_PEER_AR=off)"; return; }     // same environment on every rank
    if (!api.AllGather) { h->peer_why = "libnccl has no ncclAllGather"; return; }
    if (h->nranks < 2 || h->nranks > kPeerRanks) { h->peer_why = "needs 2.." + std::to_string(kPeerRanks) + " ranks"; return; }
    const int R = h->nranks;
    bool ok = true;
    std::string why;
    auto bad = [&](const char* what, cudaError_t e) { if (ok) { ok = false; why = std::string(what) + ": " + cudaGetErrorString(e); } cudaGetLastError(); };
    const size_t bytes = (size_t)2 * kPeerRanks * kPeerMax * sizeof(uint4);
    cudaError_t e = cudaMalloc(&h->inbox, bytes);
    if (e != cudaSuccess) { bad("cudaMalloc(inbox)", e); h->inbox = nullptr; }
    if (h->inbox && (e = cudaMemset(h->inbox, 0, bytes)) != cudaSuccess) bad("cudaMemset(inbox)", e);     // tag 0 = never written
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (h->inbox && (e = cudaIpcGetMemHandle(&mine, h->inbox)) != cudaSuccess) bad("cudaIpcGetMemHandle", e);
    // all-gather of the handles (device staging buffers; failures here would desynchronise the ranks, so they are fatal
    // for the peer path only if they are local allocation failures that every later step checks again)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    unsigned char* stage = nullptr;
    std::vector<cudaIpcMemHandle_t> all(R);
    bool gathered = false;
    if (cudaMalloc(&stage, (size_t)(R + 1) * 64) == cudaSuccess) {
        cudaMemcpy(stage, &mine, 64, cudaMemcpyHostToDevice);
        const int rc = api.AllGather(stage, stage + 64, 64, kNcclChar, h->comm, (cudaStream_t)0);
        if (rc == 0 && cudaStreamSynchronize(0) == cudaSuccess &&
            cudaMemcpy(all.data(), stage + 64, (size_t)R * 64, cudaMemcpyDeviceToHost) == cudaSuccess)
            gathered = true;
        else { ok = false; if (why.empty()) why = "all-gather of the IPC handles failed"; cudaGetLastError(); }
    } else { ok = false; why = "cudaMalloc(staging) failed"; cudaGetLastError(); }
    if (ok && gathered) {
        for (int r = 0; r < R; ++r) {
            if (r == h->rank) { h->peers.p[r] = h->inbox; continue; }
            void* ptr = nullptr;
            e = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { bad("cudaIpcOpenMemHandle", e); break; }
            h->peers.p[r] = static_cast<uint4*>(ptr);
        }
    }
    // agreement: the peer path is used only if it works everywhere
    bool agreed = false;
    if (stage) {
        double flag = ok ? 1.0 : 0.0, sum = 0.0;
        double* dflag = reinterpret_cast<double*>(stage);
        if (cudaMemcpy(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
            api.AllReduce(dflag, dflag, 1, kNcclFloat64, kNcclSum, h->comm, (cudaStream_t)0) == 0 &&
            cudaStreamSynchronize(0) == cudaSuccess &&
            cudaMemcpy(&sum, dflag, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess)
            agreed = (sum == (double)R);
        cudaGetLastError();
        cudaFree(stage);
    }
    h->peer_ok = ok && agreed;
    h->peer_why = h->peer_ok ? "ok" : (ok ? "a peer rank could not map the inboxes" : why);
    if (!h->peer_ok && getenv("NLE_B200_DEBUG"))
        fprintf(stderr, "[nle_b200 comm rank %d] peer-memory all-reduce unavailable (%s): small messages use ncclAllReduce\n", h->rank,
                h->peer_why.c_str());
}

void peer_teardown(nle_b200_comm* h) {
    for (int r = 0; r < h->nranks && r < kPeerRanks; ++r)
        if (r != h->rank && h->peers.p[r]) cudaIpcCloseMemHandle(h->peers.p[r]);
    if (h->inbox) cudaFree(h->inbox);
    cudaGetLastError();
}

template <int R>
void launch_peer(nle_b200_comm* h, double* buf, int count, cudaStream_t s) {
    unsigned tag = ++h->calls;
    if (tag == 0) tag = h->calls = 2;       // 2^32 calls later: skip 0 (= never written), keep the parity sequence
    peer_allreduce_kernel<R><<<cdiv(count, 256), 256, 0, s>>>(buf, count, h->peers, h->rank, h->nranks, tag);
    NLE_LAUNCH_CHECK();
}

}  // namespace

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
    peer_setup(h);
    *out = h;
    return NLE_B200_OK;
}

/* Which path the small all-reduces of this communicator take: *peer_path = 1 if they are peer_allreduce_kernel launches
 * over NVLink peer memory, 0 if ncclAllReduce; calls[0] / calls[1] = reductions issued so far on the peer / NCCL path.
 * Returns a static description ("ok" or why the peer path is unavailable). */
const char* nle_b200_comm_info(nle_b200_comm* h, int* peer_path, unsigned long long calls[2]) {
    if (!h) return "null communicator";
    if (peer_path) *peer_path = h->peer_ok ? 1 : 0;
    if (calls) { calls[0] = h->n_peer; calls[1] = h->n_nccl; }
    return h->peer_why.c_str();
}

/* An nle_b200_allreduce_fn: pass it as `allreduce` with `user` = the nle_b200_comm*.  In-place FP64 sum on `cuda_stream`. */
int nle_b200_comm_allreduce(void* dev_buf, size_t count, void* cuda_stream, void* user) {
    auto* h = static_cast<nle_b200_comm*>(user);
    if (!h || !h->comm) return 1;
    if (h->nranks == 1 || count == 0) return 0;
    if (h->peer_ok && count <= (size_t)kPeerMax) {
        try {
            cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
            double* buf = static_cast<double*>(dev_buf);
            if (h->nranks <= 2) launch_peer<2>(h, buf, (int)count, s);
            else if (h->nranks <= 4) launch_peer<4>(h, buf, (int)count, s);
            else if (h->nranks <= 8) launch_peer<8>(h, buf, (int)count, s);
            else launch_peer<16>(h, buf, (int)count, s);
        } catch (const CudaError& ce) { set_error(std::string("peer all-reduce launch: ") + cudaGetErrorString(ce.e)); return 1; }
        ++h->n_peer;
        return 0;
    }
    ++h->n_nccl;
    const int rc = nccl_api().AllReduce(dev_buf, dev_buf, count, kNcclFloat64, kNcclSum, h->comm, static_cast<cudaStream_t>(cuda_stream));
    if (rc != 0) { set_error(std::string("ncclAllReduce: ") + nccl_api().GetErrorString(rc)); return rc; }
    return 0;
}

void nle_b200_comm_destroy(nle_b200_comm* h) {
    if (!h) return;
    peer_teardown(h);
    if (h->comm) nccl_api().CommDestroy(h->comm);
    delete h;
}

}  // extern "C"
