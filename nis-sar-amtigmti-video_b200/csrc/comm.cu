// comm.cu -- the two exchange steps of the hot path as C entry points over NCCL:
//   nis_echo_reduce   partial echoes of scatterer shards summed across ranks (config 3, sar_vehicle_sim.py:83-126 model
//                     with T = 1e5 sharded by scatterer): all-reduce, or reduce to one root
//   nis_slc_exchange  ring shift of one focused channel to the neighbour that pairs it (HRWS / ATI channel per GPU,
//                     generalising sar_ati_dcpa_sim_csa.py:184-197, :402-419): rank k receives channel k+1
// The fused alternatives -- the echo kernel reducing into the owner's HBM (accumulate = 2), the DPCA/ATI kernel reading the
// neighbour's image in place -- need no entry point of their own: they are nis_echo_accumulate / nis_gmti_fused on
// peer-mapped pointers (nis_peer_*).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by the host process, e.g. torch's): the library
// keeps loading on machines without NCCL, and nis_comm_* then fail with NIS_ERR_UNSUPPORTED.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

using namespace nis;

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
constexpr int kNcclFloat = 7, kNcclSum = 0;   // ncclFloat32, ncclSum (nccl.h; stable across NCCL 2.x)

struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

const NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
#define NIS_SYM(field, name) *(void**)(&api.field) = dlsym(h, name)
    NIS_SYM(GetUniqueId, "ncclGetUniqueId");
    NIS_SYM(CommInitRank, "ncclCommInitRank");
    NIS_SYM(CommDestroy, "ncclCommDestroy");
    NIS_SYM(AllReduce, "ncclAllReduce");
    NIS_SYM(Reduce, "ncclReduce");
    NIS_SYM(Send, "ncclSend");
    NIS_SYM(Recv, "ncclRecv");
    NIS_SYM(GroupStart, "ncclGroupStart");
    NIS_SYM(GroupEnd, "ncclGroupEnd");
    NIS_SYM(GetErrorString, "ncclGetErrorString");
#undef NIS_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Reduce && api.Send && api.Recv &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api;
}

#define NIS_NCCL_TRY(expr)                                                                         \
    do {                                                                                           \
        const int _r = (expr);                                                                     \
        if (_r != 0) {                                                                             \
            set_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
            return NIS_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int require_nccl() {
    if (!nccl().ok) {
        set_error("NCCL (libnccl.so.2) is not available in this process");
        return NIS_ERR_UNSUPPORTED;
    }
    return NIS_OK;
}

}  // namespace

struct nis_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

extern "C" int nis_comm_unique_id(uint8_t* id128) {
    NIS_REQUIRE(id128, "nis_comm_unique_id: null argument");
    int rc = require_nccl();
    if (rc != NIS_OK) return rc;
    ncclUniqueId id;
    NIS_NCCL_TRY(nccl().GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return NIS_OK;
}

extern "C" int nis_comm_init(int32_t rank, int32_t nranks, const uint8_t* id128, nis_comm** out) {
    NIS_REQUIRE(id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, "nis_comm_init: bad argument");
    int rc = require_nccl();
    if (rc != NIS_OK) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    nis_comm* c = new nis_comm();
    c->rank = rank;
    c->nranks = nranks;
    const int r = nccl().CommInitRank(&c->comm, nranks, id, rank);   // binds to the calling thread's current device
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
        delete c;
        return NIS_ERR_CUDA;
    }
    *out = c;
    return NIS_OK;
}

extern "C" int nis_comm_destroy(nis_comm* c) {
    if (!c) return NIS_OK;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
    return NIS_OK;
}

extern "C" int nis_echo_reduce(nis_comm* c, nis_c32* raw, uint64_t n_samples, int32_t root, nis_stream stream) {
    NIS_REQUIRE(c && raw, "nis_echo_reduce: null argument");
    NIS_REQUIRE(root >= -1 && root < c->nranks, "nis_echo_reduce: root %d outside -1..%d", root, c->nranks - 1);
    if (c->nranks == 1 || n_samples == 0) return NIS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // complex64 samples travel as pairs of floats: NCCL sums real dtypes
    if (root < 0) NIS_NCCL_TRY(nccl().AllReduce(raw, raw, 2 * n_samples, kNcclFloat, kNcclSum, c->comm, st));
    else NIS_NCCL_TRY(nccl().Reduce(raw, raw, 2 * n_samples, kNcclFloat, kNcclSum, root, c->comm, st));
    return NIS_OK;
}

extern "C" int nis_slc_exchange(nis_comm* c, const nis_c32* slc_local, nis_c32* slc_next, uint64_t n_pix,
                                nis_stream stream) {
    NIS_REQUIRE(c && slc_local, "nis_slc_exchange: null argument");
    NIS_REQUIRE(c->rank == c->nranks - 1 || slc_next, "nis_slc_exchange: every rank but the last needs a receive buffer");
    if (c->nranks == 1 || n_pix == 0) return NIS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    NIS_NCCL_TRY(nccl().GroupStart());
    int r1 = 0, r2 = 0;
    if (c->rank > 0) r1 = nccl().Send(slc_local, 2 * n_pix, kNcclFloat, c->rank - 1, c->comm, st);
    if (c->rank < c->nranks - 1) r2 = nccl().Recv(slc_next, 2 * n_pix, kNcclFloat, c->rank + 1, c->comm, st);
    const int r3 = nccl().GroupEnd();
    if (r1 || r2 || r3) {
        set_error("nis_slc_exchange: NCCL send/recv failed: %s", nccl().GetErrorString(r1 ? r1 : (r2 ? r2 : r3)));
        return NIS_ERR_CUDA;
    }
    return NIS_OK;
}

// 1 when `device` can perform native atomics on memory of `peer` over their link (NVLink: yes; PCIe without atomics: 0)
extern "C" int nis_peer_native_atomics(int32_t device, int32_t peer) {
    if (device == peer) return 1;
    int v = 0;
    if (cudaDeviceGetP2PAttribute(&v, cudaDevP2PAttrNativeAtomicSupported, device, peer) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return v;
}
