// xrft_b200 -- the one exchange step of the path: sum of radial-bin partials over the ranks that share the sharded
// (non-transform) axis (reference role: the `.mean` over a dask-chunked axis after
// isotropic_power_spectrum, xrft/xrft.py:1013-1095, xrft/tests/test_xrft.py:1011-1013).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a torch process that resolves to the copy torch already
// loaded), so the library has no link-time dependency on it and single-GPU users never touch it.  One communicator per
// process/GPU; the unique id travels through whatever side channel the host has (torch.distributed in xrft_b200/shard.py).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include "../../include/xrft_b200.h"
#include "internal.h"

namespace {

struct NcclId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef int (*GetUniqueId_t)(NcclId*);
typedef int (*CommInitRank_t)(void**, int, NcclId, int);
typedef int (*CommDestroy_t)(void*);
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*GetErrorString_t)(int);
typedef int (*GetVersion_t)(int*);

struct NcclApi {
    void* handle = nullptr;
    GetUniqueId_t get_unique_id = nullptr;
    CommInitRank_t comm_init_rank = nullptr;
    CommDestroy_t comm_destroy = nullptr;
    AllReduce_t all_reduce = nullptr;
    GetErrorString_t error_string = nullptr;
    GetVersion_t get_version = nullptr;
};

std::mutex g_mu;
NcclApi g_nccl;

const NcclApi* nccl() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_nccl.handle) return &g_nccl;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { xrftb::set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return nullptr; }
    NcclApi a;
    a.handle = h;
    a.get_unique_id = reinterpret_cast<GetUniqueId_t>(dlsym(h, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<CommInitRank_t>(dlsym(h, "ncclCommInitRank"));
    a.comm_destroy = reinterpret_cast<CommDestroy_t>(dlsym(h, "ncclCommDestroy"));
    a.all_reduce = reinterpret_cast<AllReduce_t>(dlsym(h, "ncclAllReduce"));
    a.error_string = reinterpret_cast<GetErrorString_t>(dlsym(h, "ncclGetErrorString"));
    a.get_version = reinterpret_cast<GetVersion_t>(dlsym(h, "ncclGetVersion"));
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce) {
        xrftb::set_error("libnccl lacks a required symbol");
        return nullptr;
    }
    g_nccl = a;
    return &g_nccl;
}

int nccl_fail(const NcclApi* api, const char* what, int rc) {
    xrftb::set_error("%s: %s", what, api->error_string ? api->error_string(rc) : "NCCL error");
    return XRFTB_ECUDA;
}

struct Comm { void* nccl_comm; int nranks, rank; };

}  // namespace

extern "C" {

int xrftb_comm_unique_id(void* id128) {
    if (!id128) { xrftb::set_error("comm_unique_id: NULL buffer"); return XRFTB_EINVAL; }
    const NcclApi* api = nccl();
    if (!api) return XRFTB_EUNSUPPORTED;
    NcclId id;
    if (int rc = api->get_unique_id(&id)) return nccl_fail(api, "ncclGetUniqueId", rc);
    memcpy(id128, id.internal, sizeof(id.internal));
    return 0;
}

int xrftb_comm_init(void** comm, int nranks, int rank, const void* id128) {
    if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { xrftb::set_error("comm_init: bad arguments"); return XRFTB_EINVAL; }
    const NcclApi* api = nccl();
    if (!api) return XRFTB_EUNSUPPORTED;
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    void* c = nullptr;
    if (int rc = api->comm_init_rank(&c, nranks, id, rank)) return nccl_fail(api, "ncclCommInitRank", rc);
    *comm = new Comm{c, nranks, rank};
    return 0;
}

int xrftb_comm_destroy(void* comm) {
    if (!comm) return 0;
    const NcclApi* api = nccl();
    Comm* c = reinterpret_cast<Comm*>(comm);
    int rc = api ? api->comm_destroy(c->nccl_comm) : 0;
    delete c;
    return rc ? nccl_fail(api, "ncclCommDestroy", rc) : 0;
}

int xrftb_comm_nccl_version(void) {
    const NcclApi* api = nccl();
    int v = 0;
    if (!api || !api->get_version || api->get_version(&v)) return 0;
    return v;
}

int xrftb_allreduce_bins(void* comm, double* buf, size_t count, void* stream) {
    if (!comm || !buf) { xrftb::set_error("allreduce_bins: bad arguments"); return XRFTB_EINVAL; }
    if (count == 0) return 0;
    const NcclApi* api = nccl();
    if (!api) return XRFTB_EUNSUPPORTED;
    Comm* c = reinterpret_cast<Comm*>(comm);
    // ncclFloat64 = 8, ncclSum = 0; in place
    if (int rc = api->all_reduce(buf, buf, count, 8, 0, c->nccl_comm, reinterpret_cast<cudaStream_t>(stream)))
        return nccl_fail(api, "ncclAllReduce", rc);
    return 0;
}

}  // extern "C"
