// Multi-GPU edge of the C ABI: the ordered gather of per-rank output segments ("NCCL is
// used only to gather the segmented outputs back into one ordered stream").  One process (or thread) per GPU, as
// everywhere in this library; the data path itself needs no collective.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy PyTorch ships is found when it is already loaded,
// COMMS_B200_NCCL_LIB names another one), so libcomms_b200.so has no link-time dependency on it and single-GPU
// users never touch it.  Only the entry points used here are declared; their signatures are NCCL's public ABI.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace cb {

struct NcclUniqueId {
    char internal[128];
};
typedef void *NcclComm;

struct NcclApi {
    void *lib = nullptr;
    int (*get_unique_id)(NcclUniqueId *) = nullptr;
    int (*comm_init_rank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*comm_destroy)(NcclComm) = nullptr;
    int (*all_gather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*get_error_string)(int) = nullptr;
    bool ok = false;
};

static NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[3] = {getenv("COMMS_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (n == nullptr || *n == 0) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.get_unique_id = reinterpret_cast<int (*)(NcclUniqueId *)>(dlsym(api.lib, "ncclGetUniqueId"));
        api.comm_init_rank = reinterpret_cast<int (*)(NcclComm *, int, NcclUniqueId, int)>(dlsym(api.lib, "ncclCommInitRank"));
        api.comm_destroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(api.lib, "ncclCommDestroy"));
        api.all_gather = reinterpret_cast<int (*)(const void *, void *, size_t, int, NcclComm, cudaStream_t)>(
            dlsym(api.lib, "ncclAllGather"));
        api.get_error_string = reinterpret_cast<const char *(*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
        api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather;
    });
    return api;
}

static int nccl_fail(int rc, const char *what)
{
    NcclApi &a = nccl();
    set_error("%s: %s", what, a.get_error_string ? a.get_error_string(rc) : "NCCL error");
    return CB_ERR_CUDA;
}

}  // namespace cb

using namespace cb;

struct cb_comm {
    int device;
    int nranks, rank;
    NcclComm comm;
};

extern "C" {

int cb_comm_unique_id(void *id128)
{
    CB_REQUIRE(id128, CB_ERR_INVALID_ARG, "id buffer is NULL");
    NcclApi &a = nccl();
    CB_REQUIRE(a.ok, CB_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded; set COMMS_B200_NCCL_LIB");
    NcclUniqueId id;
    const int rc = a.get_unique_id(&id);
    if (rc) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof id);
    return CB_OK;
}

int cb_comm_init(int nranks, int rank, const void *id128, cb_comm **out)
{
    CB_REQUIRE(out && id128, CB_ERR_INVALID_ARG, "NULL argument");
    CB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, CB_ERR_INVALID_ARG, "comm: rank %d of %d", rank, nranks);
    int rc = ensure_device();
    if (rc) return rc;
    NcclApi &a = nccl();
    CB_REQUIRE(a.ok, CB_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded; set COMMS_B200_NCCL_LIB");
    cb_comm *c = new (std::nothrow) cb_comm();
    CB_REQUIRE(c, CB_ERR_OOM, "host allocation failed");
    c->device = current_device();
    c->nranks = nranks;
    c->rank = rank;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof id);
    rc = a.comm_init_rank(&c->comm, nranks, id, rank);
    if (rc) {
        delete c;
        return nccl_fail(rc, "ncclCommInitRank");
    }
    *out = c;
    return CB_OK;
}

int cb_comm_destroy(cb_comm *c)
{
    if (!c) return CB_OK;
    cudaSetDevice(c->device);
    nccl().comm_destroy(c->comm);
    delete c;
    return CB_OK;
}

int cb_gather_segments_dev(cb_comm *c, const float *d_seg, size_t n_samples, float *d_all, void *stream)
{
    CB_REQUIRE(c, CB_ERR_INVALID_ARG, "comm is NULL");
    if (n_samples == 0) return CB_OK;
    CB_REQUIRE(d_seg && d_all, CB_ERR_INVALID_ARG, "NULL data pointer");
    CB_CUDA(cudaSetDevice(c->device));
    // rank r's n_samples complex samples land at d_all[r * n_samples ...] on every rank: the segments in stream order
    const int rc = nccl().all_gather(d_seg, d_all, 2 * n_samples, 7 /* ncclFloat32 */, c->comm, (cudaStream_t)stream);
    if (rc) return nccl_fail(rc, "ncclAllGather");
    return CB_OK;
}

}  // extern "C"
