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
    int (*send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
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
        api.send = reinterpret_cast<int (*)(const void *, size_t, int, int, NcclComm, cudaStream_t)>(dlsym(api.lib, "ncclSend"));
        api.recv = reinterpret_cast<int (*)(void *, size_t, int, int, NcclComm, cudaStream_t)>(dlsym(api.lib, "ncclRecv"));
        api.broadcast = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t)>(
            dlsym(api.lib, "ncclBroadcast"));
        api.group_start = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupStart"));
        api.group_end = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupEnd"));
        api.get_error_string = reinterpret_cast<const char *(*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
        api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather && api.send && api.recv &&
                 api.broadcast && api.group_start && api.group_end;
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

// Equal-length segments onto EVERY rank (ncclAllGather).  A collective: every rank of the communicator must call it
// with the same n_samples (0 on all ranks is a no-op on all ranks; a rank may not skip the call on its own).  Segments
// of different lengths, or a stream wanted on one rank only: cb_gather_segments_to_root_dev / cb_allgather_segments_var_dev.
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

static int check_counts(const cb_comm *c, const size_t *counts, size_t elem_bytes)
{
    CB_REQUIRE(c, CB_ERR_INVALID_ARG, "comm is NULL");
    CB_REQUIRE(counts, CB_ERR_INVALID_ARG, "counts is NULL");
    CB_REQUIRE(elem_bytes >= 1, CB_ERR_INVALID_ARG, "elem_bytes must be >= 1");
    return CB_OK;
}

// Ordered gather of per-rank segments of ANY lengths onto one rank: counts[r] elements of elem_bytes each from rank r
// land at d_all + sum_{q<r} counts[q] on `root`, nowhere else.  Every rank passes the same counts[] (the segment table
// is host knowledge: sharding.segment_bounds), so empty segments are skipped consistently and no rank can block alone.
// One grouped ncclSend/ncclRecv: root's NVLink ingress carries (total - own) bytes once, 1/nranks of what an all-gather
// moves through the switch.  d_all may be NULL on the other ranks.
int cb_gather_segments_to_root_dev(cb_comm *c, const void *d_seg, const size_t *counts, size_t elem_bytes, int root,
                                   void *d_all, void *stream)
{
    int rc = check_counts(c, counts, elem_bytes);
    if (rc) return rc;
    CB_REQUIRE(root >= 0 && root < c->nranks, CB_ERR_INVALID_ARG, "gather: root %d of %d ranks", root, c->nranks);
    const size_t mine = counts[c->rank];
    CB_REQUIRE(d_seg || mine == 0, CB_ERR_INVALID_ARG, "NULL segment pointer");
    CB_CUDA(cudaSetDevice(c->device));
    NcclApi &a = nccl();
    cudaStream_t s = (cudaStream_t)stream;
    if (c->rank != root) {
        if (mine == 0) return CB_OK;
        rc = a.send(d_seg, mine * elem_bytes, 0 /* ncclInt8 */, root, c->comm, s);
        return rc ? nccl_fail(rc, "ncclSend") : CB_OK;
    }
    size_t total = 0;
    for (int r = 0; r < c->nranks; ++r) total += counts[r];
    CB_REQUIRE(d_all || total == 0, CB_ERR_INVALID_ARG, "NULL destination on the root rank");
    rc = a.group_start();
    if (rc) return nccl_fail(rc, "ncclGroupStart");
    size_t off = 0;
    int rc_recv = 0;
    for (int r = 0; r < c->nranks; ++r) {
        char *dst = static_cast<char *>(d_all) + off * elem_bytes;
        if (r != root && counts[r] != 0 && rc_recv == 0) rc_recv = a.recv(dst, counts[r] * elem_bytes, 0, r, c->comm, s);
        off += counts[r];
    }
    rc = a.group_end();
    if (rc_recv) return nccl_fail(rc_recv, "ncclRecv");
    if (rc) return nccl_fail(rc, "ncclGroupEnd");
    if (mine) {
        size_t my_off = 0;
        for (int r = 0; r < root; ++r) my_off += counts[r];
        char *dst = static_cast<char *>(d_all) + my_off * elem_bytes;
        if (dst != d_seg) CB_CUDA(cudaMemcpyAsync(dst, d_seg, mine * elem_bytes, cudaMemcpyDeviceToDevice, s));
    }
    return CB_OK;
}

// The same ordered stream on EVERY rank, segments of any lengths: one grouped broadcast per non-empty segment.
int cb_allgather_segments_var_dev(cb_comm *c, const void *d_seg, const size_t *counts, size_t elem_bytes, void *d_all,
                                  void *stream)
{
    int rc = check_counts(c, counts, elem_bytes);
    if (rc) return rc;
    size_t total = 0;
    for (int r = 0; r < c->nranks; ++r) total += counts[r];
    if (total == 0) return CB_OK;
    CB_REQUIRE(d_all, CB_ERR_INVALID_ARG, "NULL destination");
    CB_REQUIRE(d_seg || counts[c->rank] == 0, CB_ERR_INVALID_ARG, "NULL segment pointer");
    CB_CUDA(cudaSetDevice(c->device));
    NcclApi &a = nccl();
    rc = a.group_start();
    if (rc) return nccl_fail(rc, "ncclGroupStart");
    size_t off = 0;
    int rc_b = 0;
    for (int r = 0; r < c->nranks; ++r) {
        char *dst = static_cast<char *>(d_all) + off * elem_bytes;
        if (counts[r] != 0 && rc_b == 0)
            rc_b = a.broadcast(r == c->rank ? d_seg : dst, dst, counts[r] * elem_bytes, 0, r, c->comm, (cudaStream_t)stream);
        off += counts[r];
    }
    rc = a.group_end();
    if (rc_b) return nccl_fail(rc_b, "ncclBroadcast");
    if (rc) return nccl_fail(rc, "ncclGroupEnd");
    return CB_OK;
}

int cb_comm_rank(const cb_comm *c, int *rank, int *nranks)
{
    CB_REQUIRE(c, CB_ERR_INVALID_ARG, "comm is NULL");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return CB_OK;
}

// ---- peer-mapped output: the gather done by the producing kernel's own stores ----------------------------------------
// The root exports its gathered-stream buffer (cudaIpcGetMemHandle, 64 bytes, passed over any host channel); every other
// rank maps it (cb_peer_open) and hands `mapped + its offset` to *_run_dev as d_out: the kernel's stores go over
// NVLink / NVSwitch straight into the root's HBM, tile by tile while the filter runs, and there is no second pass.
int cb_peer_export(void *d_ptr, void *handle64)
{
    CB_REQUIRE(d_ptr && handle64, CB_ERR_INVALID_ARG, "NULL argument");
    int rc = ensure_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, sizeof h);
    return CB_OK;
}

int cb_peer_open(const void *handle64, void **d_mapped)
{
    CB_REQUIRE(handle64 && d_mapped, CB_ERR_INVALID_ARG, "NULL argument");
    int rc = ensure_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    CB_CUDA(cudaIpcOpenMemHandle(d_mapped, h, cudaIpcMemLazyEnablePeerAccess));
    return CB_OK;
}

int cb_peer_close(void *d_mapped)
{
    if (!d_mapped) return CB_OK;
    int rc = ensure_device();
    if (rc) return rc;
    CB_CUDA(cudaIpcCloseMemHandle(d_mapped));
    return CB_OK;
}

}  // extern "C"
