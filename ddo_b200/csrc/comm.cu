// comm.cu -- ddo_comm_*: the collectives of the fringe-sharded search at the C ABI (include/ddo_b200.h), NCCL over NVLink / NVSwitch.
//
// The reference shares one `Mutex<Critical>` between its worker threads (implementation/solver/parallel.rs:32-81); between processes (one
// per GPU) the same three words -- incumbent lower bound, bound of the best open node, "work remains" -- travel through ONE collective per
// wave, and open nodes move point to point when the fringes get out of balance.  NCCL is resolved at run time (dlopen of libnccl.so.2, the
// library torch ships or the system's), so that libddo_b200.so loads on hosts without it; every entry point fails loudly when it is absent.
#include <dlfcn.h>

#include <cstring>
#include <string>

#include <cuda_runtime.h>

#include "engine.hpp"

namespace {

struct ncclComm;
typedef ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
enum { NCCL_INT8 = 0, NCCL_INT64 = 4 };  // ncclDataType_t: ncclInt8 = ncclChar = 0, ncclInt64 = 4
enum { NCCL_MAX = 2 };                   // ncclRedOp_t: sum 0, prod 1, max 2, min 3

struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Api& api() {
    static Api a;
    if (a.lib) return a;
    // RTLD_NOLOAD first: a host process that already carries NCCL (torch) must not get a second copy
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
        if (!a.lib) a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    auto sym = [&](const char* n) { return dlsym(a.lib, n); };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Send && a.Recv;
    return a;
}

int fail(const std::string& what, int code = DDO_ERR_CUDA) { ddo::set_error(what); return code; }
int nccl_fail(const char* where, int rc) {
    Api& a = api();
    return fail(std::string(where) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
}

}  // namespace

struct ddo_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    void* d_buf = nullptr; void* h_buf = nullptr; size_t cap = 0;  // device scratch + pinned mirror
    int reserve(size_t bytes) {
        if (bytes <= cap) return DDO_OK;
        if (d_buf) cudaFree(d_buf);
        if (h_buf) cudaFreeHost(h_buf);
        cap = std::max<size_t>(bytes, 1 << 16);
        if (cudaMalloc(&d_buf, cap) != cudaSuccess || cudaMallocHost(&h_buf, cap) != cudaSuccess) { cap = 0; return fail("ddo_comm: scratch allocation failed"); }
        return DDO_OK;
    }
};

extern "C" {

int ddo_comm_unique_id(void* id128) {
    if (!id128) return fail("ddo_comm_unique_id: null argument", DDO_ERR_INVALID);
    Api& a = api();
    if (!a.ok) return fail("NCCL (libnccl.so.2) is not available: there is no fallback transport", DDO_ERR_UNSUPPORTED);
    NcclId id;
    const int rc = a.GetUniqueId(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    std::memcpy(id128, &id, sizeof(id));
    return DDO_OK;
}

int ddo_comm_init(int32_t nranks, int32_t rank, const void* id128, int device, ddo_comm** out) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128 || !out) return fail("ddo_comm_init: invalid argument", DDO_ERR_INVALID);
    Api& a = api();
    if (!a.ok) return fail("NCCL (libnccl.so.2) is not available: there is no fallback transport", DDO_ERR_UNSUPPORTED);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("no CUDA device (there is no CPU fallback)", DDO_ERR_NO_DEVICE);
    if (cudaSetDevice(device) != cudaSuccess) return fail("ddo_comm_init: bad device", DDO_ERR_INVALID);
    ddo_comm* c = new ddo_comm();
    c->nranks = nranks; c->rank = rank; c->device = device;
    NcclId id;
    std::memcpy(&id, id128, sizeof(id));
    const int rc = a.CommInitRank(&c->comm, nranks, id, rank);
    if (rc != 0) { delete c; return nccl_fail("ncclCommInitRank", rc); }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { a.CommDestroy(c->comm); delete c; return fail("ddo_comm_init: stream"); }
    *out = c;
    return DDO_OK;
}

void ddo_comm_destroy(ddo_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm) api().CommDestroy(c->comm);
    if (c->d_buf) cudaFree(c->d_buf);
    if (c->h_buf) cudaFreeHost(c->h_buf);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int32_t ddo_comm_size(const ddo_comm* c) { return c ? c->nranks : 0; }
int32_t ddo_comm_rank(const ddo_comm* c) { return c ? c->rank : -1; }

int ddo_comm_allreduce_max(ddo_comm* c, int64_t* values, int32_t count) {
    if (!c || !values || count < 1) return fail("ddo_comm_allreduce_max: invalid argument", DDO_ERR_INVALID);
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)count * 8;
    int rc = c->reserve(bytes);
    if (rc != DDO_OK) return rc;
    std::memcpy(c->h_buf, values, bytes);
    if (cudaMemcpyAsync(c->d_buf, c->h_buf, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return fail("ddo_comm: H2D");
    const int nrc = api().AllReduce(c->d_buf, c->d_buf, (size_t)count, NCCL_INT64, NCCL_MAX, c->comm, c->stream);
    if (nrc != 0) return nccl_fail("ncclAllReduce", nrc);
    if (cudaMemcpyAsync(c->h_buf, c->d_buf, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("ddo_comm: D2H");
    std::memcpy(values, c->h_buf, bytes);
    return DDO_OK;
}

int ddo_comm_allgather(ddo_comm* c, const int64_t* values, int32_t count, int64_t* recv) {
    if (!c || !values || !recv || count < 1) return fail("ddo_comm_allgather: invalid argument", DDO_ERR_INVALID);
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)count * 8, all = bytes * (size_t)c->nranks;
    int rc = c->reserve(bytes + all);
    if (rc != DDO_OK) return rc;
    std::memcpy(c->h_buf, values, bytes);
    char* d = (char*)c->d_buf;
    if (cudaMemcpyAsync(d, c->h_buf, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return fail("ddo_comm: H2D");
    const int nrc = api().AllGather(d, d + bytes, (size_t)count, NCCL_INT64, c->comm, c->stream);
    if (nrc != 0) return nccl_fail("ncclAllGather", nrc);
    if (cudaMemcpyAsync((char*)c->h_buf + bytes, d + bytes, all, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("ddo_comm: D2H");
    std::memcpy(recv, (char*)c->h_buf + bytes, all);
    return DDO_OK;
}

int ddo_comm_send(ddo_comm* c, const void* buf, int64_t bytes, int32_t peer) {
    if (!c || (!buf && bytes > 0) || bytes < 0 || peer < 0 || peer >= c->nranks || peer == c->rank) return fail("ddo_comm_send: invalid argument", DDO_ERR_INVALID);
    if (bytes == 0) return DDO_OK;
    cudaSetDevice(c->device);
    int rc = c->reserve((size_t)bytes);
    if (rc != DDO_OK) return rc;
    std::memcpy(c->h_buf, buf, (size_t)bytes);
    if (cudaMemcpyAsync(c->d_buf, c->h_buf, (size_t)bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return fail("ddo_comm: H2D");
    const int nrc = api().Send(c->d_buf, (size_t)bytes, NCCL_INT8, peer, c->comm, c->stream);
    if (nrc != 0) return nccl_fail("ncclSend", nrc);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("ddo_comm: send sync");
    return DDO_OK;
}

int ddo_comm_recv(ddo_comm* c, void* buf, int64_t bytes, int32_t peer) {
    if (!c || (!buf && bytes > 0) || bytes < 0 || peer < 0 || peer >= c->nranks || peer == c->rank) return fail("ddo_comm_recv: invalid argument", DDO_ERR_INVALID);
    if (bytes == 0) return DDO_OK;
    cudaSetDevice(c->device);
    int rc = c->reserve((size_t)bytes);
    if (rc != DDO_OK) return rc;
    const int nrc = api().Recv(c->d_buf, (size_t)bytes, NCCL_INT8, peer, c->comm, c->stream);
    if (nrc != 0) return nccl_fail("ncclRecv", nrc);
    if (cudaMemcpyAsync(c->h_buf, c->d_buf, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("ddo_comm: D2H");
    std::memcpy(buf, c->h_buf, (size_t)bytes);
    return DDO_OK;
}

}  // extern "C"
