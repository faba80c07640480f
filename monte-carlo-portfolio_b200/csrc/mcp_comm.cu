// Multi-GPU plumbing INSIDE libmcp: one NCCL communicator per handle.
//
// The path shards by index range and has no data-path collective; what crosses NVLink are the merges of tiny
// results: one fixed-size selection record per rank (ncclAllGather), histogram counts and tail sums of the exact
// radix select (ncclAllReduce, sum), risk ranges (min / max) and envelope bins.  All of them are issued on the
// handle's stream, between the library's own kernels, so a sharded call costs one host wait -- and a C / C++
// consumer of include/mcp.h gets the multi-GPU path without torch.distributed.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already loaded NCCL (torch ships its own
// copy under the same soname) keeps using that one copy, and a process that never calls mcp_comm_* needs no NCCL.
#include <dlfcn.h>
#include <cstring>
#include <mutex>

#include <nccl.h>

#include "mcp_context.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclCommGetAsyncError) CommGetAsyncError = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        const char* e = dlerror();
        g_nccl.error = std::string("cannot load libnccl.so.2: ") + (e ? e : "unknown dlopen error");
        return;
    }
#define MCP_NCCL_SYM(field, sym)                                                              \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, #sym));          \
    if (!g_nccl.field) { g_nccl.error = "libnccl.so.2 lacks " #sym; return; }
    MCP_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
    MCP_NCCL_SYM(CommInitRank, ncclCommInitRank)
    MCP_NCCL_SYM(CommDestroy, ncclCommDestroy)
    MCP_NCCL_SYM(CommAbort, ncclCommAbort)
    MCP_NCCL_SYM(AllReduce, ncclAllReduce)
    MCP_NCCL_SYM(AllGather, ncclAllGather)
    MCP_NCCL_SYM(GroupStart, ncclGroupStart)
    MCP_NCCL_SYM(GroupEnd, ncclGroupEnd)
    MCP_NCCL_SYM(CommGetAsyncError, ncclCommGetAsyncError)
    MCP_NCCL_SYM(GetErrorString, ncclGetErrorString)
    MCP_NCCL_SYM(GetVersion, ncclGetVersion)
#undef MCP_NCCL_SYM
}

const NcclApi* nccl_api(std::string* why) {
    std::call_once(g_nccl_once, nccl_load);
    if (!g_nccl.error.empty()) {
        if (why) *why = g_nccl.error;
        return nullptr;
    }
    return &g_nccl;
}

}  // namespace

struct mcp_comm_state {
    ncclComm_t comm = nullptr;
};

#define MCP_NCCL(h, api, call)                                                                       \
    do {                                                                                             \
        ncclResult_t _r = (call);                                                                    \
        if (_r != ncclSuccess)                                                                       \
            return mcp_fail((h), MCP_ERR_COMM, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

static_assert(sizeof(ncclUniqueId) == MCP_COMM_ID_BYTES, "MCP_COMM_ID_BYTES must match ncclUniqueId");

// ---- internal (device-buffer) collectives used by the entry points: ENQUEUE on `st`, no host wait -------------
int mcp_comm_allgather_dev(mcp_context* h, const void* send_dev, void* recv_dev, size_t bytes, cudaStream_t st) {
    if (!h->comm) return mcp_fail(h, MCP_ERR_COMM, "no communicator on this handle (call mcp_comm_init first)");
    const NcclApi* api = nccl_api(nullptr);
    MCP_NCCL(h, api, api->AllGather(send_dev, recv_dev, bytes, ncclChar, h->comm->comm, st));
    return MCP_OK;
}

int mcp_comm_allreduce_dev(mcp_context* h, void* buf_dev, size_t count, int kind, cudaStream_t st) {
    if (!h->comm) return mcp_fail(h, MCP_ERR_COMM, "no communicator on this handle (call mcp_comm_init first)");
    const NcclApi* api = nccl_api(nullptr);
    ncclDataType_t dt = ncclUint64;
    ncclRedOp_t op = ncclSum;
    switch (kind) {
        case MCP_REDUCE_U64_SUM: dt = ncclUint64; op = ncclSum; break;
        case MCP_REDUCE_F64_SUM: dt = ncclDouble; op = ncclSum; break;
        case MCP_REDUCE_F64_MIN: dt = ncclDouble; op = ncclMin; break;
        case MCP_REDUCE_F64_MAX: dt = ncclDouble; op = ncclMax; break;
        case MCP_REDUCE_U64_MAX: dt = ncclUint64; op = ncclMax; break;
        default: return mcp_fail(h, MCP_ERR_INVALID, "mcp_comm_allreduce: bad reduction kind %d", kind);
    }
    MCP_NCCL(h, api, api->AllReduce(buf_dev, buf_dev, count, dt, op, h->comm->comm, st));
    return MCP_OK;
}

// after the host has waited for the stream: did the communicator report an asynchronous failure?
int mcp_comm_check(mcp_context* h) {
    if (!h->comm) return MCP_OK;
    const NcclApi* api = nccl_api(nullptr);
    ncclResult_t async = ncclSuccess;
    MCP_NCCL(h, api, api->CommGetAsyncError(h->comm->comm, &async));
    if (async != ncclSuccess && async != ncclInProgress)
        return mcp_fail(h, MCP_ERR_COMM, "NCCL asynchronous error on rank %d of %d: %s", h->comm_rank, h->comm_size, api->GetErrorString(async));
    return MCP_OK;
}

void mcp_comm_release(mcp_context* h) {
    if (!h->comm) return;
    const NcclApi* api = nccl_api(nullptr);
    if (api && h->comm->comm) api->CommDestroy(h->comm->comm);
    delete h->comm;
    h->comm = nullptr;
    h->comm_rank = 0;
    h->comm_size = 0;
}

extern "C" {

int mcp_comm_unique_id(void* id_out) {
    if (!id_out) return mcp_fail(nullptr, MCP_ERR_INVALID, "mcp_comm_unique_id: id_out is NULL");
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!api) return mcp_fail(nullptr, MCP_ERR_COMM, "%s", why.c_str());
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return mcp_fail(nullptr, MCP_ERR_COMM, "ncclGetUniqueId failed: %s", api->GetErrorString(r));
    memcpy(id_out, &id, sizeof id);
    return MCP_OK;
}

int mcp_comm_init(mcp_handle h, const void* id, int rank, int nranks) {
    if (!h) return MCP_ERR_INVALID;
    MCP_REQUIRE(h, id != nullptr, "mcp_comm_init: id is NULL");
    MCP_REQUIRE(h, nranks >= 1 && rank >= 0 && rank < nranks, "mcp_comm_init: bad rank %d of %d", rank, nranks);
    MCP_REQUIRE(h, h->comm == nullptr, "mcp_comm_init: this handle already has a communicator (mcp_comm_destroy it first)");
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!api) return mcp_fail(h, MCP_ERR_COMM, "%s", why.c_str());
    mcp_device_guard guard(h->device);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    mcp_comm_state* c = new (std::nothrow) mcp_comm_state();
    if (!c) return mcp_fail(h, MCP_ERR_NOMEM, "mcp_comm_init: out of host memory");
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, uid, rank);
    if (r != ncclSuccess) {
        delete c;
        return mcp_fail(h, MCP_ERR_COMM, "ncclCommInitRank(rank %d of %d, device %d) failed: %s", rank, nranks, h->device, api->GetErrorString(r));
    }
    h->comm = c;
    h->comm_rank = rank;
    h->comm_size = nranks;
    return MCP_OK;
}

// One process, n handles (one per GPU): a communicator over all of them, built inside one NCCL group so a single
// host thread can do it (the collectives themselves are then issued by one host thread per handle, or inside
// ncclGroupStart / End by the caller's own scheduling -- mcportfolio uses one thread per device).
int mcp_comm_init_all(mcp_handle* handles, int n) {
    if (!handles || n < 1) return mcp_fail(nullptr, MCP_ERR_INVALID, "mcp_comm_init_all: bad arguments");
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!api) return mcp_fail(handles[0], MCP_ERR_COMM, "%s", why.c_str());
    for (int i = 0; i < n; ++i) {
        if (!handles[i]) return mcp_fail(nullptr, MCP_ERR_INVALID, "mcp_comm_init_all: handle %d is NULL", i);
        if (handles[i]->comm) return mcp_fail(handles[i], MCP_ERR_INVALID, "mcp_comm_init_all: handle %d already has a communicator", i);
        for (int j = 0; j < i; ++j)
            if (handles[j]->device == handles[i]->device)
                return mcp_fail(handles[i], MCP_ERR_INVALID, "mcp_comm_init_all: handles %d and %d share device %d", j, i, handles[i]->device);
    }
    ncclUniqueId uid;
    ncclResult_t r = api->GetUniqueId(&uid);
    if (r != ncclSuccess) return mcp_fail(handles[0], MCP_ERR_COMM, "ncclGetUniqueId failed: %s", api->GetErrorString(r));
    std::vector<mcp_comm_state*> cs((size_t)n, nullptr);
    for (int i = 0; i < n; ++i) {
        cs[i] = new (std::nothrow) mcp_comm_state();
        if (!cs[i]) {
            for (auto* c : cs) delete c;
            return mcp_fail(handles[0], MCP_ERR_NOMEM, "mcp_comm_init_all: out of host memory");
        }
    }
    int prev = -1;
    cudaGetDevice(&prev);
    r = api->GroupStart();
    for (int i = 0; i < n && r == ncclSuccess; ++i) {
        cudaSetDevice(handles[i]->device);
        r = api->CommInitRank(&cs[i]->comm, n, uid, i);
    }
    ncclResult_t r2 = api->GroupEnd();
    if (prev >= 0) cudaSetDevice(prev);
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) {
        for (auto* c : cs) {
            if (c->comm) api->CommAbort(c->comm);
            delete c;
        }
        return mcp_fail(handles[0], MCP_ERR_COMM, "ncclCommInitRank (group of %d) failed: %s", n, api->GetErrorString(r));
    }
    for (int i = 0; i < n; ++i) {
        handles[i]->comm = cs[i];
        handles[i]->comm_rank = i;
        handles[i]->comm_size = n;
    }
    return MCP_OK;
}

int mcp_comm_destroy(mcp_handle h) {
    if (!h) return MCP_ERR_INVALID;
    mcp_device_guard guard(h->device);
    cudaStreamSynchronize(h->stream);
    mcp_comm_release(h);
    return MCP_OK;
}

int mcp_comm_info(mcp_handle h, int* rank, int* nranks) {
    if (!h) return MCP_ERR_INVALID;
    if (rank) *rank = h->comm ? h->comm_rank : 0;
    if (nranks) *nranks = h->comm ? h->comm_size : 0;
    return MCP_OK;
}

int mcp_comm_nccl_version(int* version) {
    if (!version) return MCP_ERR_INVALID;
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!api) return mcp_fail(nullptr, MCP_ERR_COMM, "%s", why.c_str());
    return api->GetVersion(version) == ncclSuccess ? MCP_OK : MCP_ERR_COMM;
}

// Host-buffer collectives (tiny payloads: records, ranges, bins).  Staged through the handle's pinned + device scratch,
// stream-ordered on the handle's stream, one host wait, then the asynchronous-error check.
int mcp_comm_allgather(mcp_handle h, const void* send_host, size_t bytes, void* recv_host) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_comm_allgather", [&]() -> int {
        MCP_REQUIRE(h, h->comm != nullptr, "mcp_comm_allgather: no communicator on this handle");
        MCP_REQUIRE(h, send_host && recv_host && bytes > 0, "mcp_comm_allgather: bad arguments");
        mcp_device_guard guard(h->device);
        const size_t total = bytes * (size_t)h->comm_size;
        unsigned char* pin = nullptr;
        unsigned char* dev = nullptr;
        MCP_CHECK(mcp_pinned_reserve(h, 3, bytes + total, (void**)&pin));
        MCP_CHECK(mcp_dev_reserve(h, 14, bytes + total, (void**)&dev));
        memcpy(pin, send_host, bytes);
        cudaStream_t st = h->stream;
        MCP_CUDA(h, cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, st));
        MCP_CHECK(mcp_comm_allgather_dev(h, dev, dev + bytes, bytes, st));
        MCP_CUDA(h, cudaMemcpyAsync(pin + bytes, dev + bytes, total, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaStreamSynchronize(st));
        MCP_CHECK(mcp_comm_check(h));
        memcpy(recv_host, pin + bytes, total);
        return MCP_OK;
    });
}

int mcp_comm_allreduce(mcp_handle h, void* inout_host, size_t count, int kind) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_comm_allreduce", [&]() -> int {
        MCP_REQUIRE(h, h->comm != nullptr, "mcp_comm_allreduce: no communicator on this handle");
        MCP_REQUIRE(h, inout_host && count > 0, "mcp_comm_allreduce: bad arguments");
        mcp_device_guard guard(h->device);
        const size_t bytes = count * 8;
        unsigned char* pin = nullptr;
        unsigned char* dev = nullptr;
        MCP_CHECK(mcp_pinned_reserve(h, 3, bytes, (void**)&pin));
        MCP_CHECK(mcp_dev_reserve(h, 14, bytes, (void**)&dev));
        memcpy(pin, inout_host, bytes);
        cudaStream_t st = h->stream;
        MCP_CUDA(h, cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, st));
        MCP_CHECK(mcp_comm_allreduce_dev(h, dev, count, kind, st));
        MCP_CUDA(h, cudaMemcpyAsync(pin, dev, bytes, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaStreamSynchronize(st));
        MCP_CHECK(mcp_comm_check(h));
        memcpy(inout_host, pin, bytes);
        return MCP_OK;
    });
}

}  // extern "C"
