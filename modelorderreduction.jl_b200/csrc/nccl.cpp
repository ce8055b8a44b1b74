// NCCL plumbing for the row-sharded (one process per GPU) mode.
//
// NCCL is loaded with dlopen on first use so that (a) a single-GPU user needs no NCCL at
// all and (b) inside a torch process the library binds to the libnccl.so.2 torch already
// mapped instead of dragging in a second copy.  Only the exchange steps of SURVEY.md
// section 8e go through here: the n x n Gram all-reduce of the POD passes and the
// (value, index, row) all-gather of every DEIM step.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "internal.h"

namespace mor {

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int32_t load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return MOR_OK;
    const char* override_path = getenv("MOR_B200_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
        return MOR_ERR_CUDA;
    }
#define LOAD(field, sym)                                             \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, sym)); \
    if (!g_nccl.field) {                                             \
        set_error(std::string("NCCL symbol missing: ") + sym);       \
        dlclose(h);                                                  \
        return MOR_ERR_CUDA;                                         \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.handle = h;
    return MOR_OK;
}

int32_t nccl_fail(ncclResult_t r, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "NCCL error %d (%s) in %s", (int)r,
             g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    set_error(buf);
    return MOR_ERR_CUDA;
}
}  // namespace

int32_t comm_allreduce_sum(double* buf, size_t count) {
    Ctx& c = ctx();
    if (c.nranks <= 1 || count == 0) return MOR_OK;
    ncclResult_t r =
        g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)c.comm, c.stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    return MOR_OK;
}

int32_t comm_allgather(const void* send, void* recv, size_t bytes_per_rank) {
    Ctx& c = ctx();
    if (c.nranks <= 1) {
        if (send != recv) MOR_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, c.stream));
        return MOR_OK;
    }
    ncclResult_t r =
        g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)c.comm, c.stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    return MOR_OK;
}

int32_t comm_allgather_i64(int64_t mine, std::vector<int64_t>& all) {
    Ctx& c = ctx();
    all.assign((size_t)c.nranks, 0);
    if (c.nranks <= 1) {
        all[0] = mine;
        return MOR_OK;
    }
    DevBuf d;
    MOR_TRY(d.alloc(sizeof(int64_t) * (size_t)(c.nranks + 1)));
    int64_t* dp = d.as<int64_t>();
    MOR_CUDA(cudaMemcpyAsync(dp + c.nranks, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
    MOR_TRY(comm_allgather(dp + c.nranks, dp, sizeof(int64_t)));
    MOR_CUDA(cudaMemcpyAsync(all.data(), dp, sizeof(int64_t) * (size_t)c.nranks, cudaMemcpyDeviceToHost,
                             c.stream));
    MOR_CUDA(cudaStreamSynchronize(c.stream));
    return MOR_OK;
}

// ---- DEIM mailbox: the per-step candidate exchange goes over peer memory, not through NCCL ------------------------
static void mailbox_set_pointers(Ctx& c) {
    unsigned char* b = static_cast<unsigned char*>(c.box_base);
    c.box_ticket = reinterpret_cast<unsigned*>(b + 512);
    c.box = reinterpret_cast<double*>(b + MBOX_HEADER_BYTES);
    c.box_ll = reinterpret_cast<uint4*>(b + MBOX_HEADER_BYTES + MBOX_PLAIN_BYTES);
    c.peer_base[c.rank] = c.box_base;
    c.peer_ll[c.rank] = c.box_ll;
}

void deim_mailbox_release() {
    Ctx& c = ctx();
    for (int r = 0; r < MBOX_MAXRANKS; ++r) {
        if (c.peers_ipc && c.peer_base[r] && c.peer_base[r] != c.box_base) cudaIpcCloseMemHandle(c.peer_base[r]);
        c.peer_base[r] = nullptr;
        c.peer_ll[r] = nullptr;
    }
    if (c.box_base) cudaFree(c.box_base);
    c.box_base = nullptr;
    c.box = nullptr;
    c.box_ll = nullptr;
    c.box_ticket = nullptr;
    c.peers_tried = c.peers_ok = c.peers_ipc = false;
    c.deim_seq = 0;
    cudaGetLastError();
}

int32_t deim_mailbox() {
    Ctx& c = ctx();
    if (!c.box_base) {
        MOR_CUDA(cudaMalloc(&c.box_base, MBOX_BYTES));
        MOR_CUDA(cudaMemset(c.box_base, 0, MBOX_BYTES));
        MOR_CUDA(cudaDeviceSynchronize());      // zeroed before any peer can learn the address
        mailbox_set_pointers(c);
    }
    if (c.nranks <= 1 || c.peers_tried) return MOR_OK;
    c.peers_tried = true;
    // One process per GPU: export the mailbox with CUDA IPC, all-gather the handles (the one NCCL call of the setup),
    // map every peer's mailbox.  All ranks or none: otherwise the exchange stays on ncclAllGather.
    const char* ex = getenv("MOR_B200_DEIM_EXCHANGE");
    bool ok = !(ex && strcmp(ex, "nccl") == 0);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine, c.box_base) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
    }
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    std::vector<cudaIpcMemHandle_t> all((size_t)c.nranks);
    {
        DevBuf d;
        MOR_TRY(d.alloc(hb * (size_t)(c.nranks + 1)));
        unsigned char* dp = d.as<unsigned char>();
        MOR_CUDA(cudaMemcpyAsync(dp + hb * (size_t)c.nranks, &mine, hb, cudaMemcpyHostToDevice, c.stream));
        MOR_TRY(comm_allgather(dp + hb * (size_t)c.nranks, dp, hb));
        MOR_CUDA(cudaMemcpyAsync(all.data(), dp, hb * (size_t)c.nranks, cudaMemcpyDeviceToHost, c.stream));
        MOR_CUDA(cudaStreamSynchronize(c.stream));
    }
    void* mapped[MBOX_MAXRANKS] = {nullptr};
    for (int r = 0; r < c.nranks && ok; ++r) {
        if (r == c.rank) continue;
        if (cudaIpcOpenMemHandle(&mapped[r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            mapped[r] = nullptr;
            ok = false;
        }
    }
    std::vector<int64_t> votes;
    MOR_TRY(comm_allgather_i64(ok ? 1 : 0, votes));
    for (int64_t v : votes) ok = ok && v != 0;
    if (!ok) {
        for (int r = 0; r < c.nranks; ++r)
            if (mapped[r]) cudaIpcCloseMemHandle(mapped[r]);
        cudaGetLastError();
        return MOR_OK;
    }
    for (int r = 0; r < c.nranks; ++r) {
        if (r == c.rank) continue;
        unsigned char* b = static_cast<unsigned char*>(mapped[r]);
        c.peer_base[r] = mapped[r];
        c.peer_ll[r] = reinterpret_cast<uint4*>(b + MBOX_HEADER_BYTES + MBOX_PLAIN_BYTES);
    }
    c.peers_ipc = true;
    c.peers_ok = true;
    return MOR_OK;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_comm_unique_id(void* id128) {
    MOR_API_LOCK;
    MOR_ARG(id128 != nullptr, "mor_b200_comm_unique_id: null buffer");
    MOR_TRY(load_nccl());
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return MOR_OK;
}

int32_t mor_b200_comm_init(int32_t rank, int32_t nranks, const void* id128) {
    MOR_API_LOCK;
    MOR_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "mor_b200_comm_init: bad rank/nranks");
    if (multi_on() && !is_worker()) {
        set_error("mor_b200_comm_init: the library is in single-process multi-GPU mode; it owns the communicator");
        return MOR_ERR_ARG;
    }
    MOR_TRY(require_ctx());
    Ctx& c = ctx();
    if (c.comm) mor_b200_comm_destroy();
    deim_mailbox_release();   // flags / step counter restart with the communicator
    if (nranks == 1) {
        c.rank = 0;
        c.nranks = 1;
        return MOR_OK;
    }
    MOR_ARG(id128 != nullptr, "mor_b200_comm_init: null unique id");
    MOR_TRY(load_nccl());
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    c.comm = comm;
    c.rank = rank;
    c.nranks = nranks;
    return MOR_OK;
}

int32_t mor_b200_comm_info(int32_t* rank, int32_t* nranks) {
    if (rank) *rank = ctx().rank;
    if (nranks) *nranks = ctx().nranks;
    return MOR_OK;
}

int32_t mor_b200_comm_destroy(void) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return MOR_OK;   // the workers' communicators go with mor_b200_shutdown
    Ctx& c = ctx();
    if (c.comm && g_nccl.CommDestroy) {
        if (c.ready) {
            cudaSetDevice(c.device);
            cudaStreamSynchronize(c.stream);
        }
        g_nccl.CommDestroy((ncclComm_t)c.comm);
    }
    if (c.ready) cudaSetDevice(c.device);
    deim_mailbox_release();
    c.comm = nullptr;
    c.rank = 0;
    c.nranks = 1;
    return MOR_OK;
}

}  // extern "C"
