// comm.cu -- run-time binding to NCCL and the asb_comm entry points of the C ABI (see comm.cuh).
#include <dlfcn.h>

#include <mutex>

#include "comm.cuh"

namespace {

// the slice of nccl.h this library uses (types restated so that no NCCL header is needed at build time)
typedef void *nccl_comm_t;
typedef struct {
    char internal[128];
} nccl_unique_id;
enum { NCCL_INT8 = 0, NCCL_INT64 = 4, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*CommCount)(nccl_comm_t, int *) = nullptr;
    int (*CommUserRank)(nccl_comm_t, int *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        bool all = true;
#define ASB_NCCL_SYM(field, sym)                                            \
    do {                                                                    \
        *(void **)(&api.field) = dlsym(api.lib, sym);                       \
        if (!api.field) {                                                   \
            all = false;                                                    \
            api.error = std::string("NCCL symbol missing: ") + sym;         \
        }                                                                   \
    } while (0)
        ASB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        ASB_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        ASB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        ASB_NCCL_SYM(CommCount, "ncclCommCount");
        ASB_NCCL_SYM(CommUserRank, "ncclCommUserRank");
        ASB_NCCL_SYM(AllReduce, "ncclAllReduce");
        ASB_NCCL_SYM(AllGather, "ncclAllGather");
        ASB_NCCL_SYM(Broadcast, "ncclBroadcast");
        ASB_NCCL_SYM(Send, "ncclSend");
        ASB_NCCL_SYM(Recv, "ncclRecv");
        ASB_NCCL_SYM(GroupStart, "ncclGroupStart");
        ASB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        ASB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef ASB_NCCL_SYM
        api.ok = all;
    });
    return api;
}

int nccl_fail(asb_ctx *ctx, int rc, const char *what) {
    NcclApi &a = nccl();
    char b[512];
    snprintf(b, sizeof(b), "NCCL error in %s: %s", what, (a.ok && a.GetErrorString) ? a.GetErrorString(rc) : "?");
    ctx->last_error = b;
    return ASB_ERR_NCCL;
}

#define ASB_NCCL(ctx, expr, what)                        \
    do {                                                 \
        int _rc = (expr);                                \
        if (_rc != 0) return nccl_fail(ctx, _rc, what);  \
    } while (0)

int need_nccl(asb_ctx *ctx) {
    NcclApi &a = nccl();
    if (!a.ok) {
        ctx->last_error = a.error.empty() ? "NCCL unavailable" : a.error;
        return ASB_ERR_NCCL;
    }
    return ASB_OK;
}

}  // namespace

bool asb_nccl_available(std::string *why) {
    NcclApi &a = nccl();
    if (!a.ok && why) *why = a.error;
    return a.ok;
}

int asb_comm_bcast_bytes(asb_ctx *ctx, asb_comm *comm, void *buf_d, size_t bytes, int root) {
    if (!comm || comm->nranks == 1 || bytes == 0) return ASB_OK;
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().Broadcast(buf_d, buf_d, bytes, NCCL_INT8, root, comm->nccl, ctx->stream), "ncclBroadcast");
    return ASB_OK;
}
int asb_comm_allreduce_f64(asb_ctx *ctx, asb_comm *comm, double *buf_d, size_t count, AsbRedOp op) {
    if (!comm || comm->nranks == 1 || count == 0) return ASB_OK;
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().AllReduce(buf_d, buf_d, count, NCCL_FLOAT64, (int)op, comm->nccl, ctx->stream), "ncclAllReduce");
    return ASB_OK;
}
int asb_comm_allreduce_i64(asb_ctx *ctx, asb_comm *comm, long long *buf_d, size_t count, AsbRedOp op) {
    if (!comm || comm->nranks == 1 || count == 0) return ASB_OK;
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().AllReduce(buf_d, buf_d, count, NCCL_INT64, (int)op, comm->nccl, ctx->stream), "ncclAllReduce");
    return ASB_OK;
}
int asb_comm_allgather_bytes(asb_ctx *ctx, asb_comm *comm, const void *send_d, void *recv_d, size_t bytes_per_rank) {
    if (!comm || comm->nranks == 1) {
        if (send_d != recv_d && bytes_per_rank)
            ASB_CUDA(ctx, cudaMemcpyAsync(recv_d, send_d, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return ASB_OK;
    }
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().AllGather(send_d, recv_d, bytes_per_rank, NCCL_INT8, comm->nccl, ctx->stream), "ncclAllGather");
    return ASB_OK;
}
int asb_comm_send_bytes(asb_ctx *ctx, asb_comm *comm, const void *buf_d, size_t bytes, int peer) {
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().Send(buf_d, bytes, NCCL_INT8, peer, comm->nccl, ctx->stream), "ncclSend");
    return ASB_OK;
}
int asb_comm_recv_bytes(asb_ctx *ctx, asb_comm *comm, void *buf_d, size_t bytes, int peer) {
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().Recv(buf_d, bytes, NCCL_INT8, peer, comm->nccl, ctx->stream), "ncclRecv");
    return ASB_OK;
}
int asb_comm_group_start(asb_ctx *ctx) {
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().GroupStart(), "ncclGroupStart");
    return ASB_OK;
}
int asb_comm_group_end(asb_ctx *ctx) {
    ASB_TRY(need_nccl(ctx));
    ASB_NCCL(ctx, nccl().GroupEnd(), "ncclGroupEnd");
    return ASB_OK;
}

extern "C" {

int asb_comm_unique_id(asb_ctx *ctx, void *id_out) {
    if (!ctx || !id_out) return ASB_ERR_INVALID;
    ASB_TRY(need_nccl(ctx));
    nccl_unique_id id;
    ASB_NCCL(ctx, nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof(id));
    return ASB_OK;
}

int asb_comm_init_rank(asb_ctx *ctx, const void *unique_id, int nranks, int rank, asb_comm **out) {
    if (!ctx || !unique_id || !out || nranks < 1 || rank < 0 || rank >= nranks) return ASB_ERR_INVALID;
    *out = nullptr;
    ASB_CUDA(ctx, cudaSetDevice(ctx->device));
    asb_comm *c = new asb_comm();
    c->rank = rank;
    c->nranks = nranks;
    if (nranks > 1) {
        int rc0 = need_nccl(ctx);
        if (rc0 != ASB_OK) {
            delete c;
            return rc0;
        }
        nccl_unique_id id;
        memcpy(&id, unique_id, sizeof(id));
        nccl_comm_t comm = nullptr;
        int rc = nccl().CommInitRank(&comm, nranks, id, rank);
        if (rc != 0) {
            delete c;
            return nccl_fail(ctx, rc, "ncclCommInitRank");
        }
        c->nccl = comm;
        c->owned = true;
    }
    *out = c;
    return ASB_OK;
}

int asb_comm_from_nccl(asb_ctx *ctx, void *nccl_comm, asb_comm **out) {
    if (!ctx || !nccl_comm || !out) return ASB_ERR_INVALID;
    *out = nullptr;
    ASB_TRY(need_nccl(ctx));
    asb_comm *c = new asb_comm();
    c->nccl = nccl_comm;
    c->owned = false;
    int rc = nccl().CommCount(nccl_comm, &c->nranks);
    if (rc == 0) rc = nccl().CommUserRank(nccl_comm, &c->rank);
    if (rc != 0) {
        delete c;
        return nccl_fail(ctx, rc, "ncclCommCount / ncclCommUserRank");
    }
    *out = c;
    return ASB_OK;
}

void asb_comm_destroy(asb_comm *comm) {
    if (!comm) return;
    if (comm->owned && comm->nccl && nccl().ok) nccl().CommDestroy(comm->nccl);
    delete comm;
}

int asb_comm_rank(const asb_comm *comm) { return comm ? comm->rank : 0; }
int asb_comm_size(const asb_comm *comm) { return comm ? comm->nranks : 1; }

}  // extern "C"
