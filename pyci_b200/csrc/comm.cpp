// NCCL plumbing for the row-sharded operator: one communicator per context (one process per GPU of
// one box), used for the per-iteration all-gather of the trial vector and the small all-reduces of
// the Davidson solver and the RDM tensors.  libnccl.so.2 is resolved lazily with dlopen so that the
// single-GPU path has no NCCL dependency and so that, inside a torch process, the copy of NCCL that
// torch already loaded is the one that is used.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle)
                break;
        }
        if (!a.handle) {
            a.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
            return;
        }
#define LOAD(sym)                                                                                  \
    a.sym = reinterpret_cast<decltype(a.sym)>(dlsym(a.handle, "nccl" #sym));                       \
    if (!a.sym) {                                                                                  \
        a.why = "libnccl lacks nccl" #sym;                                                         \
        return;                                                                                    \
    }
        LOAD(GetUniqueId)
        LOAD(CommInitRank)
        LOAD(CommDestroy)
        LOAD(AllGather)
        LOAD(AllReduce)
        LOAD(Broadcast)
        LOAD(Send)
        LOAD(Recv)
        LOAD(GroupStart)
        LOAD(GroupEnd)
        LOAD(GetErrorString)
#undef LOAD
        a.ok = true;
    });
    return a;
}

#define PYCI_NCCL(expr)                                                                            \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            pyci_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, api().GetErrorString(_r)); \
            return PYCI_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

int need_api() {
    if (!api().ok)
        PYCI_FAIL(PYCI_ERR_CUDA, "%s", api().why.c_str());
    return PYCI_OK;
}

} // namespace

int comm_unique_id(void *out128) {
    PYCI_TRY(need_api());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    PYCI_NCCL(api().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return PYCI_OK;
}

int comm_init(pyci_ctx *ctx, int rank, int nranks, const void *id128) {
    PYCI_TRY(need_api());
    if (!id128)
        PYCI_FAIL(PYCI_ERR_VALUE, "null NCCL unique id");
    if (ctx->comm)
        comm_destroy(ctx);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    PYCI_NCCL(api().CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return PYCI_OK;
}

void comm_destroy(pyci_ctx *ctx) {
    if (ctx->comm && api().ok)
        api().CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
}

int comm_allgather_f64(pyci_ctx *ctx, const double *send_dev, double *recv_dev, long count_per_rank) {
    if (ctx->nranks == 1) {
        if (send_dev != recv_dev)
            PYCI_CUDA(cudaMemcpyAsync(recv_dev, send_dev, sizeof(double) * count_per_rank, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
        return PYCI_OK;
    }
    PYCI_NCCL(api().AllGather(send_dev, recv_dev, (size_t)count_per_rank, ncclDouble, (ncclComm_t)ctx->comm,
                              ctx->stream));
    return PYCI_OK;
}

// All-gather of unequal shards: rank p contributes bounds[p+1] - bounds[p] elements, which land at recv + bounds[p] on
// every rank (the nnz-balanced row partition of a selected space).  One NCCL group of broadcasts, one per rank.
int comm_allgatherv_f64(pyci_ctx *ctx, const double *send_dev, double *recv_dev, const long *bounds) {
    const int R = ctx->nranks, me = ctx->rank;
    if (R == 1) {
        if (send_dev != recv_dev && bounds[1] > bounds[0])
            PYCI_CUDA(cudaMemcpyAsync(recv_dev + bounds[0], send_dev, sizeof(double) * (size_t)(bounds[1] - bounds[0]),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
        return PYCI_OK;
    }
    PYCI_NCCL(api().GroupStart());
    ncclResult_t r = ncclSuccess;
    for (int p = 0; p < R && r == ncclSuccess; ++p) {
        const long cnt = bounds[p + 1] - bounds[p];
        if (cnt <= 0)
            continue;
        r = api().Broadcast(p == me ? (const void *)send_dev : (const void *)(recv_dev + bounds[p]), recv_dev + bounds[p],
                            (size_t)cnt, ncclDouble, p, (ncclComm_t)ctx->comm, ctx->stream);
    }
    const ncclResult_t g = api().GroupEnd();
    PYCI_NCCL(r);
    PYCI_NCCL(g);
    return PYCI_OK;
}

int comm_allreduce_sum_f64(pyci_ctx *ctx, double *buf_dev, long count) {
    if (ctx->nranks == 1)
        return PYCI_OK;
    PYCI_NCCL(api().AllReduce(buf_dev, buf_dev, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->comm,
                              ctx->stream));
    return PYCI_OK;
}

int comm_allreduce_sum_i64_host(pyci_ctx *ctx, long *vals, int count) {
    if (ctx->nranks == 1)
        return PYCI_OK;
    long *d = nullptr;
    PYCI_CUDA(dev_malloc(&d, sizeof(long) * count)); // stream-ordered pool: no device-wide synchronisation
    cudaError_t e = cudaMemcpyAsync(d, vals, sizeof(long) * count, cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess)
        r = api().AllReduce(d, d, (size_t)count, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
    if (e == cudaSuccess && r == ncclSuccess) {
        e = cudaMemcpyAsync(vals, d, sizeof(long) * count, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(ctx->stream);
    }
    dev_free(d);
    PYCI_CUDA(e);
    PYCI_NCCL(r);
    return PYCI_OK;
}

// Personalised exchange of 64-bit words: rank p receives send[soff[p] .. soff[p] + scount[p]) of every rank into
// recv[roff[q] ..) (q = sender).  One NCCL group of point-to-point sends / receives; the local part is a copy.
int comm_alltoallv_u64(pyci_ctx *ctx, const unsigned long long *send, const long *scount, const long *soff,
                       unsigned long long *recv, const long *rcount, const long *roff) {
    const int R = ctx->nranks, me = ctx->rank;
    if (scount[me] > 0)
        PYCI_CUDA(cudaMemcpyAsync(recv + roff[me], send + soff[me], sizeof(unsigned long long) * (size_t)scount[me],
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    if (R == 1)
        return PYCI_OK;
    PYCI_NCCL(api().GroupStart());
    ncclResult_t r = ncclSuccess;
    for (int p = 0; p < R && r == ncclSuccess; ++p) {
        if (p == me)
            continue;
        if (scount[p] > 0)
            r = api().Send(send + soff[p], (size_t)scount[p], ncclUint64, p, (ncclComm_t)ctx->comm, ctx->stream);
        if (r == ncclSuccess && rcount[p] > 0)
            r = api().Recv(recv + roff[p], (size_t)rcount[p], ncclUint64, p, (ncclComm_t)ctx->comm, ctx->stream);
    }
    const ncclResult_t g = api().GroupEnd();
    PYCI_NCCL(r);
    PYCI_NCCL(g);
    return PYCI_OK;
}

// The same exchange in bytes (arrays of 4-byte columns, 8-byte values, ...): counts and offsets are byte counts.
int comm_alltoallv_bytes(pyci_ctx *ctx, const void *send, const long *scount, const long *soff, void *recv,
                         const long *rcount, const long *roff) {
    const int R = ctx->nranks, me = ctx->rank;
    const char *s = static_cast<const char *>(send);
    char *d = static_cast<char *>(recv);
    if (scount[me] > 0)
        PYCI_CUDA(cudaMemcpyAsync(d + roff[me], s + soff[me], (size_t)scount[me], cudaMemcpyDeviceToDevice, ctx->stream));
    if (R == 1)
        return PYCI_OK;
    PYCI_NCCL(api().GroupStart());
    ncclResult_t r = ncclSuccess;
    for (int p = 0; p < R && r == ncclSuccess; ++p) {
        if (p == me)
            continue;
        if (scount[p] > 0)
            r = api().Send(s + soff[p], (size_t)scount[p], ncclChar, p, (ncclComm_t)ctx->comm, ctx->stream);
        if (r == ncclSuccess && rcount[p] > 0)
            r = api().Recv(d + roff[p], (size_t)rcount[p], ncclChar, p, (ncclComm_t)ctx->comm, ctx->stream);
    }
    const ncclResult_t g = api().GroupEnd();
    PYCI_NCCL(r);
    PYCI_NCCL(g);
    return PYCI_OK;
}
