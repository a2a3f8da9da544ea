// iterate.cu -- the iterated (power-iteration) SpMV mode behind the C ABI: communicator + iterator.
//
// The reference stops at one device (it enumerates up to 8 GPUs, csr.c:12,22-30, and breaks out of
// the device loop after the first, csr.c:279), so everything here is new.  One rank = one context =
// one GPU; ranks may be processes (torchrun, peers mapped with b200_ipc_*) or threads of one
// process (the C drivers with --gpus N, peers enabled with b200_ctx_enable_peer_access).
//
//   b200_comm      an NCCL communicator bound to a context's stream.  libnccl.so.2 is opened at run
//                  time (dlopen), so the library itself has no NCCL link dependency and the
//                  single-GPU drivers run on boxes without it.
//   b200_iterator  `steps` steps of  y = A_r x / ||x||,  ||y||^2 all-reduced,  x <- y exchanged,
//                  in one of two formulations, issued directly or recorded ONCE into a CUDA graph
//                  (kernel + memset + NCCL nodes) and replayed -- a step lasts 0.1-0.2 ms on the
//                  device, less than a host needs to issue its three calls on 8 ranks in lock-step:
//       B200_ITER_FUSED      the SELL kernel stores each row straight into the x buffers of the ranks
//                            that read it (own block + halos, b200_spmv_sell_halo_f64) and
//                            accumulates ||y||^2; one 256-byte all-reduce per step is the only
//                            collective and doubles as the barrier that orders the peer stores.
//       B200_ITER_ALLGATHER  the formulation BASELINE.json names: SpMV (CSR or SELL) into the rank's
//                            segment, sum of squares, 1-element all-reduce, scale, in-place
//                            ncclAllGather of the segments.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
};

NcclApi g_nccl;

template <typename F>
bool load_sym(void *h, const char *name, F &out)
{
    out = reinterpret_cast<F>(dlsym(h, name));
    return out != nullptr;
}

// returns B200_SUCCESS once libnccl.so.2 is open and every entry point resolved
int nccl_api()
{
    if (g_nccl.handle) return B200_SUCCESS;
    // RTLD_NOLOAD first: inside a process that already carries NCCL (torch bundles its own copy under
    // the same soname) use THAT copy -- two NCCL instances in one process would each claim the GPUs
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        b200_set_error("multi-GPU mode needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
        return B200_ERR_UNSUPPORTED;
    }
    NcclApi a;
    a.handle = h;
    const bool ok = load_sym(h, "ncclGetVersion", a.GetVersion) && load_sym(h, "ncclGetUniqueId", a.GetUniqueId) &&
                    load_sym(h, "ncclCommInitRank", a.CommInitRank) && load_sym(h, "ncclCommDestroy", a.CommDestroy) &&
                    load_sym(h, "ncclCommAbort", a.CommAbort) &&
                    load_sym(h, "ncclCommGetAsyncError", a.CommGetAsyncError) &&
                    load_sym(h, "ncclGetErrorString", a.GetErrorString) && load_sym(h, "ncclAllReduce", a.AllReduce) &&
                    load_sym(h, "ncclAllGather", a.AllGather);
    if (!ok) {
        b200_set_error("libnccl.so.2 lacks an entry point this library needs: %s", dlerror());
        dlclose(h);
        return B200_ERR_UNSUPPORTED;
    }
    g_nccl = a;
    return B200_SUCCESS;
}

int nccl_fail(ncclResult_t r, const char *what, const char *file, int line)
{
    b200_set_error("NCCL error %d (%s) at %s:%d: %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?",
                   file, line, what);
    return B200_ERR_COMM;
}

#define B200_NCCL(call)                                                              \
    do {                                                                             \
        ncclResult_t r__ = (call);                                                   \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call, __FILE__, __LINE__);    \
    } while (0)

}  // namespace

struct b200_comm {
    b200_ctx *ctx;
    ncclComm_t comm;
    int rank, world;
};

struct b200_iterator {
    b200_ctx *ctx;
    b200_comm *comm;  // NULL when world == 1
    b200_iter_desc d;
    b200_block_f64 a;
    std::vector<double *> xs[2];  // x[b][r]
    std::vector<int> halo_lo, halo_hi;
    unsigned long long step;  // steps issued so far
    double *acc;              // device: FUSED 2 x B200_SUMSQ_SLOTS partial sums, ALLGATHER 2 x 1
    b200_graph *graph[2];     // d.graph_steps steps recorded from a step of parity p (NULL: not yet)
    b200_sell_plan *sell_plan;  // ALLGATHER on a SELL block: tells the SpMV whether the block is a stencil (pipelined kernel)
    unsigned long long launches;  // kernels + collectives issued (bench.py's gpu_launches)
};

namespace {

// one step, issued on the context's stream (or recorded, when a graph capture is open)
int issue_step(b200_iterator *it, unsigned long long k)
{
    b200_ctx *ctx = it->ctx;
    const b200_iter_desc &d = it->d;
    const b200_block_f64 &a = it->a;
    const int cur = (int)(k & 1), nxt = cur ^ 1;
    const long long offset = (long long)d.rank * d.rows_per_rank;
    int rc;
    if (d.mode == B200_ITER_FUSED_MCAST) {
        double *sums = nullptr;
        const double *summed = nullptr;
        rc = b200_mcast_step_buffers(d.mcast, &sums, &summed);
        if (rc) return rc;
        rc = b200_spmv_sell_halo_f64(ctx, a.data, a.indices, it->xs[cur][d.rank], a.ptr, 32, a.n_slices, a.n_rows,
                                     k > 0 ? summed : nullptr, sums, it->xs[nxt].data(), d.world, offset,
                                     it->halo_lo.empty() ? nullptr : it->halo_lo.data(),
                                     it->halo_hi.empty() ? nullptr : it->halo_hi.data());
        if (rc) return rc;
        rc = b200_mcast_allreduce_barrier(d.mcast);  // sums over ranks through the switch + the barrier
        if (rc) return rc;
        it->launches += 2;
        return B200_SUCCESS;
    }
    if (d.mode == B200_ITER_FUSED) {
        double *sums = it->acc + cur * B200_SUMSQ_SLOTS;
        const double *scale = k > 0 ? it->acc + nxt * B200_SUMSQ_SLOTS : nullptr;  // slots of step k-1
        rc = b200_memset_async(ctx, sums, 0, sizeof(double) * B200_SUMSQ_SLOTS);
        if (rc) return rc;
        rc = b200_spmv_sell_halo_f64(ctx, a.data, a.indices, it->xs[cur][d.rank], a.ptr, 32, a.n_slices, a.n_rows, scale,
                                     sums, it->xs[nxt].data(), d.world, offset,
                                     it->halo_lo.empty() ? nullptr : it->halo_lo.data(),
                                     it->halo_hi.empty() ? nullptr : it->halo_hi.data());
        if (rc) return rc;
        it->launches += 1;
        if (it->comm) {  // the norm AND the barrier that orders the peer stores
            B200_NCCL(g_nccl.AllReduce(sums, sums, B200_SUMSQ_SLOTS, ncclDouble, ncclSum, it->comm->comm, ctx->stream));
            it->launches += 1;
        }
        return B200_SUCCESS;
    }
    // B200_ITER_ALLGATHER
    double *x_cur = it->xs[cur][d.rank], *x_next = it->xs[nxt][d.rank];
    double *seg = x_next + offset;
    double *sum = it->acc + cur;
    if (a.format == B200_FORMAT_CSR)
        rc = b200_spmv_csr_f64(ctx, a.ptr, a.indices, a.data, x_cur, seg, a.n_rows, a.csr_plan);
    else
        rc = b200_spmv_sell_f64(ctx, a.data, a.indices, x_cur, seg, a.ptr, 32, a.n_slices, a.n_rows, nullptr, it->sell_plan);
    if (rc) return rc;
    rc = b200_memset_async(ctx, sum, 0, sizeof(double));
    if (rc) return rc;
    rc = b200_sumsq_f64(ctx, seg, a.n_rows, sum);
    if (rc) return rc;
    if (it->comm) B200_NCCL(g_nccl.AllReduce(sum, sum, 1, ncclDouble, ncclSum, it->comm->comm, ctx->stream));
    rc = b200_scale_f64(ctx, seg, a.n_rows, sum, 1);
    if (rc) return rc;
    it->launches += 3;
    if (it->comm) {
        B200_NCCL(g_nccl.AllGather(seg, x_next, (size_t)d.rows_per_rank, ncclDouble, it->comm->comm, ctx->stream));
        it->launches += 2;
    }
    return B200_SUCCESS;
}

}  // namespace

extern "C" {

int b200_comm_get_unique_id(unsigned char id[B200_COMM_ID_BYTES])
{
    B200_REQUIRE(id, "null id");
    static_assert(sizeof(ncclUniqueId) == B200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    int rc = nccl_api();
    if (rc) return rc;
    ncclUniqueId u;
    B200_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return B200_SUCCESS;
}

int b200_comm_create(b200_ctx *ctx, const unsigned char id[B200_COMM_ID_BYTES], int rank, int world, b200_comm **comm)
{
    B200_TRACE("b200 comm create");
    B200_ENTER(ctx);
    B200_REQUIRE(id && comm && world >= 1 && rank >= 0 && rank < world, "bad argument");
    *comm = nullptr;
    int rc = nccl_api();
    if (rc) return rc;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    B200_NCCL(g_nccl.CommInitRank(&c, world, u, rank));
    b200_comm *out = new b200_comm();
    out->ctx = ctx;
    out->comm = c;
    out->rank = rank;
    out->world = world;
    *comm = out;
    return B200_SUCCESS;
}

int b200_comm_destroy(b200_comm *comm)
{
    if (!comm) return B200_SUCCESS;
    cudaSetDevice(comm->ctx->device);
    cudaStreamSynchronize(comm->ctx->stream);
    ncclResult_t async = ncclSuccess;
    // a communicator that saw an asynchronous error cannot be destroyed collectively: abort it
    if (g_nccl.CommGetAsyncError(comm->comm, &async) != ncclSuccess || async != ncclSuccess)
        g_nccl.CommAbort(comm->comm);
    else
        g_nccl.CommDestroy(comm->comm);
    delete comm;
    return B200_SUCCESS;
}

int b200_comm_info(const b200_comm *comm, int *rank, int *world, int *nccl_version)
{
    B200_REQUIRE(comm, "null communicator");
    if (rank) *rank = comm->rank;
    if (world) *world = comm->world;
    if (nccl_version) B200_NCCL(g_nccl.GetVersion(nccl_version));
    return B200_SUCCESS;
}

int b200_comm_check(b200_comm *comm)
{
    B200_REQUIRE(comm, "null communicator");
    ncclResult_t async = ncclSuccess;
    B200_NCCL(g_nccl.CommGetAsyncError(comm->comm, &async));
    if (async != ncclSuccess && async != ncclInProgress) return nccl_fail(async, "asynchronous error", __FILE__, __LINE__);
    return B200_SUCCESS;
}

int b200_comm_allreduce_sum_f64(b200_comm *comm, double *buf, long long count)
{
    B200_TRACE("b200 allreduce");
    B200_REQUIRE(comm && buf && count >= 0, "bad argument");
    B200_ENTER(comm->ctx);
    if (count == 0) return B200_SUCCESS;
    B200_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, comm->comm, comm->ctx->stream));
    return B200_SUCCESS;
}

int b200_comm_allgather_f64(b200_comm *comm, double *full, long long count_per_rank)
{
    B200_TRACE("b200 allgather");
    B200_REQUIRE(comm && full && count_per_rank >= 0, "bad argument");
    B200_ENTER(comm->ctx);
    if (count_per_rank == 0) return B200_SUCCESS;
    B200_NCCL(g_nccl.AllGather(full + (long long)comm->rank * count_per_rank, full, (size_t)count_per_rank, ncclDouble,
                               comm->comm, comm->ctx->stream));
    return B200_SUCCESS;
}

int b200_comm_allgather_bytes(b200_comm *comm, void *full, long long bytes_per_rank)
{
    B200_TRACE("b200 allgather");
    B200_REQUIRE(comm && full && bytes_per_rank >= 0, "bad argument");
    B200_ENTER(comm->ctx);
    if (bytes_per_rank == 0) return B200_SUCCESS;
    B200_NCCL(g_nccl.AllGather(static_cast<char *>(full) + (long long)comm->rank * bytes_per_rank, full, (size_t)bytes_per_rank,
                               ncclChar, comm->comm, comm->ctx->stream));
    return B200_SUCCESS;
}

int b200_ctx_enable_peer_access(b200_ctx *ctx, int peer_device)
{
    B200_ENTER(ctx);
    if (peer_device == ctx->device) return B200_SUCCESS;
    int can = 0;
    B200_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) {
        b200_set_error("device %d cannot access device %d's memory", ctx->device, peer_device);
        return B200_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return B200_SUCCESS;
    }
    B200_CUDA(e);
    return B200_SUCCESS;
}

int b200_halo_rows(const int *col_min, const int *col_max, int world, int rank, long long rows_per_rank,
                   long long n_rows_total, int *lo, int *hi)
{
    B200_REQUIRE(col_min && col_max && lo && hi && world >= 1 && rank >= 0 && rank < world && rows_per_rank >= 0,
                 "bad argument");
    long long b0 = (long long)rank * rows_per_rank, b1 = b0 + rows_per_rank;
    if (b0 > n_rows_total) b0 = n_rows_total;
    if (b1 > n_rows_total) b1 = n_rows_total;
    for (int d = 0; d < world; ++d) {
        if (d == rank) {  // its own block in full: the next x of its own rows AND the result
            lo[d] = 0;
            hi[d] = (int)(b1 - b0);
            continue;
        }
        const long long first = col_min[d] > b0 ? col_min[d] : b0;
        const long long last = (long long)col_max[d] + 1 < b1 ? (long long)col_max[d] + 1 : b1;
        if (last <= first) {
            lo[d] = hi[d] = 0;
        } else {
            lo[d] = (int)(first - b0);
            hi[d] = (int)(last - b0);
        }
    }
    return B200_SUCCESS;
}

int b200_iterator_create(b200_ctx *ctx, b200_comm *comm, const b200_block_f64 *block, const b200_iter_desc *desc,
                         b200_iterator **iterator)
{
    B200_ENTER(ctx);
    B200_REQUIRE(block && desc && iterator, "null argument");
    *iterator = nullptr;
    const b200_iter_desc &d = *desc;
    B200_REQUIRE(d.mode == B200_ITER_FUSED || d.mode == B200_ITER_ALLGATHER || d.mode == B200_ITER_FUSED_MCAST, "unknown mode");
    B200_REQUIRE(d.world >= 1 && d.rank >= 0 && d.rank < d.world, "bad rank / world");
    if (d.mode == B200_ITER_FUSED_MCAST) {
        B200_REQUIRE(d.mcast != nullptr && comm == nullptr, "the multicast mode takes a bound b200_mcast block and no communicator");
    } else {
        B200_REQUIRE((d.world == 1) == (comm == nullptr), "give a communicator exactly when world > 1");
    }
    B200_REQUIRE(!comm || (comm->world == d.world && comm->rank == d.rank && comm->ctx == ctx),
                 "communicator belongs to another rank / world / context");
    B200_REQUIRE(d.rows_per_rank >= block->n_rows && d.rows_per_rank % 32 == 0, "rows_per_rank must be a multiple of 32 and >= n_rows");
    B200_REQUIRE(d.x[0] && d.x[1], "null x buffer tables");
    B200_REQUIRE(d.graph_steps >= 0 && d.graph_steps % 2 == 0, "graph_steps must be even (the x buffers alternate)");
    B200_REQUIRE(block->format == B200_FORMAT_CSR || block->format == B200_FORMAT_SELL, "block format must be CSR or SELL");
    B200_REQUIRE(d.mode == B200_ITER_ALLGATHER || block->format == B200_FORMAT_SELL, "the fused exchange is a SELL-32 kernel");
    B200_REQUIRE(d.mode == B200_ITER_ALLGATHER || d.world <= 16, "the fused exchange serves at most 16 ranks");
    B200_REQUIRE((d.halo_lo == nullptr) == (d.halo_hi == nullptr), "give both halo arrays or neither");
    b200_iterator *it = new b200_iterator();
    it->ctx = ctx;
    it->comm = comm;
    it->d = d;
    it->a = *block;
    it->step = 0;
    it->acc = nullptr;
    it->graph[0] = it->graph[1] = nullptr;
    it->sell_plan = nullptr;
    it->launches = 0;
    for (int b = 0; b < 2; ++b) {
        it->xs[b].resize(d.world, nullptr);
        for (int r = 0; r < d.world; ++r) {
            // the all-gather formulation touches only this rank's own buffers
            if (d.mode == B200_ITER_ALLGATHER && r != d.rank) continue;
            it->xs[b][r] = d.x[b][r];
            if (!it->xs[b][r]) {
                delete it;
                b200_set_error("b200_iterator_create: null x buffer (buffer %d, rank %d)", b, r);
                return B200_ERR_INVALID_VALUE;
            }
        }
    }
    if (d.halo_lo) {
        it->halo_lo.assign(d.halo_lo, d.halo_lo + d.world);
        it->halo_hi.assign(d.halo_hi, d.halo_hi + d.world);
    }
    it->d.x[0] = it->d.x[1] = nullptr;  // the caller's tables are not kept
    it->d.halo_lo = it->d.halo_hi = nullptr;
    cudaError_t e = cudaMalloc(&it->acc, sizeof(double) * 2 * B200_SUMSQ_SLOTS);
    if (e == cudaSuccess) e = cudaMemsetAsync(it->acc, 0, sizeof(double) * 2 * B200_SUMSQ_SLOTS, ctx->stream);
    if (e != cudaSuccess) {
        if (it->acc) cudaFree(it->acc);
        delete it;
        return b200_cuda_fail(e, "iterator scratch", __FILE__, __LINE__);
    }
    if (d.mode == B200_ITER_ALLGATHER && block->format == B200_FORMAT_SELL && block->n_slices > 0) {
        const int rc = b200_sell_plan_create(ctx, block->ptr, block->n_slices, &it->sell_plan);
        if (rc != B200_SUCCESS) {
            cudaFree(it->acc);
            delete it;
            return rc;
        }
    }
    *iterator = it;
    return B200_SUCCESS;
}

int b200_iterator_run(b200_iterator *it, int steps)
{
    B200_TRACE("b200 iterate");
    B200_REQUIRE(it && steps >= 0, "bad argument");
    b200_ctx *ctx = it->ctx;
    B200_ENTER(ctx);
    if (it->comm) {
        int rc = b200_comm_check(it->comm);
        if (rc) return rc;
    }
    const int G = it->d.graph_steps;
    while (steps > 0) {
        // step 0 has no previous norm to scale by and is also the first use of the communicator
        // (NCCL sets itself up lazily): always issued directly, never recorded
        if (G >= 2 && it->step >= 1 && steps >= G) {
            const int p = (int)(it->step & 1);
            if (!it->graph[p]) {
                const unsigned long long before = it->launches;
                int rc = b200_graph_begin(ctx);
                if (rc) return rc;
                for (int i = 0; i < G && rc == B200_SUCCESS; ++i) rc = issue_step(it, it->step + i);
                b200_graph *g = nullptr;
                const int rc_end = b200_graph_end(ctx, &g);  // always close the capture
                it->launches = before;
                if (rc || rc_end) {
                    if (g) b200_graph_destroy(g);
                    return rc ? rc : rc_end;
                }
                it->graph[p] = g;
            }
            int rc = b200_graph_launch(ctx, it->graph[p]);
            if (rc) return rc;
            const int per_step = it->d.mode == B200_ITER_FUSED_MCAST ? 2
                                 : it->d.mode == B200_ITER_FUSED     ? (it->comm ? 2 : 1)
                                                                     : (it->comm ? 5 : 3);
            it->launches += (unsigned long long)per_step * G;
            it->step += G;
            steps -= G;
        } else {
            int rc = issue_step(it, it->step);
            if (rc) return rc;
            it->step += 1;
            steps -= 1;
        }
    }
    return B200_SUCCESS;
}

int b200_iterator_norm(b200_iterator *it, double *norm)
{
    B200_REQUIRE(it && norm, "null argument");
    b200_ctx *ctx = it->ctx;
    B200_ENTER(ctx);
    *norm = NAN;
    if (it->step == 0) return B200_SUCCESS;
    const int last = (int)((it->step - 1) & 1);
    double host[B200_SUMSQ_SLOTS];
    const int n = it->d.mode == B200_ITER_ALLGATHER ? 1 : B200_SUMSQ_SLOTS;
    const double *src = it->d.mode == B200_ITER_FUSED ? it->acc + last * B200_SUMSQ_SLOTS : it->acc + last;
    if (it->d.mode == B200_ITER_FUSED_MCAST) {
        int rc0 = b200_mcast_step_buffers(it->d.mcast, nullptr, &src);  // the sums over ranks of the last step
        if (rc0) return rc0;
    }
    int rc = b200_memcpy_d2h(ctx, host, src, sizeof(double) * n);
    if (rc) return rc;
    if (it->comm) {
        rc = b200_comm_check(it->comm);
        if (rc) return rc;
    }
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += host[i];
    *norm = sqrt(s);
    return B200_SUCCESS;
}

int b200_iterator_state(const b200_iterator *it, unsigned long long *steps_done, double **x_current,
                        unsigned long long *launches)
{
    B200_REQUIRE(it, "null iterator");
    if (steps_done) *steps_done = it->step;
    if (x_current) *x_current = it->xs[it->step & 1][it->d.rank];
    if (launches) *launches = it->launches;
    return B200_SUCCESS;
}

int b200_iterator_destroy(b200_iterator *it)
{
    if (!it) return B200_SUCCESS;
    cudaSetDevice(it->ctx->device);
    cudaStreamSynchronize(it->ctx->stream);
    for (int p = 0; p < 2; ++p)
        if (it->graph[p]) b200_graph_destroy(it->graph[p]);
    if (it->acc) cudaFree(it->acc);
    if (it->sell_plan) b200_sell_plan_destroy(it->sell_plan);
    delete it;
    return B200_SUCCESS;
}

}  // extern "C"
