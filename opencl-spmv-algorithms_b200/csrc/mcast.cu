// mcast.cu -- NVSwitch multicast (multimem.*) for the per-step hand-over of the iterated mode.
//
// What a power-iteration step has to exchange besides the halo rows is tiny and goes to EVERYBODY: the 32
// partial sums of ||y||^2 and a "this rank is done" signal.  That is what NVLink multicast is for: a
// multicast object spans one 2 MiB block of every GPU's memory; a `multimem.red` to its address is
// carried out by the switch on EVERY GPU's copy (and a `multimem.st` stores to every copy), so
//     all-gather  = every rank stores its sum into ITS slot of the multicast block   (1 instruction)
//     barrier     = every rank adds 1 to a multicast counter and polls its OWN copy  (1 instruction)
//     all-reduce  = after the barrier every rank adds the slots in rank order (bit-identical everywhere)
// -- no ring, no tree, no peer loops, no NCCL call: one 32-thread kernel per step.  The halo rows stay
// unicast peer stores from the SpMV kernel's epilogue (each boundary row has exactly one reader, so
// there is nothing to multicast there).
//
// Set-up (driver API through cudaGetDriverEntryPoint: the library links neither libcuda nor NCCL):
//   rank 0: b200_mcast_create (cuMulticastCreate + a POSIX file descriptor to share)
//   others: b200_mcast_import_pid_fd (pidfd_getfd + cuMemImportFromShareableHandle) between processes,
//           b200_mcast_share between threads of one process
//   all   : b200_mcast_add_device, BARRIER, b200_mcast_bind (local 2 MiB + cuMulticastBindMem + the two
//           mappings), BARRIER, then use.
#include <cuda.h>
#include <errno.h>
#include <math.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "common.cuh"

namespace {

struct DriverApi {
    bool loaded = false;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                           unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
};
DriverApi g_drv;

template <typename F>
bool entry(const char *name, F &out)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        (void)cudaGetLastError();
        return false;
    }
    out = reinterpret_cast<F>(fn);
    return true;
}

int driver_api()
{
    if (g_drv.loaded) return B200_SUCCESS;
    DriverApi d;
    const bool ok = entry("cuMulticastCreate", d.MulticastCreate) && entry("cuMulticastAddDevice", d.MulticastAddDevice) &&
                    entry("cuMulticastBindMem", d.MulticastBindMem) && entry("cuMulticastUnbind", d.MulticastUnbind) &&
                    entry("cuMulticastGetGranularity", d.MulticastGetGranularity) && entry("cuMemCreate", d.MemCreate) &&
                    entry("cuMemRelease", d.MemRelease) && entry("cuMemAddressReserve", d.MemAddressReserve) &&
                    entry("cuMemAddressFree", d.MemAddressFree) && entry("cuMemMap", d.MemMap) &&
                    entry("cuMemUnmap", d.MemUnmap) && entry("cuMemSetAccess", d.MemSetAccess) &&
                    entry("cuMemExportToShareableHandle", d.MemExportToShareableHandle) &&
                    entry("cuMemImportFromShareableHandle", d.MemImportFromShareableHandle) &&
                    entry("cuDeviceGet", d.DeviceGet) && entry("cuDeviceGetAttribute", d.DeviceGetAttribute) &&
                    entry("cuGetErrorString", d.GetErrorString);
    if (!ok) {
        b200_set_error("this CUDA driver lacks the multicast / virtual-memory entry points");
        return B200_ERR_UNSUPPORTED;
    }
    d.loaded = true;
    g_drv = d;
    return B200_SUCCESS;
}

int cu_fail(CUresult r, const char *what, const char *file, int line)
{
    const char *s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    b200_set_error("CUDA driver error %d (%s) at %s:%d: %s", (int)r, s ? s : "?", file, line, what);
    return (r == CUDA_ERROR_NOT_SUPPORTED || r == CUDA_ERROR_NOT_PERMITTED) ? B200_ERR_UNSUPPORTED : B200_ERR_CUDA;
}
#define B200_CU(call)                                                           \
    do {                                                                        \
        CUresult r__ = (call);                                                  \
        if (r__ != CUDA_SUCCESS) return cu_fail(r__, #call, __FILE__, __LINE__); \
    } while (0)

// layout of the multicast block, in 8-byte words
constexpr int kMcFlag = 0;    // arrivals so far, summed over ranks and steps (multimem.red +1 per rank and step)
constexpr int kMcSlots = 16;  // [parity][16] one slot per rank: that rank's ||y_r||^2 of the step
constexpr int kMcWords = kMcSlots + 2 * 16;
constexpr unsigned long long kWaitLimitNs = 2000000000ull;

__device__ __forceinline__ unsigned long long timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One warp per rank and step, launched right after the fused SpMV kernel on the same queue (so all of that
// kernel's stores -- its own x block and the halo rows in the peers' memory -- have been performed):
//   1. fold this rank's 32 partial sums (shuffles, fixed order) and clear the accumulator,
//   2. lane 0: ONE multimem.st puts the sum into this rank's slot of the multicast block on every GPU, ONE
//      multimem.red (release) adds 1 to the multicast arrival counter.  Two multicast operations per rank
//      and step: a first version sent the 32 partial sums as 32 multimem.red.add.f64 and its cost grew
//      with the square of the rank count (every operation fans out to every GPU): 0.29 ms per step at 4
//      GPUs, 0.68 at 8, against 0.18 / 0.19 with NCCL's all-reduce;
//   3. poll the LOCAL copy of the counter until all `world` ranks of this step have arrived (acquire),
//   4. add the ranks' slots in rank order -- every GPU computes bit-identical sums, unlike reductions that
//      land in arrival order -- and hand the total to the next SpMV kernel (scale_out).
// Slots alternate with the step parity: a rank overwrites its slot of parity p only two steps later, after
// everybody has passed the barrier in between.  The step number lives in device memory so that the launch
// can be replayed from a graph.
__global__ void mcast_sync_kernel(double *__restrict__ acc_local, double *__restrict__ scale_out,
                                  unsigned long long *__restrict__ step_counter, unsigned long long *mc,
                                  unsigned long long *uc, int world, int my_rank, int *__restrict__ err_flag)
{
    const int lane = threadIdx.x;
    const unsigned long long step = *step_counter;
    const int p = (int)(step & 1);
    double v = acc_local[lane];
    acc_local[lane] = 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) {
        double *slot = reinterpret_cast<double *>(mc + kMcSlots) + p * 16 + my_rank;
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot), "d"(v) : "memory");
        asm volatile("multimem.red.release.sys.global.add.u64 [%0], %1;" ::"l"(mc + kMcFlag), "l"(1ull) : "memory");
    }
    const unsigned long long want = (unsigned long long)world * (step + 1);
    const unsigned long long t0 = timer_ns();
    for (;;) {
        unsigned long long f = 0;
        if (lane == 0) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(uc + kMcFlag) : "memory");
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f >= want) break;
        __nanosleep(32);
        if (timer_ns() - t0 > kWaitLimitNs) {  // a rank never arrived: flag it, do not hang
            if (lane == 0) atomicExch(err_flag, 3);
            break;
        }
    }
    if (lane == 0) {
        const double *slots = reinterpret_cast<const double *>(uc + kMcSlots) + p * 16;
        double total = 0.0;
        for (int r = 0; r < world; ++r) {
            double t;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(t) : "l"(slots + r) : "memory");
            total += t;
        }
        scale_out[0] = total;
        *step_counter = step + 1;
    } else {
        scale_out[lane] = 0.0;  // the consumer adds the 32 slots up
    }
}

}  // namespace

struct b200_mcast {
    b200_ctx *ctx;
    int world, rank;
    size_t bytes;  // rounded up to the multicast granularity
    CUmemGenericAllocationHandle mc_handle, mem_handle;
    CUdeviceptr mc_va, uc_va;
    bool have_mc, have_mem, added, bound;
    int export_fd;
    // per-rank local (ordinary device memory): accumulator, summed sums, step counter
    double *acc_local, *scale_out;
    unsigned long long *step_counter;
};

namespace {

int mcast_alloc(b200_ctx *ctx, int world, size_t bytes, b200_mcast **out)
{
    int rc = driver_api();
    if (rc) return rc;
    CUdevice dev;
    B200_CU(g_drv.DeviceGet(&dev, ctx->device));
    int supported = 0;
    B200_CU(g_drv.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
    if (!supported) {
        b200_set_error("device %d does not support NVLink multicast", ctx->device);
        return B200_ERR_UNSUPPORTED;
    }
    CUmulticastObjectProp prop;
    memset(&prop, 0, sizeof prop);
    prop.numDevices = (unsigned)world;
    prop.size = bytes;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    B200_CU(g_drv.MulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (gran == 0) gran = 2u << 20;
    b200_mcast *m = new b200_mcast();
    memset(m, 0, sizeof *m);
    m->ctx = ctx;
    m->world = world;
    m->bytes = (bytes + gran - 1) / gran * gran;
    m->export_fd = -1;
    *out = m;
    return B200_SUCCESS;
}

}  // namespace

extern "C" {

int b200_mcast_supported(b200_ctx *ctx, int *supported)
{
    B200_ENTER(ctx);
    B200_REQUIRE(supported, "null argument");
    *supported = 0;
    if (driver_api() != B200_SUCCESS) return B200_SUCCESS;
    CUdevice dev;
    int yes = 0;
    if (g_drv.DeviceGet(&dev, ctx->device) == CUDA_SUCCESS &&
        g_drv.DeviceGetAttribute(&yes, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS)
        *supported = yes;
    return B200_SUCCESS;
}

int b200_mcast_create(b200_ctx *ctx, int world, b200_mcast **mcast, int *export_fd)
{
    B200_ENTER(ctx);
    B200_REQUIRE(mcast && world >= 1, "bad argument");
    *mcast = nullptr;
    b200_mcast *m = nullptr;
    int rc = mcast_alloc(ctx, world, (size_t)kMcWords * 8, &m);
    if (rc) return rc;
    CUmulticastObjectProp prop;
    memset(&prop, 0, sizeof prop);
    prop.numDevices = (unsigned)world;
    prop.size = m->bytes;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUresult r = g_drv.MulticastCreate(&m->mc_handle, &prop);
    if (r != CUDA_SUCCESS) {
        delete m;
        return cu_fail(r, "cuMulticastCreate", __FILE__, __LINE__);
    }
    m->have_mc = true;
    if (export_fd) {
        int fd = -1;
        r = g_drv.MemExportToShareableHandle(&fd, m->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) {
            b200_mcast_destroy(m);
            return cu_fail(r, "cuMemExportToShareableHandle", __FILE__, __LINE__);
        }
        m->export_fd = fd;
        *export_fd = fd;
    }
    *mcast = m;
    return B200_SUCCESS;
}

int b200_mcast_import_fd(b200_ctx *ctx, int world, int fd, b200_mcast **mcast)
{
    B200_ENTER(ctx);
    B200_REQUIRE(mcast && world >= 1 && fd >= 0, "bad argument");
    *mcast = nullptr;
    b200_mcast *m = nullptr;
    int rc = mcast_alloc(ctx, world, (size_t)kMcWords * 8, &m);
    if (rc) return rc;
    CUresult r = g_drv.MemImportFromShareableHandle(&m->mc_handle, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) {
        delete m;
        return cu_fail(r, "cuMemImportFromShareableHandle", __FILE__, __LINE__);
    }
    m->have_mc = true;
    *mcast = m;
    return B200_SUCCESS;
}

int b200_mcast_import_pid_fd(b200_ctx *ctx, int world, int owner_pid, int owner_fd, b200_mcast **mcast)
{
    B200_REQUIRE(mcast && owner_pid > 0 && owner_fd >= 0, "bad argument");
    *mcast = nullptr;
#if defined(SYS_pidfd_open) && defined(SYS_pidfd_getfd)
    const int pidfd = (int)syscall(SYS_pidfd_open, owner_pid, 0);
    if (pidfd < 0) {
        b200_set_error("pidfd_open(%d) failed: errno %d", owner_pid, errno);
        return B200_ERR_UNSUPPORTED;
    }
    const int fd = (int)syscall(SYS_pidfd_getfd, pidfd, owner_fd, 0);
    const int err = errno;
    close(pidfd);
    if (fd < 0) {
        b200_set_error("pidfd_getfd(pid %d, fd %d) failed: errno %d (needs ptrace permission on the owner)", owner_pid, owner_fd, err);
        return B200_ERR_UNSUPPORTED;
    }
    const int rc = b200_mcast_import_fd(ctx, world, fd, mcast);
    close(fd);
    return rc;
#else
    b200_set_error("this libc has no pidfd_getfd: pass the descriptor yourself (b200_mcast_import_fd)");
    return B200_ERR_UNSUPPORTED;
#endif
}

int b200_mcast_share(b200_ctx *ctx, const b200_mcast *owner, b200_mcast **mcast)
{
    B200_ENTER(ctx);
    B200_REQUIRE(owner && owner->have_mc && mcast, "bad argument");
    *mcast = nullptr;
    b200_mcast *m = nullptr;
    int rc = mcast_alloc(ctx, owner->world, owner->bytes, &m);
    if (rc) return rc;
    m->mc_handle = owner->mc_handle;  // generic allocation handles are valid process-wide
    m->have_mc = false;               // ... and released by their owner only
    *mcast = m;
    return B200_SUCCESS;
}

int b200_mcast_add_device(b200_mcast *m)
{
    B200_REQUIRE(m, "null multicast block");
    B200_ENTER(m->ctx);
    CUdevice dev;
    B200_CU(g_drv.DeviceGet(&dev, m->ctx->device));
    B200_CU(g_drv.MulticastAddDevice(m->mc_handle, dev));
    m->added = true;
    return B200_SUCCESS;
}

int b200_mcast_bind(b200_mcast *m, int rank)
{
    B200_REQUIRE(m && m->added, "b200_mcast_add_device (on every rank, then a barrier) comes first");
    B200_REQUIRE(rank >= 0 && rank < m->world && m->world <= 16, "rank must be in [0, world), world at most 16");
    m->rank = rank;
    b200_ctx *ctx = m->ctx;
    B200_ENTER(ctx);
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = ctx->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    B200_CU(g_drv.MemCreate(&m->mem_handle, m->bytes, &prop, 0));
    m->have_mem = true;
    B200_CU(g_drv.MulticastBindMem(m->mc_handle, 0, m->mem_handle, 0, m->bytes, 0));
    m->bound = true;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = ctx->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    B200_CU(g_drv.MemAddressReserve(&m->uc_va, m->bytes, m->bytes, 0, 0));
    B200_CU(g_drv.MemMap(m->uc_va, m->bytes, 0, m->mem_handle, 0));
    B200_CU(g_drv.MemSetAccess(m->uc_va, m->bytes, &acc, 1));
    B200_CU(g_drv.MemAddressReserve(&m->mc_va, m->bytes, m->bytes, 0, 0));
    B200_CU(g_drv.MemMap(m->mc_va, m->bytes, 0, m->mc_handle, 0));
    B200_CU(g_drv.MemSetAccess(m->mc_va, m->bytes, &acc, 1));
    B200_CUDA(cudaMemsetAsync(reinterpret_cast<void *>(m->uc_va), 0, m->bytes, ctx->stream));
    B200_CUDA(cudaMalloc(&m->acc_local, sizeof(double) * 64 + 64));
    m->scale_out = m->acc_local + 32;
    m->step_counter = reinterpret_cast<unsigned long long *>(m->acc_local + 64);
    B200_CUDA(cudaMemsetAsync(m->acc_local, 0, sizeof(double) * 64 + 64, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_SUCCESS;
}

int b200_mcast_pointers(const b200_mcast *m, void **multicast_ptr, void **local_ptr, size_t *bytes)
{
    B200_REQUIRE(m && m->bound, "multicast block is not bound yet");
    if (multicast_ptr) *multicast_ptr = reinterpret_cast<void *>(m->mc_va);
    if (local_ptr) *local_ptr = reinterpret_cast<void *>(m->uc_va);
    if (bytes) *bytes = m->bytes;
    return B200_SUCCESS;
}

int b200_mcast_step_buffers(const b200_mcast *m, double **sumsq_accumulator, const double **summed_over_ranks)
{
    B200_REQUIRE(m && m->bound, "multicast block is not bound yet");
    if (sumsq_accumulator) *sumsq_accumulator = m->acc_local;
    if (summed_over_ranks) *summed_over_ranks = m->scale_out;
    return B200_SUCCESS;
}

int b200_mcast_allreduce_barrier(b200_mcast *m)
{
    B200_REQUIRE(m && m->bound, "multicast block is not bound yet");
    b200_ctx *ctx = m->ctx;
    B200_ENTER(ctx);
    mcast_sync_kernel<<<1, 32, 0, ctx->stream>>>(m->acc_local, m->scale_out, m->step_counter,
                                                 reinterpret_cast<unsigned long long *>(m->mc_va),
                                                 reinterpret_cast<unsigned long long *>(m->uc_va), m->world, m->rank,
                                                 ctx->scratch + kWatchFlag);
    B200_LAUNCH_CHECK();
    ctx->watch_flag = true;
    return B200_SUCCESS;
}

int b200_mcast_destroy(b200_mcast *m)
{
    if (!m) return B200_SUCCESS;
    b200_ctx *ctx = m->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (m->acc_local) cudaFree(m->acc_local);
    if (m->mc_va) {
        g_drv.MemUnmap(m->mc_va, m->bytes);
        g_drv.MemAddressFree(m->mc_va, m->bytes);
    }
    if (m->uc_va) {
        g_drv.MemUnmap(m->uc_va, m->bytes);
        g_drv.MemAddressFree(m->uc_va, m->bytes);
    }
    if (m->bound) {
        CUdevice dev;
        if (g_drv.DeviceGet(&dev, ctx->device) == CUDA_SUCCESS) g_drv.MulticastUnbind(m->mc_handle, dev, 0, m->bytes);
    }
    if (m->have_mem) g_drv.MemRelease(m->mem_handle);
    if (m->have_mc) g_drv.MemRelease(m->mc_handle);
    if (m->export_fd >= 0) close(m->export_fd);
    delete m;
    return B200_SUCCESS;
}

}  // extern "C"
