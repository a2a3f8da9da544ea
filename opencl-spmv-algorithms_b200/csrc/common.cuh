// common.cuh -- shared device/host helpers of libb200spmv (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/b200spmv.h"

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "libb200spmv is written for sm_100a (B200) only"
#endif

// Tuning hooks.  Every hook has an environment variable of the same name; the environment is read
// ONCE, in b200_ctx_create, into ctx->opt[] (no getenv in front of a 10 us launch), and
// b200_ctx_set_option changes a hook on a live context.  kOptUnset = the library decides.
enum b200_opt {
    OPT_CSR_LANES, OPT_ELL_LANES, OPT_CSR_UNROLL, OPT_ELL_UNROLL, OPT_CSR_STREAM, OPT_SELL_WPC,
    OPT_SELL_UNROLL, OPT_SELL_TMA, OPT_SELL_TMA_BLOCKS, OPT_SELL_TMA_SUSPEND_NS, OPT_COO_U, OPT_CMRS_U,
    OPT_CMRS_WPS, OPT_CMRS_STREAM, OPT_ELLCM_Q, OPT_RING_FLUSH, OPT_RING_POLL, OPT_RING_SLEEP_NS,
    OPT_BCAST_U, OPT_CSR_STREAM_G, OPT_SELL_PIPE, OPT_COUNT
};
constexpr int kOptUnset = INT_MIN;
extern const char *const b200_opt_names[OPT_COUNT];

struct b200_ctx {
    int device;
    int opt[OPT_COUNT];    // tuning hooks (kOptUnset = automatic)
    cudaStream_t stream;
    bool owns_stream;
    int sm_count;
    int l2_bytes;
    int max_persist_l2;
    int *scratch;          // small device scratch (flags / counters), 4 KiB
    void *host_scratch;    // pinned, 4 KiB, for small readbacks
    bool watch_flag;       // a bulk-copy kernel ran since the last sync: b200_sync reads kWatchFlag
    bool watch_saved;      // watch_flag as it was when b200_graph_begin started recording
    bool persist_set;      // b200_ctx_set_l2_persist reserved a (device-wide) persisting-L2 carve-out
    bool overlap;          // b200_ctx_set_launch_overlap: SpMV kernels are launched as programmatic
                           // dependents (they stream their matrix arrays while earlier work drains)
    mutable bool needs_order;  // something other than an SpMV launch entered this context since the last
                           // SpMV launch (an upload, a build, a memset, an event ...): the next SpMV launch
                           // is then fully stream-ordered, never a programmatic dependent -- only
                           // back-to-back SpMV launches overlap, so matrix arrays are never read early
};
// scratch[kWatchFlag]: set by a kernel whose mbarrier wait ran into its spin limit (never expected;
// reported by the next b200_sync / b200_memcpy_d2h instead of hanging the device)
constexpr int kWatchFlag = 512;

struct b200_event {
    cudaEvent_t ev;
    int device;
};

void b200_set_error(const char *fmt, ...);
int b200_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define B200_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return b200_cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define B200_REQUIRE(cond, msg)                                                \
    do {                                                                       \
        if (!(cond)) {                                                         \
            b200_set_error("%s: %s", __func__, msg);                           \
            return B200_ERR_INVALID_VALUE;                                     \
        }                                                                      \
    } while (0)

#define B200_LAUNCH_CHECK() B200_CUDA(cudaGetLastError())

static inline int b200_ctx_enter(const b200_ctx *ctx)
{
    if (!ctx) {
        b200_set_error("null context");
        return B200_ERR_INVALID_VALUE;
    }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return b200_cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    ctx->needs_order = true;
    return B200_SUCCESS;
}
#define B200_ENTER(ctx)                          \
    do {                                         \
        int rc__ = b200_ctx_enter(ctx);          \
        if (rc__) return rc__;                   \
    } while (0)
// the SpMV entry points: entering does not by itself break a chain of overlapping launches
#define B200_ENTER_SPMV(ctx)                                     \
    do {                                                         \
        const bool order__ = (ctx) ? (ctx)->needs_order : true;  \
        int rc__ = b200_ctx_enter(ctx);                          \
        if (rc__) return rc__;                                   \
        (ctx)->needs_order = order__;                            \
    } while (0)
// launch overlap applies to launches of at most a few waves: there the ~2 us of ramp and drain
// between kernels matter (a cant-sized SpMV lasts ~10 us); a 1 GB matrix keeps the plain kernels
static inline bool ovl_on(const b200_ctx *ctx, long long threads)
{
    return ctx->overlap && threads <= 8ll * ctx->sm_count * 2048;
}

// value of a tuning hook, or `dflt` when it is unset
static inline int opt_or(const b200_ctx *ctx, b200_opt o, int dflt)
{
    return ctx->opt[o] == kOptUnset ? dflt : ctx->opt[o];
}
static inline bool opt_set(const b200_ctx *ctx, b200_opt o) { return ctx->opt[o] != kOptUnset; }

// NVTX range around a C-ABI call (build / H2D / spmv / exchange show up as named
// ranges in Nsight Systems; without a tool attached a push/pop is one predicted branch each)
struct b200_nvtx_scope {
    explicit b200_nvtx_scope(const char *name) { nvtxRangePushA(name); }
    ~b200_nvtx_scope() { nvtxRangePop(); }
};
#define B200_TRACE(name) b200_nvtx_scope nvtx_scope__(name)

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline unsigned ceil_div_u(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

#ifdef __CUDACC__
// Launch on the context's stream; ovl = the kernel is an OVL variant (see pdl_wait below): it becomes
// a programmatic dependent of the previous kernel on the stream, unless anything but SpMV launches
// entered the context since (needs_order).
template <typename... KArgs, typename... Args>
static inline cudaError_t b200_launch(const b200_ctx *ctx, bool ovl, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                                      size_t smem, Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (ovl && !ctx->needs_order) ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (ovl) ctx->needs_order = false;  // the chain (re)starts here: this kernel waits before touching x / y
    return e;
}

// ---------------------------------------------------------------------------------------
// Loads.  Matrix arrays (indices, values) are read exactly once per SpMV: stream them with
// evict-first (ld.global.cs) so they do not push x out of L1/L2.  x is gathered through the
// read-only path (ld.global.nc) and is the only data worth caching.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ld_stream(const int4 *p) { return __ldcs(p); }
__device__ __forceinline__ int2 ld_stream(const int2 *p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ long long ld_stream(const long long *p) { return __ldcs(p); }

template <typename T>
__device__ __forceinline__ T ld_x(const T *x, int c)
{
    return __ldg(x + c);
}

// four consecutive values as one (fp32) or two (fp64) 128-bit loads
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    float v[4];
    __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = v[3] = 0.f; }
    __device__ __forceinline__ void load(const float *p)
    {
        float4 t = ld_stream(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void load_shared(const float *p)  // p in shared memory, 16-byte aligned
    {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <>
struct Vec4<double> {
    double v[4];
    __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = v[3] = 0.0; }
    __device__ __forceinline__ void load(const double *p)
    {
        double2 a = ld_stream(reinterpret_cast<const double2 *>(p));
        double2 b = ld_stream(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    __device__ __forceinline__ void load_shared(const double *p)
    {
        double2 a = *reinterpret_cast<const double2 *>(p);
        double2 b = *(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
struct IVec4 {
    int v[4];
    __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = v[3] = 0; }
    __device__ __forceinline__ void load(const int *p)
    {
        int4 t = ld_stream(reinterpret_cast<const int4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void load_shared(const int *p)
    {
        int4 t = *reinterpret_cast<const int4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};

// Programmatic dependent launch (PDL), template flag OVL on the SpMV kernels.  An OVL kernel lets its
// successor start at once (launch_dependents at the top) and reads x / writes y only after
// pdl_wait_once(), i.e. after all earlier work on the stream has completed and is visible; what it
// does BEFORE the wait -- row pointers and the first batch of index/value loads -- touches only the
// matrix arrays.  x must then be gathered with COHERENT loads (ld.global.ca): ptxas treats
// ld.global.nc data as immutable and hoists such loads above griddepcontrol.wait (seen in SASS: the
// gathers landed before ACQBULK), which would read x while the previous launch is still writing it.
// OVL = false (the default launch path) compiles to exactly the kernels without any of this.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
template <bool OVL>
__device__ __forceinline__ void pdl_wait_once(bool &waited)
{
    if (OVL && !waited) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        waited = true;
    }
}
template <bool OVL, typename T>
__device__ __forceinline__ T ld_xo(const T *x, int c)
{
    return OVL ? __ldca(x + c) : __ldg(x + c);
}

// Load batching.  The kernels issue U groups of matrix loads, then the x gathers, then the FMAs.
// ptxas, left alone, sinks the later loads of a batch below the first gathers (fewer live
// registers), and the in-order warp then sits out a full memory latency before those loads are even
// issued.  batch_hold() returns a value that is 0 at run time -- column indices are >= 0, so the
// sign bit of (index & value bits) is clear -- but data-dependent on every load of the batch; adding
// it to the gather indices pins the schedule to "all loads, then all gathers" (checked with
// cuobjdump -sass).  Groups that were not loaded must have c zeroed.
__device__ __forceinline__ int hold_bits(const IVec4 &c, const Vec4<float> &v)
{
    return c.v[0] & __float_as_int(v.v[0]);
}
__device__ __forceinline__ int hold_bits(const IVec4 &c, const Vec4<double> &v)
{
    return c.v[0] & __double2hiint(v.v[0]) & __double2hiint(v.v[2]);
}
template <int U, typename T>
__device__ __forceinline__ int batch_hold(const IVec4 (&c)[U], const Vec4<T> (&v)[U])
{
    if (U == 1) return 0;
    int h = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) h |= hold_bits(c[u], v[u]);
    return h >> 31;
}

template <int LANES, typename T>
__device__ __forceinline__ T subwarp_sum(T v)
{
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
#endif  // __CUDACC__
