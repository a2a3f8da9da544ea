// spmv_sell_ell.cu -- SELL-32(-sigma) and column-major ELL SpMV for sm_100a.
//
// SELL replaces kernels/Sigma_C.cl:1-18 (one 32-lane work-group per slice, lane = row, scalar
// loads).  B200 design: a chunk (C = 32 rows) is exactly one warp's worth of rows and is stored
// column-major, so a chunk of width w is a flat run of 32*w entries.  The warp walks that run
// with 128-bit loads: lane L fetches entries 4g..4g+3 of group g = 32*it + L, i.e. rows
// 4*(L%8)..+3 of column 4*it + L/8.  Each lane therefore owns four row accumulators for one
// quarter of the columns; two __shfl_xor_sync steps (8, 16) fold the quarters and lanes 0-7
// store the 32 results as 128-bit writes (or scatter through the sigma permutation).
//
// Variants in this file: 1-8 warps per chunk for matrices with few chunks (WPC), a lane = row path for
// stencil-width chunks, an opt-in kernel that stages the chunks through shared memory with the TMA
// engine (sell32_tma_kernel), the fused SpMV + exchange kernels of the multi-GPU power iteration
// (sell32_bcast_kernel: halo-limited peer stores; ring_sync_kernel: collective-free hand-over of the
// norm), and sell32_pipe_kernel: the persistent, software-pipelined kernel that serves large stencil
// matrices both in the fused step and in the plain SpMV (next chunk's loads in flight while the current
// one is gathered; profiles/r2_pipelined_kernels.md).
//
// Column-major ELL is the same idea with one "chunk" spanning the whole matrix: thread t owns rows
// 4t..4t+3 and walks the K columns (optionally split over KS thread-slices when the matrix has
// too few rows to fill 148 SMs).  Bytes: P*(4+V) + (S+1)*P_bytes [+R*4 perm] + Cn*V + R*V for
// SELL, R*K*(4+V) + Cn*V + R*V for ELL; both HBM-bound.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;

// Narrow chunks (width <= kNarrowW columns: stencil matrices, 7 columns for the 7-point Laplacian).
// The 128-bit mapping below gives such a chunk only one or two groups per lane -- a chain of
// dependent round trips (pointer -> group -> gather -> group -> gather) on a warp that moves 2.7 KB --
// and its gathers touch 4 columns x 2 lines per instruction.  Here lane L owns row L (the reference's
// mapping, kernels/Sigma_C.cl:8-15): ALL of the chunk's columns are loaded at once (w coalesced 128-
// byte index loads and 128/256-byte value loads), then all w gathers (each 32 consecutive rows of one
// column: 2-3 lines), then the FMAs.  Two round trips per chunk, whatever w; the result is already
// "lane = row", so the store is one coalesced 128/256-byte write.
constexpr int kNarrowW = 8;

template <typename T, bool OVL>
__device__ __forceinline__ T narrow_chunk_dot(const T *__restrict__ dp, const int *__restrict__ ip,
                                              const T *__restrict__ x, int w, int lane, bool &waited)
{
    int c[kNarrowW];
    T v[kNarrowW];
#pragma unroll
    for (int k = 0; k < kNarrowW; ++k) {
        c[k] = 0;
        v[k] = T(0);
        if (k < w) {
            c[k] = ld_stream(ip + 32 * k + lane);
            v[k] = ld_stream(dp + 32 * k + lane);
        }
    }
    // 0 at run time (column indices are >= 0), but it makes every gather wait for ALL loads above
    int hold = 0;
#pragma unroll
    for (int k = 0; k < kNarrowW; ++k)
        hold |= c[k] & (sizeof(T) == 8 ? __double2hiint((double)v[k]) : __float_as_int((float)v[k]));
    hold >>= 31;
    pdl_wait_once<OVL>(waited);
    T xv[kNarrowW];
#pragma unroll
    for (int k = 0; k < kNarrowW; ++k) xv[k] = ld_xo<OVL>(x, c[k] + hold);  // unused slots: x[0] * 0
    T a0 = 0, a1 = 0;
#pragma unroll
    for (int k = 0; k < kNarrowW; k += 2) {
        a0 += v[k] * xv[k];
        a1 += v[k + 1] * xv[k + 1];
    }
    return a0 + a1;
}

// Columns [seg*wmax, (seg+1)*wmax) of one chunk.  seg == 0 is the main pass (plain store, wmax = 0
// means the whole chunk); seg > 0 are the extra segments of chunks wider than wmax columns
// (power-law inputs: a hub row makes one chunk 10^5 columns wide), listed by the plan and
// accumulated with atomics after the main pass.
//   U   = groups of four entries a lane loads per round trip: all 2U/3U 128-bit loads first, then
//         the 4U gathers, then the FMAs (see row_dot_vec in spmv_csr.cu for why this is explicit);
//   WPC = warps that share one chunk (main pass only).  A matrix with few chunks (cant: 1952) gives
//         one-warp-per-chunk only 13 warps per SM, each walking ~20 dependent round trips; WPC warps
//         take the chunk's 32-group rounds in turn and their 32 row sums meet in shared memory
//         (fixed order: deterministic, no atomics).
template <typename T, typename P, bool EXTRA, int WPC, int U, bool OVL = false>
__global__ void __launch_bounds__(kBlock)
sell32_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
              T *__restrict__ y, const P *__restrict__ slice_ptr, int n_work, int n_out,
              const int *__restrict__ perm, int wmax, const int2 *__restrict__ items)
{
    static_assert(!EXTRA || WPC == 1, "extra segments are one warp each");
    __shared__ T red[WPC > 1 ? kBlock / 32 : 1][32];
    pdl_launch_dependents();
    bool waited = false;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long work = ((long long)blockIdx.x * kBlock + threadIdx.x) / (32 * WPC);
    const int part = warp % WPC;  // which of the chunk's WPC warps this is
    const bool active = work < n_work;
    if (WPC == 1 && !active) return;  // whole warps leave together
    long long slice = work;
    int seg = 0;
    if (EXTRA) {
        const int2 it = items[work];
        slice = it.x;
        seg = it.y;
    }
    T acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    if (active) {
        const long long chunk_base = slice_ptr[slice];
        long long n_groups = ((long long)slice_ptr[slice + 1] - chunk_base) >> 2;  // 8 per column
        long long g_begin = 0;
        if (wmax > 0) {
            g_begin = (long long)seg * wmax * 8;
            n_groups = min(n_groups, g_begin + (long long)wmax * 8);
        }
        const int *ip = idx + chunk_base;
        const T *dp = data + chunk_base;
        if (WPC == 1 && !EXTRA && wmax == 0 && n_groups <= 8 * kNarrowW) {  // warp-uniform
            const T a = narrow_chunk_dot<T, OVL>(dp, ip, x, (int)(n_groups >> 3), lane, waited);
            pdl_wait_once<OVL>(waited);
            const long long r = slice * 32 + lane;
            if (r < n_out) y[perm ? perm[r] : r] = a;
            return;
        }
        for (long long g0 = g_begin + 32 * part + lane; g0 < n_groups; g0 += 32 * WPC * U) {
            IVec4 c[U];
            Vec4<T> v[U];
            T xv[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long g = g0 + 32 * WPC * u;
                c[u].zero();
                v[u].zero();
                if (g < n_groups) {
                    c[u].load(ip + (g << 2));
                    v[u].load(dp + (g << 2));
                }
            }
            const int hold = batch_hold<U, T>(c, v);  // 0; orders the gathers after ALL loads (common.cuh)
            pdl_wait_once<OVL>(waited);  // x may still be being written by the previous launch
            // a group that was not loaded reads x[0] and multiplies it by 0: every group of a chunk
            // is whole (padding slots are (col 0, value 0) in the format itself)
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) xv[u][k] = ld_xo<OVL>(x, c[u].v[k] + hold);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc0 += v[u].v[0] * xv[u][0];
                acc1 += v[u].v[1] * xv[u][1];
                acc2 += v[u].v[2] * xv[u][2];
                acc3 += v[u].v[3] * xv[u][3];
            }
        }
    }
    pdl_wait_once<OVL>(waited);  // empty chunks never waited: y must not be written early either
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, off);
        acc3 += __shfl_xor_sync(0xffffffffu, acc3, off);
    }
    if (WPC == 1) {
        if (lane < 8) {
            const long long r = slice * 32 + lane * 4;
            const T a[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (r + k < n_out) {
                    T *dst = y + (perm ? perm[r + k] : r + k);
                    if (EXTRA) atomicAdd(dst, a[k]);
                    else *dst = a[k];
                }
            }
        }
    } else {
        if (lane < 8) {
            red[warp][lane * 4 + 0] = acc0;
            red[warp][lane * 4 + 1] = acc1;
            red[warp][lane * 4 + 2] = acc2;
            red[warp][lane * 4 + 3] = acc3;
        }
        __syncthreads();
        if (part == 0 && active) {
            T total = red[warp][lane];
#pragma unroll
            for (int w = 1; w < WPC; ++w) total += red[warp + w][lane];
            const long long r = slice * 32 + lane;
            if (r < n_out) y[perm ? perm[r] : r] = total;
        }
    }
}

// ---- SELL through the TMA engine (bulk asynchronous copies into shared memory) ----------------
// A chunk is one contiguous run of 32*w entries in both arrays, so the copy engine can stream it:
// every warp owns a two-stage ring of kPiece-entry buffers in shared memory and one mbarrier per
// stage.  Lane 0 arms the barrier with the byte count and issues two cp.async.bulk copies (indices,
// values); the warp waits on the barrier, reads its groups with conflict-free LDS.128, gathers x
// and accumulates, then re-arms the freed stage with the piece two ahead.  The ring runs ACROSS
// chunk boundaries (a warp walks chunks gw, gw + W, ...), so a warp always has one or two pieces
// (4-12 KiB) in flight without holding a single register for them: ~100 KiB in flight per SM at 16
// resident warps, where the register-staged kernel needs 64 warps x U groups for the same.
// Persistent grid: blocks = min(chunks / 8, SMs x blocks that fit in 227 KiB of shared memory).
constexpr int kPiece = 512;   // entries per stage: 16 columns of a chunk
constexpr int kStages = 2;
constexpr unsigned long long kWaitLimitNs = 2000000000ull;  // a 2 s wait is a bug: flag it, do not hang

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// suspend_ns > 0: the warp may be parked for up to that long before "not yet" is reported
// (NANOSLEEP.SYNCS in SASS): a waiting warp should sleep, not poll
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity, unsigned suspend_ns)
{
    unsigned done;
    if (suspend_ns > 0) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(suspend_ns)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes,
                                         unsigned long long *bar, unsigned long long policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

template <typename T, typename P>
__global__ void __launch_bounds__(kBlock)
sell32_tma_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
                  T *__restrict__ y, const P *__restrict__ slice_ptr, int n_slices, int n_out,
                  const int *__restrict__ perm, int *__restrict__ err_flag, unsigned suspend_ns)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kWarps = kBlock / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *s_idx = reinterpret_cast<int *>(smem_raw) + warp * kStages * kPiece;
    T *s_val = reinterpret_cast<T *>(smem_raw + (size_t)kWarps * kStages * kPiece * sizeof(int)) +
               warp * kStages * kPiece;
    unsigned long long *bars =
        reinterpret_cast<unsigned long long *>(smem_raw + (size_t)kWarps * kStages * kPiece * (sizeof(int) + sizeof(T))) +
        warp * kStages;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kStages; ++st) mbar_init(bars + st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned long long policy;  // matrix data is read once: evict-first in L2, like ld.global.cs
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    const long long n_warps = (long long)gridDim.x * kWarps;
    const long long gw = (long long)blockIdx.x * kWarps + warp;

    // producer cursor: the next piece to request = entries [p_off, p_off + kPiece) of chunk p_slice
    long long p_slice = gw, p_base = 0, p_n = 0, p_off = 0;
    bool p_valid = p_slice < n_slices;
    if (p_valid) {
        p_base = slice_ptr[p_slice];
        p_n = (long long)slice_ptr[p_slice + 1] - p_base;
    }
    auto p_normalize = [&]() {  // skip past exhausted (or empty) chunks
        while (p_valid && p_off >= p_n) {
            p_slice += n_warps;
            p_off = 0;
            p_valid = p_slice < n_slices;
            if (p_valid) {
                p_base = slice_ptr[p_slice];
                p_n = (long long)slice_ptr[p_slice + 1] - p_base;
            }
        }
    };
    auto p_issue = [&](int stage) {
        const unsigned cnt = (unsigned)min((long long)kPiece, p_n - p_off);
        if (lane == 0) {
            mbar_expect_tx(bars + stage, cnt * (unsigned)(sizeof(int) + sizeof(T)));
            bulk_g2s(s_idx + stage * kPiece, idx + p_base + p_off, cnt * (unsigned)sizeof(int), bars + stage, policy);
            bulk_g2s(s_val + stage * kPiece, data + p_base + p_off, cnt * (unsigned)sizeof(T), bars + stage, policy);
        }
        p_off += kPiece;
        p_normalize();
    };
    p_normalize();
#pragma unroll
    for (int st = 0; st < kStages; ++st)
        if (p_valid) p_issue(st);

    unsigned consumed = 0;
    for (long long slice = gw; slice < n_slices; slice += n_warps) {
        const long long base = slice_ptr[slice];
        const long long n = (long long)slice_ptr[slice + 1] - base;
        T acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
        for (long long off = 0; off < n; off += kPiece) {
            const int stage = consumed % kStages;
            const unsigned parity = (consumed / kStages) & 1u;
            if (!mbar_try_wait(bars + stage, parity, suspend_ns)) {
                const unsigned long long t0 = global_timer_ns();
                while (!mbar_try_wait(bars + stage, parity, suspend_ns)) {
                    if (global_timer_ns() - t0 > kWaitLimitNs) {
                        if (lane == 0) atomicExch(err_flag, 1);
                        return;
                    }
                }
            }
            const int groups = (int)(min((long long)kPiece, n - off) >> 2);
            const int *si = s_idx + stage * kPiece;
            const T *sv = s_val + stage * kPiece;
            constexpr int G = kPiece / 4 / 32;  // groups per lane and piece
            IVec4 c[G];
            Vec4<T> v[G];
            T xv[G][4];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const int g = lane + 32 * u;
                c[u].zero();
                v[u].zero();
                if (g < groups) {
                    c[u].load_shared(si + 4 * g);
                    v[u].load_shared(sv + 4 * g);
                }
            }
#pragma unroll
            for (int u = 0; u < G; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) xv[u][k] = ld_x(x, c[u].v[k]);
#pragma unroll
            for (int u = 0; u < G; ++u) {
                acc0 += v[u].v[0] * xv[u][0];
                acc1 += v[u].v[1] * xv[u][1];
                acc2 += v[u].v[2] * xv[u][2];
                acc3 += v[u].v[3] * xv[u][3];
            }
            __syncwarp();  // every lane has read the stage: it may be overwritten
            ++consumed;
            if (p_valid) p_issue(stage);
        }
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
            acc2 += __shfl_xor_sync(0xffffffffu, acc2, off);
            acc3 += __shfl_xor_sync(0xffffffffu, acc3, off);
        }
        if (lane < 8) {
            const long long r = slice * 32 + lane * 4;
            const T a[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (r + k < n_out) y[perm ? perm[r + k] : r + k] = a[k];
        }
    }
}

// ---- plan: extra (chunk, segment) work items of chunks wider than kSellWmax columns ---------
constexpr int kSellWmax = 256;

template <typename P>
__global__ void sell_wide_items_kernel(const P *__restrict__ slice_ptr, int n_slices, int wmax,
                                       int *counter, int2 *items)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    const long long width = ((long long)slice_ptr[s + 1] - (long long)slice_ptr[s]) >> 5;
    if (!items && width > kNarrowW) atomicAdd(counter + 1, 1);  // first pass only: chunks off the lane = row path
    const int extra = (int)((width + wmax - 1) / wmax) - 1;
    if (extra <= 0) return;
    const int at = atomicAdd(counter, extra);
    if (items)
        for (int k = 0; k < extra; ++k) items[at + k] = make_int2((int)s, k + 1);
}

// ---- fused SpMV + exchange for the iterated (power-iteration) mode -------------------------
// y = (A_r . x) / sqrt(*scale2) is written STRAIGHT INTO the gather buffers of all ranks: dst.p[d]
// is the next-x buffer of rank d (its own for d == rank, the others mapped through CUDA IPC, i.e.
// plain stores over NVLink/NVSwitch), so the all-gather of the reference-style formulation
// disappears -- the transfer overlaps the SpMV chunk by chunk.  The only synchronisation left per
// step is the 1-element all-reduce of ||y||^2, which also orders the peer writes.
constexpr int kMaxPeers = 16;
template <typename T>
struct PeerDst {
    T *p[kMaxPeers];
    // rows [lo[d], hi[d]) of THIS rank's block (local numbering) that destination d reads as columns:
    // the whole block for the rank itself, the halo for a neighbour, nothing (lo >= hi) for a rank
    // whose rows never touch these columns.  A banded matrix then moves kilobytes per step instead of
    // the full vector (7-point Laplacian: one 160 000-row plane per neighbour instead of 8 M rows to
    // all 7 peers).
    int lo[kMaxPeers], hi[kMaxPeers];
};

// ||y||^2 is accumulated by the same kernel into kSumsqSlots partial sums (one atomic per block,
// spread over the slots so that no single L2 address serialises them); the consumer -- the next
// step's kernel -- adds the slots up after the ranks' all-reduce.
constexpr int kSumsqSlots = 32;

// RING variant: no NCCL call in the loop at all.  Every rank owns a small "sync block" in device
// memory that all ranks map (CUDA IPC), laid out in 8-byte words:
//   [kSyncFlags + r]              steps completed by rank r, written by r into EVERY rank's block
//   [kSyncSums + (p*16 + r)*32 + s]  the 32 partial sums of ||y_r||^2 at step parity p, written by r
//   [kSyncDone]                   this rank's "blocks finished" ticket counter (local use)
//   [kSyncAcc + s]                this rank's running partial sums of the current step (local use)
//   [kSyncScale + s]              sum over ranks of the previous step's partial sums (local use)
// Step k = two launches: the plain fused SpMV + halo-store kernel (scale from [kSyncScale], ||y||^2
// share added to [kSyncAcc]) and ring_sync_kernel below.
constexpr int kSyncFlags = 0;
constexpr int kSyncSums = 16;
constexpr int kSyncDone = kSyncSums + 2 * 16 * 32;
constexpr int kSyncAcc = kSyncDone + 2;
constexpr int kSyncScale = kSyncAcc + 32;
constexpr int kSyncWords = kSyncScale + 32;
static_assert(kSyncWords * 8 <= B200_SYNC_BLOCK_BYTES, "sync block layout exceeds B200_SYNC_BLOCK_BYTES");

struct PeerSync {
    unsigned long long *blk[kMaxPeers];  // sync block of every rank (own + IPC-mapped peers), by rank
    unsigned long long *mine;            // == blk[my_rank]
    int my_rank, world;
    unsigned long long step;
    int *err_flag;
    int flush_after_flag;  // B200_RING_FLUSH (default 1): system fence AFTER the flag stores as well
    int poll_mode;         // B200_RING_POLL: 0 acquire loads (default), 1 relaxed loads + one fence at the end
    int sleep_ns;          // B200_RING_SLEEP_NS between polls (default 100; 0 = none)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename T, typename P>
__global__ void __launch_bounds__(kBlock)
sell32_bcast_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
                    const P *__restrict__ slice_ptr, int n_slices, int n_rows,
                    const T *__restrict__ scale2, T *__restrict__ sumsq_out, PeerDst<T> dst, int n_dst,
                    long long dst_offset)
{
    __shared__ T warp_sq[kBlock / 32];
    const int lane = threadIdx.x & 31;
    const long long slice = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const bool active = slice < n_slices;  // no early return: the block reduces ||y||^2 together
    // 1/||x||: one coalesced load of the 32 partial sums per warp, issued now and folded (shuffles +
    // rsqrt) only after the dot product, so its latency hides behind the matrix loads
    T scale_part = 0;
    if (scale2) scale_part = __ldg(scale2 + lane);
    T mine = 0;  // row `lane` of this warp's chunk
    if (active) {
        const long long chunk_base = slice_ptr[slice];
        const long long n_groups = ((long long)slice_ptr[slice + 1] - chunk_base) >> 2;
        const int *ip = idx + chunk_base;
        const T *dp = data + chunk_base;
        if (n_groups <= 8 * kNarrowW) {  // warp-uniform: stencil-width chunk, lane = row (see narrow_chunk_dot)
            bool waited = true;
            mine = narrow_chunk_dot<T, false>(dp, ip, x, (int)(n_groups >> 3), lane, waited);
        } else {
            T acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll 4
            for (long long g = lane; g < n_groups; g += 32) {
                IVec4 c;
                Vec4<T> v;
                c.load(ip + (g << 2));
                v.load(dp + (g << 2));
                acc0 += v.v[0] * ld_x(x, c.v[0]);
                acc1 += v.v[1] * ld_x(x, c.v[1]);
                acc2 += v.v[2] * ld_x(x, c.v[2]);
                acc3 += v.v[3] * ld_x(x, c.v[3]);
            }
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
                acc2 += __shfl_xor_sync(0xffffffffu, acc2, off);
                acc3 += __shfl_xor_sync(0xffffffffu, acc3, off);
            }
            // lanes 0-7 hold rows 4L..4L+3; re-deal them so that lane j holds row j
            const int src = lane >> 2;
            const T b0 = __shfl_sync(0xffffffffu, acc0, src), b1 = __shfl_sync(0xffffffffu, acc1, src);
            const T b2 = __shfl_sync(0xffffffffu, acc2, src), b3 = __shfl_sync(0xffffffffu, acc3, src);
            mine = (lane & 2) ? ((lane & 1) ? b3 : b2) : ((lane & 1) ? b1 : b0);
        }
    }
    if (scale2) mine *= rsqrt(subwarp_sum<32>(scale_part));
    const long long r = slice * 32 + lane;
    T sq = 0;
    if (active && r < n_rows) {
        sq = mine * mine;
        // the warp writes its 32 results as ONE contiguous 256-byte store per destination (full NVLink
        // packets).  Fully unrolled with a static index: dst lives in the constant bank; a
        // runtime-indexed loop would copy the whole struct to local memory in every thread
#pragma unroll
        for (int d = 0; d < kMaxPeers; ++d)
            if (d < n_dst && r >= dst.lo[d] && r < dst.hi[d]) dst.p[d][dst_offset + r] = mine;
    }
    if (sumsq_out) {
        sq = subwarp_sum<32>(sq);
        if (lane == 0) warp_sq[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            T total = 0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) total += warp_sq[w];
            atomicAdd(sumsq_out + (blockIdx.x & (kSumsqSlots - 1)), total);
        }
    }
}

// The same step as a PERSISTENT, software-pipelined kernel for stencil-width chunks.  sell32_bcast_kernel
// walks three dependent round trips per 2.7 KB chunk (chunk pointers -> index/value loads -> gathers), and
// only during the second one does a warp have matrix bytes in flight: measured 5.2 TB/s on the 7-point
// Laplacian = warps x 2.7 KB x 1/3 of the time / latency.  Here a warp stays resident and walks chunks
// gw, gw + W, gw + 2W, ... (W = warps of the grid; neighbouring warps work on neighbouring chunks, so the
// x gathers of a block still share lines): while chunk i is gathered, multiplied and stored, the index and
// value loads of chunk i+1 are already in flight and the pointers of chunk i+2 are being fetched, so every
// resident warp keeps one whole chunk in flight all the time.  1/||x|| is folded once per warp instead of
// once per chunk, and ||y||^2 costs one atomic per block of the (SMs x blocks-per-SM)-block grid instead of
// 31 250.  Chunks wider than kNarrowW columns take the general 128-bit path inline (not pipelined).
constexpr int kPipeDefaultBlocks = 3;
template <typename T>
struct ChunkRegs {
    int c[kNarrowW];
    T v[kNarrowW];
};

template <typename T, typename P>
__device__ __forceinline__ void chunk_issue(ChunkRegs<T> &r, const T *__restrict__ data, const int *__restrict__ idx,
                                            P base, int w, int lane)
{
#pragma unroll
    for (int k = 0; k < kNarrowW; ++k) {
        r.c[k] = 0;
        r.v[k] = T(0);
        if (k < w) {
            r.c[k] = ld_stream(idx + (long long)base + 32 * k + lane);
            r.v[k] = ld_stream(data + (long long)base + 32 * k + lane);
        }
    }
}

template <typename T>
__device__ __forceinline__ T chunk_dot(const ChunkRegs<T> &r, const T *__restrict__ x)
{
    int hold = 0;  // 0 at run time; makes every gather wait for ALL of the chunk's loads (see narrow_chunk_dot)
#pragma unroll
    for (int k = 0; k < kNarrowW; ++k)
        hold |= r.c[k] & (sizeof(T) == 8 ? __double2hiint((double)r.v[k]) : __float_as_int((float)r.v[k]));
    hold >>= 31;
    T xv[kNarrowW];
#pragma unroll
    for (int k = 0; k < kNarrowW; ++k) xv[k] = ld_x(x, r.c[k] + hold);
    T a0 = 0, a1 = 0;
#pragma unroll
    for (int k = 0; k < kNarrowW; k += 2) {
        a0 += r.v[k] * xv[k];
        a1 += r.v[k + 1] * xv[k + 1];
    }
    return a0 + a1;
}

// FUSED = false is the plain SpMV (b200_spmv_sell_*) of a matrix whose plan found stencil-width chunks only:
// y[perm ? perm[r] : r] = row sum, nothing else.
template <typename T, typename P, int MINB, bool FUSED>
__global__ void __launch_bounds__(kBlock, MINB)
sell32_pipe_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
                   const P *__restrict__ slice_ptr, int n_slices, int n_rows,
                   const T *__restrict__ scale2, T *__restrict__ sumsq_out, PeerDst<T> dst, int n_dst,
                   long long dst_offset, T *__restrict__ y, const int *__restrict__ perm)
{
    __shared__ T warp_sq[kBlock / 32];
    // the destinations that take any row at all, compacted once per block (the halo-limited exchange of a
    // banded matrix has 3 of up to 16: own block + two neighbours): the per-chunk store loop then runs over
    // those only -- the fully unrolled 16-way test of sell32_bcast_kernel is ~160 instructions per chunk
    __shared__ T *s_ptr[kMaxPeers];
    __shared__ int s_lo[kMaxPeers], s_hi[kMaxPeers], s_n;
    if (FUSED) {
        if (threadIdx.x == 0) {
            int n = 0;
#pragma unroll
            for (int d = 0; d < kMaxPeers; ++d)
                if (d < n_dst && dst.hi[d] > dst.lo[d]) {
                    s_ptr[n] = dst.p[d] + dst_offset;
                    s_lo[n] = dst.lo[d];
                    s_hi[n] = dst.hi[d];
                    ++n;
                }
            s_n = n;
        }
        __syncthreads();
    }
    const int n_act = FUSED ? s_n : 0;
    const int lane = threadIdx.x & 31;
    const long long W = ((long long)gridDim.x * kBlock) >> 5;
    long long s = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    T inv_norm = T(1);
    if (FUSED && scale2) inv_norm = rsqrt(subwarp_sum<32>(__ldg(scale2 + lane)));
    // pointers of the first two chunks of this warp, loads of the first
    P b0 = 0, e0 = 0, b1 = 0, e1 = 0;
    if (s < n_slices) {
        b0 = slice_ptr[s];
        e0 = slice_ptr[s + 1];
    }
    if (s + W < n_slices) {
        b1 = slice_ptr[s + W];
        e1 = slice_ptr[s + W + 1];
    }
    ChunkRegs<T> cur, nxt;
    int w0 = (int)(((long long)e0 - (long long)b0) >> 5);  // columns of the chunk
    if (s < n_slices && w0 <= kNarrowW) chunk_issue<T, P>(cur, data, idx, b0, w0, lane);
    T sq = 0;
    for (; s < n_slices; s += W) {
        // chunk i+2: pointers;  chunk i+1: index / value loads -- all issued before chunk i is touched
        P b2 = 0, e2 = 0;
        if (s + 2 * W < n_slices) {
            b2 = slice_ptr[s + 2 * W];
            e2 = slice_ptr[s + 2 * W + 1];
        }
        const int w1 = (int)(((long long)e1 - (long long)b1) >> 5);
        const bool have1 = s + W < n_slices;
        if (have1 && w1 <= kNarrowW) chunk_issue<T, P>(nxt, data, idx, b1, w1, lane);
        T mine;
        if (w0 <= kNarrowW) {
            mine = chunk_dot<T>(cur, x);
        } else {  // general chunk: 128-bit groups, lanes 0-7 end up with rows 4L..4L+3, re-dealt to lane = row
            const long long n_groups = ((long long)e0 - (long long)b0) >> 2;
            const int *ip = idx + (long long)b0;
            const T *dp = data + (long long)b0;
            T acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll 2
            for (long long g = lane; g < n_groups; g += 32) {
                IVec4 c;
                Vec4<T> v;
                c.load(ip + (g << 2));
                v.load(dp + (g << 2));
                acc0 += v.v[0] * ld_x(x, c.v[0]);
                acc1 += v.v[1] * ld_x(x, c.v[1]);
                acc2 += v.v[2] * ld_x(x, c.v[2]);
                acc3 += v.v[3] * ld_x(x, c.v[3]);
            }
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
                acc2 += __shfl_xor_sync(0xffffffffu, acc2, off);
                acc3 += __shfl_xor_sync(0xffffffffu, acc3, off);
            }
            const int src = lane >> 2;
            const T q0 = __shfl_sync(0xffffffffu, acc0, src), q1 = __shfl_sync(0xffffffffu, acc1, src);
            const T q2 = __shfl_sync(0xffffffffu, acc2, src), q3 = __shfl_sync(0xffffffffu, acc3, src);
            mine = (lane & 2) ? ((lane & 1) ? q3 : q2) : ((lane & 1) ? q1 : q0);
        }
        const long long r = s * 32 + lane;
        if (FUSED) {
            mine *= inv_norm;
            if (r < n_rows) {
                sq += mine * mine;
                for (int i = 0; i < n_act; ++i)
                    if (r >= s_lo[i] && r < s_hi[i]) s_ptr[i][r] = mine;
            }
        } else if (r < n_rows) {
            y[perm ? perm[r] : r] = mine;
        }
        cur = nxt;
        w0 = w1;
        b0 = b1;
        e0 = e1;
        b1 = b2;
        e1 = e2;
    }
    if (FUSED && sumsq_out) {
        sq = subwarp_sum<32>(sq);
        if (lane == 0) warp_sq[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            T total = 0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) total += warp_sq[w];
            atomicAdd(sumsq_out + (blockIdx.x & (kSumsqSlots - 1)), total);
        }
    }
}

// Second (one-warp) launch of a ring step -- the whole "all-reduce + barrier" of the step:
//   publish  this rank's 32 partial sums of ||y||^2 go to every rank's sync block, a system-wide fence,
//            then flag[my_rank] = k+1 is released in every rank's block.  The kernel runs after the SpMV
//            kernel on the same stream, i.e. after ALL of its stores -- own and peer -- have been
//            performed, so the release covers this rank's halo stores too;
//   wait     until every rank's flag in THIS rank's block is k+1 (acquire): their x rows and sums are
//            then in this GPU's memory, and they are done reading the buffers step k+1 will overwrite;
//   fold     the ranks' sums into the 32 local slots the next SpMV kernel takes 1/||x|| from.
// (A first version did all this inside the SpMV kernel -- flag wait in every block, "last block"
// ticket -- and cost +0.03 ms per step on ONE GPU: every block sat out an acquire and a ticket round
// trip.  Two plain launches are cheaper.)
__global__ void ring_sync_kernel(PeerSync sync)
{
    const int lane = threadIdx.x;
    double *acc = reinterpret_cast<double *>(sync.mine + kSyncAcc);
    const double v = acc[lane];
    acc[lane] = 0.0;  // ready for the next step's SpMV kernel
    const long long base = (long long)(sync.step & 1) * 16 * 32;
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d)
        if (d < sync.world) reinterpret_cast<double *>(sync.blk[d] + kSyncSums)[base + (long long)sync.my_rank * 32 + lane] = v;
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < kMaxPeers; ++d)
            if (d < sync.world) st_release_sys(sync.blk[d] + kSyncFlags + sync.my_rank, sync.step + 1);
    }
    // push the flag stores out now: without a fence behind them they may sit in the write path until
    // the kernel ends -- and this kernel only ends when the PEERS' flags have arrived
    if (sync.flush_after_flag) __threadfence_system();
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        unsigned long long f = ~0ull;
        if (lane < sync.world)
            f = sync.poll_mode ? ld_relaxed_sys(sync.mine + kSyncFlags + lane) : ld_acquire_sys(sync.mine + kSyncFlags + lane);
        if (__all_sync(0xffffffffu, f >= sync.step + 1)) break;
        if (sync.sleep_ns > 0) __nanosleep(sync.sleep_ns);
        if (global_timer_ns() - t0 > kWaitLimitNs) {  // a peer died: flag it, do not hang
            if (lane == 0) atomicExch(sync.err_flag, 2);
            break;
        }
    }
    if (sync.poll_mode) __threadfence_system();
    const double *sums = reinterpret_cast<const double *>(sync.mine + kSyncSums) + base;
    double total = 0;
    for (int r = 0; r < sync.world; ++r) total += __ldcg(sums + r * 32 + lane);
    reinterpret_cast<double *>(sync.mine + kSyncScale)[lane] = total;
}

// scalar-load variant for unaligned arrays: lane = row (the reference's mapping)
template <typename T, typename P>
__global__ void __launch_bounds__(kBlock)
sell32_scalar_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
                     T *__restrict__ y, const P *__restrict__ slice_ptr, int n_slices, int n_out,
                     const int *__restrict__ perm)
{
    const int lane = threadIdx.x & 31;
    const long long slice = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (slice >= n_slices) return;
    const long long end = slice_ptr[slice + 1];
    T acc = 0;
    for (long long j = (long long)slice_ptr[slice] + lane; j < end; j += 32)
        acc += ld_stream(data + j) * ld_x(x, ld_stream(idx + j));
    const long long r = slice * 32 + lane;
    if (r < n_out) y[perm ? perm[r] : r] = acc;
}

// column-major ELL: blockDim = (BX threads, KS column slices); a thread owns Q quads of 4 rows,
// quad q of thread t = rows 4*(q*BX + t) .. +3 of the block's BX*4*Q-row panel, so that a warp
// instruction still reads 512 contiguous bytes and a block reads BX*16*Q contiguous bytes per
// column and array (the column stride is pitch*4 bytes: fewer, longer DRAM bursts per column).
template <typename T, int KS, int BX, int Q>
__global__ void __launch_bounds__(BX * KS)
ellcm_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
             T *__restrict__ y, int n_rows, int row_size, int pitch)
{
    __shared__ T red[KS > 1 ? KS : 1][KS > 1 ? BX : 1][4 * Q];
    const long long panel = (long long)blockIdx.x * BX * 4 * Q;
    T acc[4 * Q];
#pragma unroll
    for (int i = 0; i < 4 * Q; ++i) acc[i] = 0;
#pragma unroll 4
    for (int k = threadIdx.y; k < row_size; k += KS) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const long long r0 = panel + ((long long)q * BX + threadIdx.x) * 4;
            if (r0 < pitch) {
                IVec4 c;
                Vec4<T> v;
                c.load(idx + (long long)k * pitch + r0);
                v.load(data + (long long)k * pitch + r0);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[4 * q + i] += v.v[i] * ld_x(x, c.v[i]);
            }
        }
    }
    if (KS > 1) {
#pragma unroll
        for (int i = 0; i < 4 * Q; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
        __syncthreads();
        if (threadIdx.y != 0) return;
#pragma unroll
        for (int s = 1; s < KS; ++s)
#pragma unroll
            for (int i = 0; i < 4 * Q; ++i) acc[i] += red[s][threadIdx.x][i];
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const long long r0 = panel + ((long long)q * BX + threadIdx.x) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r0 + i < n_rows) y[r0 + i] = acc[4 * q + i];
    }
}

}  // namespace

struct b200_sell_plan {
    int device;
    int n_slices;
    int n_items;   // extra (chunk, segment) work items
    int wmax;      // columns per segment
    int2 *items;   // device
    int n_not_narrow;  // chunks wider than kNarrowW columns (0: a stencil matrix, served by the pipelined kernel)
};

namespace {

template <typename P>
int sell_plan_create_impl(b200_ctx *ctx, const P *slice_ptr, int n_slices, b200_sell_plan **plan)
{
    B200_ENTER(ctx);
    B200_REQUIRE(slice_ptr && plan && n_slices >= 0, "bad argument");
    *plan = nullptr;
    b200_sell_plan *p = new b200_sell_plan();
    p->device = ctx->device;
    p->n_slices = n_slices;
    p->n_items = 0;
    p->n_not_narrow = n_slices;
    p->wmax = kSellWmax;
    p->items = nullptr;
    if (n_slices > 0) {
        int *counter = ctx->scratch + 192;
        const unsigned blocks = ceil_div_u(n_slices, 256);
        int counts[2] = {0, 0};
        cudaError_t e = cudaMemsetAsync(counter, 0, 2 * sizeof(int), ctx->stream);
        sell_wide_items_kernel<P><<<blocks, 256, 0, ctx->stream>>>(slice_ptr, n_slices, p->wmax, counter, nullptr);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(counts, counter, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        const int count = counts[0];
        p->n_not_narrow = counts[1];
        if (e == cudaSuccess && count > 0) {
            e = cudaMalloc(&p->items, sizeof(int2) * (size_t)count);
            if (e == cudaSuccess) e = cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream);
            if (e == cudaSuccess) {
                sell_wide_items_kernel<P><<<blocks, 256, 0, ctx->stream>>>(slice_ptr, n_slices, p->wmax, counter, p->items);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        }
        if (e != cudaSuccess) {
            if (p->items) cudaFree(p->items);
            delete p;
            return b200_cuda_fail(e, "sell plan", __FILE__, __LINE__);
        }
        p->n_items = count;
    }
    *plan = p;
    return B200_SUCCESS;
}

template <typename T, typename P>
int spmv_sell_impl(b200_ctx *ctx, const T *data, const int *idx, const T *x, T *y, const P *slice_ptr,
                   int chunk, int n_slices, int n_out, const int *perm, const b200_sell_plan *plan)
{
    B200_TRACE("b200 spmv sell");
    B200_ENTER_SPMV(ctx);
    B200_REQUIRE(x && y && slice_ptr && n_slices >= 0 && n_out >= 0, "bad argument");
    if (chunk != 32) {
        b200_set_error("SELL chunk must be 32 (warp-aligned), got %d", chunk);
        return B200_ERR_UNSUPPORTED;
    }
    B200_REQUIRE((long long)n_out <= (long long)n_slices * 32, "n_out exceeds n_slices*32");
    B200_REQUIRE(!plan || plan->n_slices == n_slices, "plan was built for a different matrix");
    if (n_slices == 0) return B200_SUCCESS;
    unsigned blocks = ceil_div_u((long long)n_slices * 32, kBlock);
    if (aligned16(data) && aligned16(idx)) {
        const int wmax = plan && plan->n_items > 0 ? plan->wmax : 0;
        // B200_SELL_TMA=1 selects the bulk-copy (TMA engine) staged kernel with its persistent grid.
        // Opt-in: measured on B200 (profiles/r1e_variant_sweep.md) it is the fastest SELL kernel in
        // isolation on wide chunks (banded fp32: 157-161 us vs 162-174, a third of the L1 traffic),
        // equal within 1 % in sustained power-capped runs (0.1689 vs 0.1671 ms), 5 % slower in fp64,
        // and much slower on narrow chunks (7-point stencil: 2.7 KB pieces, 0.248 vs 0.161 ms).
        const bool tma = opt_or(ctx, OPT_SELL_TMA, 0) != 0;
        if (tma && wmax == 0) {
            constexpr size_t smem = (size_t)(kBlock / 32) * kStages * (kPiece * (sizeof(int) + sizeof(T)) + 8);
            B200_CUDA(cudaFuncSetAttribute(sell32_tma_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // resident blocks per SM (B200_SELL_TMA_BLOCKS=1|2|3): each holds 8 warps x 2 stages in
            // flight; the shared-memory carve-out is sized for exactly that many, the rest stays L1
            // for the x gather
            const long long fit = (227 * 1024) / (long long)(smem + 1024);
            long long per_sm = fit < 2 ? fit : 2;
            {
                const int v = opt_or(ctx, OPT_SELL_TMA_BLOCKS, 0);
                if (v >= 1 && v <= fit) per_sm = v;
            }
            const int carve = (int)min(100ll, (per_sm * (long long)(smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
            B200_CUDA(cudaFuncSetAttribute(sell32_tma_kernel<T, P>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            const long long want = ((long long)n_slices + kBlock / 32 - 1) / (kBlock / 32);
            const unsigned grid = (unsigned)min(want, (long long)ctx->sm_count * per_sm);
            // how long a waiting warp may sleep before it polls again (B200_SELL_TMA_SUSPEND_NS; 0 = poll)
            const unsigned suspend_ns = (unsigned)opt_or(ctx, OPT_SELL_TMA_SUSPEND_NS, 1000);
            sell32_tma_kernel<T, P><<<grid, kBlock, smem, ctx->stream>>>(data, idx, x, y, slice_ptr, n_slices, n_out, perm,
                                                                        ctx->scratch + kWatchFlag, suspend_ns);
            B200_LAUNCH_CHECK();
            ctx->watch_flag = true;
            return B200_SUCCESS;
        }
        // Stencil matrices (the plan found no chunk wider than kNarrowW columns) in launches of at least 4 chunks
        // per resident warp: the persistent software-pipelined kernel (sell32_pipe_kernel; 7-point Laplacian,
        // 8 M rows, fp64: see profiles/r2_pipelined_kernels.md).  B200_SELL_PIPE = 0 turns it off, 2 | 3 | 4 force
        // it at that many blocks per SM (any matrix: wider chunks take its inline general path, un-pipelined).
        {
            int pipe = opt_or(ctx, OPT_SELL_PIPE, -1);
            if (pipe < 0)
                pipe = (plan && plan->n_not_narrow == 0 && wmax == 0 &&
                        (long long)n_slices >= 4ll * ctx->sm_count * kPipeDefaultBlocks * (kBlock / 32))
                           ? kPipeDefaultBlocks
                           : 0;
            if (pipe >= 2 && wmax == 0) {
                void (*kern)(const T *, const int *, const T *, const P *, int, int, const T *, T *, PeerDst<T>, int,
                             long long, T *, const int *) = pipe == 2   ? sell32_pipe_kernel<T, P, 2, false>
                                                           : pipe == 3 ? sell32_pipe_kernel<T, P, 3, false>
                                                                       : sell32_pipe_kernel<T, P, 4, false>;
                int per_sm = 0;
                B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlock, 0));
                if (per_sm < 1) per_sm = 1;
                unsigned pgrid = (unsigned)ctx->sm_count * (unsigned)per_sm;
                if (pgrid > blocks) pgrid = blocks;
                PeerDst<T> none;
                memset(&none, 0, sizeof none);
                B200_CUDA(b200_launch(ctx, false, kern, dim3(pgrid), dim3(kBlock), 0, data, idx, x, slice_ptr, n_slices, n_out,
                                      (const T *)nullptr, (T *)nullptr, none, 0, 0ll, y, perm));
                return B200_SUCCESS;
            }
        }
        // warps per chunk: enough warps for ~32 per SM (tuning hook B200_SELL_WPC=1|2|4|8); chunks
        // that the plan splits by columns anyway stay one warp each
        int wpc = 1;
        while (wmax == 0 && wpc < 8 && (long long)n_slices * wpc < (long long)ctx->sm_count * 32) wpc <<= 1;
        {
            const int v = opt_or(ctx, OPT_SELL_WPC, 0);
            if (wmax == 0 && (v == 1 || v == 2 || v == 4 || v == 8)) wpc = v;
        }
        // groups per lane and round trip (tuning hook B200_SELL_UNROLL=1|2|4): 4 for whole-chunk
        // warps, 2 once the chunk is shared by 4+ warps (cant: 11.3 us at WPC 4 / U 2 vs 18.3 at 1 / 1)
        int u = wpc >= 4 ? 2 : 4;
        {
            const int v = opt_or(ctx, OPT_SELL_UNROLL, 0);
            if (v == 1 || v == 2 || v == 4) u = v;
        }
        const bool ovl = ovl_on(ctx, (long long)n_slices * 32 * wpc);
#define B200_SELL_MAIN(W, UU)                                                                              \
    B200_CUDA(ovl                                                                                   \
                  ? b200_launch(ctx, ovl, sell32_kernel<T, P, false, W, UU, true>,                                \
                                dim3(ceil_div_u((long long)n_slices * 32 * W, kBlock)), dim3(kBlock), 0, data, idx, x, \
                                y, slice_ptr, n_slices, n_out, perm, wmax, nullptr)                          \
                  : b200_launch(ctx, ovl, sell32_kernel<T, P, false, W, UU, false>,                               \
                                dim3(ceil_div_u((long long)n_slices * 32 * W, kBlock)), dim3(kBlock), 0, data, idx, x, \
                                y, slice_ptr, n_slices, n_out, perm, wmax, nullptr))
#define B200_SELL_MAIN_U(W)                  \
    do {                                     \
        if (u == 4) B200_SELL_MAIN(W, 4);    \
        else if (u == 2) B200_SELL_MAIN(W, 2); \
        else B200_SELL_MAIN(W, 1);           \
    } while (0)
        switch (wpc) {
        case 8: B200_SELL_MAIN_U(8); break;
        case 4: B200_SELL_MAIN_U(4); break;
        case 2: B200_SELL_MAIN_U(2); break;
        default: B200_SELL_MAIN_U(1); break;
        }
#undef B200_SELL_MAIN_U
#undef B200_SELL_MAIN
        if (wmax > 0)
            sell32_kernel<T, P, true, 1, 2><<<ceil_div_u((long long)plan->n_items * 32, kBlock), kBlock, 0, ctx->stream>>>(
                data, idx, x, y, slice_ptr, plan->n_items, n_out, perm, wmax, plan->items);
    } else {
        sell32_scalar_kernel<T, P><<<blocks, kBlock, 0, ctx->stream>>>(data, idx, x, y, slice_ptr, n_slices, n_out, perm);
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T, int KS>
int launch_ellcm(b200_ctx *ctx, const T *data, const int *idx, const T *x, T *y, int n_rows,
                 int row_size, int pitch)
{
    // KS == 1 (enough rows): 256 threads x Q quads of rows.  Q = 2 (8 KiB contiguous per column
    // and array in fp32) measured no better than Q = 1 on B200 (0.198 vs 0.192 ms on the banded
    // workload); kept as a tuning hook: B200_ELLCM_Q=1|2
    constexpr int BX = KS == 1 ? 256 : 64;
    dim3 block(BX, KS);
    int q = 1;
    if (KS == 1 && opt_or(ctx, OPT_ELLCM_Q, 1) == 2) q = 2;
    unsigned blocks = ceil_div_u(pitch / 4, BX * q);
    if (q == 2)
        ellcm_kernel<T, KS, BX, (KS == 1 ? 2 : 1)><<<blocks, block, 0, ctx->stream>>>(data, idx, x, y, n_rows, row_size, pitch);
    else
        ellcm_kernel<T, KS, BX, 1><<<blocks, block, 0, ctx->stream>>>(data, idx, x, y, n_rows, row_size, pitch);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int spmv_ellcm_impl(b200_ctx *ctx, const T *data, const int *idx, const T *x, T *y, int n_rows,
                    int row_size, int pitch)
{
    B200_ENTER(ctx);
    B200_REQUIRE(x && y && n_rows >= 0 && row_size >= 0, "bad argument");
    B200_REQUIRE(pitch >= n_rows && pitch % 32 == 0, "pitch must be a multiple of 32 and >= n_rows");
    if (n_rows == 0) return B200_SUCCESS;
    B200_REQUIRE(row_size == 0 || (data && idx), "null data/indices");
    B200_REQUIRE(aligned16(data) && aligned16(idx), "column-major ELL arrays must be 16-byte aligned");
    // enough threads for ~2 full waves: split the K columns over KS slices when rows are few
    long long quads = pitch / 4;
    long long want = (long long)ctx->sm_count * 2048;
    int ks = 1;
    while (ks < 16 && quads * ks < want && ks * 2 <= row_size) ks <<= 1;
    switch (ks) {
    case 1: return launch_ellcm<T, 1>(ctx, data, idx, x, y, n_rows, row_size, pitch);
    case 2: return launch_ellcm<T, 2>(ctx, data, idx, x, y, n_rows, row_size, pitch);
    case 4: return launch_ellcm<T, 4>(ctx, data, idx, x, y, n_rows, row_size, pitch);
    case 8: return launch_ellcm<T, 8>(ctx, data, idx, x, y, n_rows, row_size, pitch);
    default: return launch_ellcm<T, 16>(ctx, data, idx, x, y, n_rows, row_size, pitch);
    }
}

}  // namespace

static int sell_exchange_impl(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                              const int *row_indices, int chunk, int n_slices, int n_rows,
                              const double *scale_sumsq, double *sumsq_out, double *const *dst, int n_dst,
                              long long dst_offset, const int *dst_row_lo, const int *dst_row_hi,
                              void *const *sync_blocks, int my_rank, unsigned long long step)
{
    B200_ENTER(ctx);
    B200_REQUIRE(vect && row_indices && dst && n_slices >= 0 && n_rows >= 0 && dst_offset >= 0, "bad argument");
    B200_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers, "n_dst must be in 1..16");
    B200_REQUIRE((long long)n_rows <= (long long)n_slices * 32, "n_rows exceeds n_slices*32");
    if (chunk != 32) {
        b200_set_error("SELL chunk must be 32 (warp-aligned), got %d", chunk);
        return B200_ERR_UNSUPPORTED;
    }
    B200_REQUIRE(aligned16(data) && aligned16(indices), "SELL arrays must be 16-byte aligned");
    B200_REQUIRE(!sync_blocks || n_slices > 0, "ring exchange needs at least one chunk per rank");
    if (n_slices == 0) return B200_SUCCESS;
    PeerDst<double> d;
    B200_REQUIRE((dst_row_lo == nullptr) == (dst_row_hi == nullptr), "give both row-range arrays or neither");
    for (int i = 0; i < kMaxPeers; ++i) {
        d.p[i] = i < n_dst ? dst[i] : nullptr;
        d.lo[i] = (i < n_dst && dst_row_lo) ? dst_row_lo[i] : 0;
        d.hi[i] = i < n_dst ? (dst_row_hi ? dst_row_hi[i] : n_rows) : 0;
    }
    for (int i = 0; i < n_dst; ++i) B200_REQUIRE(d.p[i], "null destination buffer");
    // rows are stored as 8-byte words (lane = row, 256 contiguous bytes per warp): any double-aligned
    // destination and any offset are fine
    for (int i = 0; i < n_dst; ++i)
        B200_REQUIRE((reinterpret_cast<uintptr_t>(d.p[i]) & 7) == 0, "destination buffers must be 8-byte aligned");
    for (int i = 0; i < n_dst; ++i) B200_REQUIRE(d.lo[i] >= 0, "negative row range");
    PeerSync sync;
    memset(&sync, 0, sizeof sync);
    const unsigned grid = ceil_div_u((long long)n_slices * 32, kBlock);
    // B200_BCAST_U = 2 / 3 / 4: the persistent pipelined kernel at that many blocks per SM; 1: one warp per
    // chunk, one chunk per warp.  Default: pipelined when the launch has at least 4 chunks per resident warp
    // (7-point Laplacian, 8 M rows: 0.151 -> 0.122 ms at 3 blocks per SM, 0.140 at 2, 0.172 at 4 -- spills).
    int pipe = opt_or(ctx, OPT_BCAST_U, 0);
    if (pipe == 0) pipe = (long long)n_slices >= 4ll * ctx->sm_count * kPipeDefaultBlocks * (kBlock / 32) ? kPipeDefaultBlocks : 1;
    unsigned pgrid = 0;
    void (*pipe_kernel)(const double *, const int *, const double *, const int *, int, int, const double *, double *,
                        PeerDst<double>, int, long long, double *, const int *) = nullptr;
    if (pipe >= 2) {
        pipe_kernel = pipe == 2   ? sell32_pipe_kernel<double, int, 2, true>
                      : pipe == 3 ? sell32_pipe_kernel<double, int, 3, true>
                                  : sell32_pipe_kernel<double, int, 4, true>;
        int per_sm = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pipe_kernel, kBlock, 0));
        if (per_sm < 1) per_sm = 1;
        pgrid = (unsigned)ctx->sm_count * (unsigned)per_sm;
        if (pgrid > grid) pgrid = grid;
    }
    auto launch_fused = [&](const double *scale, double *sums) {
        if (pipe_kernel)
            pipe_kernel<<<pgrid, kBlock, 0, ctx->stream>>>(data, indices, vect, row_indices, n_slices, n_rows, scale, sums, d,
                                                           n_dst, dst_offset, nullptr, nullptr);
        else
            sell32_bcast_kernel<double, int><<<grid, kBlock, 0, ctx->stream>>>(data, indices, vect, row_indices, n_slices,
                                                                               n_rows, scale, sums, d, n_dst, dst_offset);
    };
    if (sync_blocks) {
        for (int i = 0; i < n_dst; ++i) {
            B200_REQUIRE(sync_blocks[i], "null sync block");
            sync.blk[i] = static_cast<unsigned long long *>(sync_blocks[i]);
        }
        sync.mine = sync.blk[my_rank];
        sync.my_rank = my_rank;
        sync.world = n_dst;
        sync.step = step;
        sync.err_flag = ctx->scratch + kWatchFlag;
        sync.flush_after_flag = opt_or(ctx, OPT_RING_FLUSH, 1) != 0;
        sync.poll_mode = opt_or(ctx, OPT_RING_POLL, 0) != 0;
        sync.sleep_ns = opt_or(ctx, OPT_RING_SLEEP_NS, 100);
        double *acc = reinterpret_cast<double *>(sync.mine + kSyncAcc);
        const double *scale = step > 0 ? reinterpret_cast<const double *>(sync.mine + kSyncScale) : nullptr;
        launch_fused(scale, acc);
        ring_sync_kernel<<<1, 32, 0, ctx->stream>>>(sync);
        ctx->watch_flag = true;
    } else {
        launch_fused(scale_sumsq, sumsq_out);
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}


extern "C" {

int b200_sell_plan_create(b200_ctx *ctx, const int *row_indices, int n_slices, b200_sell_plan **plan)
{
    return sell_plan_create_impl<int>(ctx, row_indices, n_slices, plan);
}
int b200_sell64_plan_create(b200_ctx *ctx, const long long *slice_ptr, int n_slices, b200_sell_plan **plan)
{
    return sell_plan_create_impl<long long>(ctx, slice_ptr, n_slices, plan);
}
int b200_sell_plan_extra_items(const b200_sell_plan *plan, int *n_items)
{
    B200_REQUIRE(plan && n_items, "null argument");
    *n_items = plan->n_items;
    return B200_SUCCESS;
}
int b200_sell_plan_destroy(b200_sell_plan *plan)
{
    if (!plan) return B200_SUCCESS;
    cudaSetDevice(plan->device);
    if (plan->items) cudaFree(plan->items);
    delete plan;
    return B200_SUCCESS;
}

int b200_spmv_sell_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                       double *output, const int *row_indices, int chunk, int n_slices, int n_out,
                       const int *perm, const b200_sell_plan *plan)
{
    return spmv_sell_impl<double, int>(ctx, data, indices, vect, output, row_indices, chunk, n_slices, n_out, perm, plan);
}
int b200_spmv_sell_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                       float *output, const int *row_indices, int chunk, int n_slices, int n_out,
                       const int *perm, const b200_sell_plan *plan)
{
    return spmv_sell_impl<float, int>(ctx, data, indices, vect, output, row_indices, chunk, n_slices, n_out, perm, plan);
}
int b200_spmv_sell64_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                         double *output, const long long *slice_ptr, int chunk, int n_slices,
                         int n_out, const int *perm, const b200_sell_plan *plan)
{
    return spmv_sell_impl<double, long long>(ctx, data, indices, vect, output, slice_ptr, chunk, n_slices, n_out, perm, plan);
}
int b200_spmv_sell64_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                         float *output, const long long *slice_ptr, int chunk, int n_slices,
                         int n_out, const int *perm, const b200_sell_plan *plan)
{
    return spmv_sell_impl<float, long long>(ctx, data, indices, vect, output, slice_ptr, chunk, n_slices, n_out, perm, plan);
}

int b200_spmv_sell_bcast_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                             const int *row_indices, int chunk, int n_slices, int n_rows,
                             const double *scale_sumsq, double *sumsq_out, double *const *dst, int n_dst,
                             long long dst_offset)
{
    return b200_spmv_sell_halo_f64(ctx, data, indices, vect, row_indices, chunk, n_slices, n_rows, scale_sumsq,
                                   sumsq_out, dst, n_dst, dst_offset, nullptr, nullptr);
}

int b200_spmv_sell_halo_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                            const int *row_indices, int chunk, int n_slices, int n_rows,
                            const double *scale_sumsq, double *sumsq_out, double *const *dst, int n_dst,
                            long long dst_offset, const int *dst_row_lo, const int *dst_row_hi)
{
    return sell_exchange_impl(ctx, data, indices, vect, row_indices, chunk, n_slices, n_rows, scale_sumsq, sumsq_out,
                              dst, n_dst, dst_offset, dst_row_lo, dst_row_hi, nullptr, 0, 0);
}

int b200_spmv_sell_ring_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                            const int *row_indices, int chunk, int n_slices, int n_rows, double *const *dst,
                            int n_dst, long long dst_offset, const int *dst_row_lo, const int *dst_row_hi,
                            void *const *sync_blocks, int my_rank, unsigned long long step)
{
    B200_REQUIRE(sync_blocks && my_rank >= 0 && my_rank < n_dst, "bad sync arguments");
    return sell_exchange_impl(ctx, data, indices, vect, row_indices, chunk, n_slices, n_rows, nullptr, nullptr, dst,
                              n_dst, dst_offset, dst_row_lo, dst_row_hi, sync_blocks, my_rank, step);
}

int b200_ipc_get_handle(b200_ctx *ctx, void *dptr, unsigned char handle[64])
{
    B200_ENTER(ctx);
    B200_REQUIRE(dptr && handle, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    B200_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return B200_SUCCESS;
}

int b200_ipc_open_handle(b200_ctx *ctx, const unsigned char handle[64], void **peer_dptr)
{
    B200_ENTER(ctx);
    B200_REQUIRE(handle && peer_dptr, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    *peer_dptr = nullptr;
    B200_CUDA(cudaIpcOpenMemHandle(peer_dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200_SUCCESS;
}

int b200_ipc_close_handle(b200_ctx *ctx, void *peer_dptr)
{
    B200_ENTER(ctx);
    if (peer_dptr) B200_CUDA(cudaIpcCloseMemHandle(peer_dptr));
    return B200_SUCCESS;
}

int b200_spmv_ellcm_f64(b200_ctx *ctx, const double *data_cm, const int *indices_cm,
                        const double *vect, double *output, int n_rows, int row_size, int pitch)
{
    return spmv_ellcm_impl<double>(ctx, data_cm, indices_cm, vect, output, n_rows, row_size, pitch);
}
int b200_spmv_ellcm_f32(b200_ctx *ctx, const float *data_cm, const int *indices_cm,
                        const float *vect, float *output, int n_rows, int row_size, int pitch)
{
    return spmv_ellcm_impl<float>(ctx, data_cm, indices_cm, vect, output, n_rows, row_size, pitch);
}

}  // extern "C"
