// spmv_csr.cu -- CSR and row-major ELL SpMV for sm_100a.
//
// Replaces kernels/Csr.cl:1-17 (scalar thread-per-row, 32 work-groups) and kernels/Ell.cl:1-39
// (16-lane work-group per row with a local-memory tree).  B200 design:
//   * LPR (2..32) lanes cooperate on one row; LPR is picked from the row-length statistics so
//     that a lane walks ~4 groups of four entries (measured: the x gather's L1 wavefronts, not the
//     matrix stream, are the limit, so a warp should cover ADJACENT rows);
//   * indices and values are fetched as 128-bit loads on 16-byte ADDRESS-aligned groups of four
//     entries; the ragged first/last group of a row is masked, never re-read;
//   * U groups per lane are loaded before any x is gathered (explicit batches, batch_hold());
//   * matrix data streams with evict-first, x goes through the read-only path (coherent loads in
//     the launch-overlap variants, OVL = true);
//   * rows far longer than the mean (power-law inputs) are skipped by the vector kernel and
//     handled by a block-per-row kernel that splits very long rows across blocks (one atomic
//     per block); short-row matrices take the nnz-split stream kernel instead;
//   * reduction is __shfl_xor_sync only -- no shared memory, no barriers.
// HBM-bound: algorithmic bytes = nnz*(4+V) + (R+1)*4 + Cn*V + R*V.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;

// dot product of entries [s, e) of (col, data) with x, cooperatively by LPR lanes.
// One round trip = U aligned groups of four entries per lane: first ALL 2U (fp32) / 3U (fp64)
// 128-bit matrix loads are issued, then all 4U x gathers, then the FMAs (two accumulators).  Written
// as three separate phases over register arrays because a plain `#pragma unroll U` loop compiles to
// load -> gather -> FMA -> load ... (cuobjdump -sass), i.e. one round trip per group: fine when
// dozens of warps per SM hide it, a latency chain on matrices of one or two waves (cant).
// U = 1 keeps the register count at 32 (full occupancy); U = 2 / 4 trade occupancy for loads in flight.
template <typename T, int LPR, int U, bool OVL>
__device__ __forceinline__ T row_dot_vec(const int *__restrict__ col, const T *__restrict__ data,
                                         const T *__restrict__ x, long long s, long long e, int lane,
                                         bool &waited)
{
    T acc0 = 0, acc1 = 0;
    const long long g_end = (e + 3) >> 2;
    for (long long g0 = (s >> 2) + lane; g0 < g_end; g0 += (long long)LPR * U) {
        IVec4 c[U];
        Vec4<T> v[U];
        T xv[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long g = g0 + (long long)u * LPR;
            c[u].zero();
            v[u].zero();
            if (g < g_end) {
                c[u].load(col + (g << 2));
                v[u].load(data + (g << 2));
            }
        }
        const int hold = batch_hold<U, T>(c, v);  // 0; orders the gathers after ALL loads (common.cuh)
        pdl_wait_once<OVL>(waited);  // x may still be being written by the previous launch (common.cuh)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = (g0 + (long long)u * LPR) << 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // j + k < e implies the group was loaded; masked slots contribute exactly +0
                const bool ok = j + k >= s && j + k < e;
                xv[u][k] = ok ? ld_xo<OVL>(x, c[u].v[k] + hold) : T(0);
                if (!ok) v[u].v[k] = T(0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc0 += v[u].v[0] * xv[u][0];
            acc1 += v[u].v[1] * xv[u][1];
            acc0 += v[u].v[2] * xv[u][2];
            acc1 += v[u].v[3] * xv[u][3];
        }
    }
    return acc0 + acc1;
}

// same, scalar loads: used only when an array is not 16-byte aligned
template <typename T, int LPR, bool OVL>
__device__ __forceinline__ T row_dot_scalar(const int *__restrict__ col, const T *__restrict__ data,
                                            const T *__restrict__ x, long long s, long long e, int lane,
                                            bool &waited)
{
    T acc = 0;
    pdl_wait_once<OVL>(waited);
    for (long long j = s + lane; j < e; j += LPR) acc += ld_stream(data + j) * ld_xo<OVL>(x, ld_stream(col + j));
    return acc;
}

template <typename T, int LPR, bool VEC, int U, bool OVL>
__global__ void __launch_bounds__(kBlock)
csr_vector_kernel(const int *__restrict__ ptr, const int *__restrict__ col, const T *__restrict__ data,
                  const T *__restrict__ x, T *__restrict__ y, int n_rows, int long_threshold)
{
    pdl_launch_dependents();
    bool waited = false;
    const int lane = threadIdx.x & (LPR - 1);
    const long long row = ((long long)blockIdx.x * kBlock + threadIdx.x) / LPR;
    int s = 0, e = 0;
    if (row < n_rows) {
        s = __ldg(ptr + row);
        e = __ldg(ptr + row + 1);
    }
    const bool is_long = (e - s) > long_threshold;
    T acc = 0;
    if (!is_long)
        acc = VEC ? row_dot_vec<T, LPR, U, OVL>(col, data, x, s, e, lane, waited)
                  : row_dot_scalar<T, LPR, OVL>(col, data, x, s, e, lane, waited);
    pdl_wait_once<OVL>(waited);  // rows without entries never waited: y must not be written early either
    acc = subwarp_sum<LPR>(acc);
    // long rows: publish 0 here; csr_long_rows_kernel accumulates into it afterwards
    if (lane == 0 && row < n_rows) y[row] = is_long ? T(0) : acc;
}

// one (row, split) pair per block; gridDim = (n_long, n_split)
template <typename T, bool VEC>
__global__ void __launch_bounds__(kBlock)
csr_long_rows_kernel(const int *__restrict__ ptr, const int *__restrict__ col,
                     const T *__restrict__ data, const T *__restrict__ x, T *__restrict__ y,
                     const int *__restrict__ long_rows)
{
    __shared__ T warp_part[kBlock / 32];
    const int row = long_rows[blockIdx.x];
    const long long s0 = ptr[row], e0 = ptr[row + 1];
    // split [s0,e0) into gridDim.y pieces on 4-entry boundaries
    long long per = ((e0 - s0 + gridDim.y - 1) / gridDim.y + 3) & ~3ll;
    long long s = s0 + per * blockIdx.y;
    long long e = s + per < e0 ? s + per : e0;
    T acc = 0;
    if (s < e) {
        if (VEC) {
            const long long g_end = (e + 3) >> 2;
#pragma unroll 2
            for (long long g = (s >> 2) + threadIdx.x; g < g_end; g += kBlock) {
                const long long j = g << 2;
                IVec4 c;
                Vec4<T> v;
                c.load(col + j);
                v.load(data + j);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (j + k >= s && j + k < e) acc += v.v[k] * ld_x(x, c.v[k]);
            }
        } else {
            for (long long j = s + threadIdx.x; j < e; j += kBlock)
                acc += ld_stream(data + j) * ld_x(x, ld_stream(col + j));
        }
    }
    acc = subwarp_sum<32>(acc);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        T t = threadIdx.x < kBlock / 32 ? warp_part[threadIdx.x] : T(0);
        t = subwarp_sum<32>(t);
        if (threadIdx.x == 0 && s < e) atomicAdd(y + row, t);
    }
}

// ---- short rows: nnz-split "stream" kernel -------------------------------------------------
// When the mean row has only a few entries, lanes-per-row kernels waste lanes and re-touch the
// ragged 16-byte groups of neighbouring rows.  Here a block owns a fixed tile of kTile consecutive
// ENTRIES instead: every thread loads one aligned group of four (perfectly coalesced, nothing
// masked but the matrix tail), the products go to shared memory, and then one thread per row sums
// that row's slice of the tile.  Rows [tile_lo[b], tile_lo[b+1]) START in tile b and are stored by
// it (a partial sum if the row runs on into later tiles); the piece of a row that started in an
// earlier tile is a carry, added by csr_stream_fixup_kernel afterwards (same stream, so ordered).
// tile_lo[] is part of the plan (one binary search per tile, done once).
constexpr int kTile = 1024;  // entries per 256-thread block and group (a block owns G * kTile entries)

__global__ void csr_tile_rows_kernel(const int *__restrict__ ptr, int n_rows, int n_tiles, int tile,
                                     int *__restrict__ tile_lo)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_tiles) return;
    if (b == n_tiles) {
        tile_lo[b] = n_rows;  // the last tile also owns trailing empty rows
        return;
    }
    const int e0 = b * tile;
    int lo = 0, hi = n_rows;  // smallest r in [0, n_rows] with ptr[r] >= e0
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (ptr[mid] < e0) lo = mid + 1;
        else hi = mid;
    }
    tile_lo[b] = lo;
}

// G = groups of four entries per thread (tile = G * 1024 entries): all G index/value loads of a thread
// are issued before its gathers, and one barrier pair serves G times the entries (on power-law inputs
// the kernel is bound by the gather's L1/L2 sector traffic and the barrier wait showed up as its
// largest single stall, profiles/r2_ncu_summary.md)
template <typename T, bool VEC, int G>
__global__ void __launch_bounds__(kBlock)
csr_stream_kernel(const int *__restrict__ ptr, const int *__restrict__ col, const T *__restrict__ data,
                  const T *__restrict__ x, T *__restrict__ y, int nnz, const int *__restrict__ tile_lo,
                  int *__restrict__ carry_row, T *__restrict__ carry_val)
{
    constexpr int TILE = kTile * G;
    __shared__ T prod[TILE];
    __shared__ int long_list[TILE / 32];  // rows with > 32 entries inside this tile (fewer than TILE / 32)
    __shared__ int n_long;
    if (threadIdx.x == 0) n_long = 0;
    const int b = blockIdx.x;
    const int e0 = b * TILE, e1 = min(e0 + TILE, nnz);
    const int lo = __ldg(tile_lo + b), r_end = __ldg(tile_lo + b + 1);
    T p[G][4];
    if (VEC) {
        IVec4 c[G];
        Vec4<T> v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int j = e0 + 4 * (threadIdx.x + kBlock * g);
            c[g].zero();
            v[g].zero();
            if (j < e1) {
                c[g].load(col + j);
                v[g].load(data + j);
            }
        }
        const int hold = batch_hold<G, T>(c, v);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int j = e0 + 4 * (threadIdx.x + kBlock * g);
#pragma unroll
            for (int k = 0; k < 4; ++k) p[g][k] = (j + k < e1) ? v[g].v[k] * ld_x(x, c[g].v[k] + hold) : T(0);
        }
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int j = e0 + 4 * (threadIdx.x + kBlock * g);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                p[g][k] = (j + k < e1) ? ld_stream(data + j + k) * ld_x(x, ld_stream(col + j + k)) : T(0);
        }
    }
    // this thread's first row: pointers fetched before the barrier so the loads overlap the gather
    int r = lo + threadIdx.x;
    int s = 0, e = 0;
    if (r < r_end) {
        s = __ldg(ptr + r);
        e = __ldg(ptr + r + 1);
    }
    const int first_start = __ldg(ptr + lo);  // lo <= n_rows: ptr has n_rows + 1 entries
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k) prod[4 * (threadIdx.x + kBlock * g) + k] = p[g][k];
    __syncthreads();

    // carry: entries [e0, first_start) belong to row lo-1, which started in an earlier tile
    if (threadIdx.x < 32) {
        const int c_end = min(first_start, e1) - e0;
        if (c_end > 0) {
            T part = 0;
            for (int i = threadIdx.x; i < c_end; i += 32) part += prod[i];
            part = subwarp_sum<32>(part);
            if (threadIdx.x == 0) {
                carry_row[b] = lo - 1;
                carry_val[b] = part;
            }
        } else if (threadIdx.x == 0) {
            carry_row[b] = -1;
        }
    }
    // rows that start in this tile: one thread per row; a row with more than 32 entries inside
    // the tile is left to a whole warp below (keeps skewed, power-law inputs balanced)
    while (r < r_end) {
        const int stop = min(e, e1) - e0;
        if (stop - (s - e0) > 32) {
            long_list[atomicAdd(&n_long, 1)] = r;
        } else {
            T acc = 0;
            for (int i = s - e0; i < stop; ++i) acc += prod[i];
            y[r] = acc;
        }
        r += kBlock;
        if (r < r_end) {
            s = __ldg(ptr + r);
            e = __ldg(ptr + r + 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x >> 5; i < n_long; i += kBlock / 32) {
        const int lr = long_list[i];
        const int ls = __ldg(ptr + lr) - e0, lstop = min(__ldg(ptr + lr + 1), e1) - e0;
        T part = 0;
        for (int k = ls + (threadIdx.x & 31); k < lstop; k += 32) part += prod[k];
        part = subwarp_sum<32>(part);
        if ((threadIdx.x & 31) == 0) y[lr] = part;
    }
}

template <typename T>
__global__ void csr_stream_fixup_kernel(T *__restrict__ y, int n_tiles, const int *__restrict__ carry_row,
                                        const T *__restrict__ carry_val)
{
    // A row's carry tiles are consecutive.  The thread of the FIRST of them adds the whole run in
    // tile order: no atomics, and the result does not depend on scheduling (bit-reproducible).
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_tiles) return;
    const int r = carry_row[b];
    if (r < 0 || (b > 0 && carry_row[b - 1] == r)) return;
    T sum = 0;
    for (int k = b; k < n_tiles && carry_row[k] == r; ++k) sum += carry_val[k];
    y[r] += sum;
}

// row-major ELL (the reference's arrays): row r owns entries [r*K, (r+1)*K)
template <typename T, int LPR, bool VEC, int U, bool OVL>
__global__ void __launch_bounds__(kBlock)
ell_rowmajor_kernel(const T *__restrict__ data, const int *__restrict__ col, const T *__restrict__ x,
                    T *__restrict__ y, int n_rows, int row_size)
{
    pdl_launch_dependents();
    bool waited = false;
    const int lane = threadIdx.x & (LPR - 1);
    const long long row = ((long long)blockIdx.x * kBlock + threadIdx.x) / LPR;
    long long s = 0, e = 0;
    if (row < n_rows) {
        s = row * row_size;
        e = s + row_size;
    }
    T acc = VEC ? row_dot_vec<T, LPR, U, OVL>(col, data, x, s, e, lane, waited)
                : row_dot_scalar<T, LPR, OVL>(col, data, x, s, e, lane, waited);
    pdl_wait_once<OVL>(waited);
    acc = subwarp_sum<LPR>(acc);
    if (lane == 0 && row < n_rows) y[row] = acc;
}

// ---- plan: row-length statistics on the device --------------------------------------------
struct PlanStats {
    int min_len, max_len, n_long, pad;
};

__global__ void csr_stats_kernel(const int *__restrict__ ptr, int n_rows, int long_threshold,
                                 PlanStats *out)
{
    int lo = 0x7fffffff, hi = 0, cnt = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (long long)gridDim.x * blockDim.x) {
        int len = ptr[r + 1] - ptr[r];
        lo = min(lo, len);
        hi = max(hi, len);
        cnt += len > long_threshold;
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out->min_len, lo);
        atomicMax(&out->max_len, hi);
        if (cnt) atomicAdd(&out->n_long, cnt);
    }
}

__global__ void csr_collect_long_kernel(const int *__restrict__ ptr, int n_rows, int long_threshold,
                                        int *counter, int *list)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (long long)gridDim.x * blockDim.x) {
        if (ptr[r + 1] - ptr[r] > long_threshold) list[atomicAdd(counter, 1)] = (int)r;
    }
}

int pick_lanes(const b200_ctx *ctx, double mean_len, b200_opt hook)
{
    // Measured on B200 (profiles/r1_lanes_sweep.md): the kernel is bound by L1 wavefronts of the x
    // gather, not by the matrix loads.  Fewer lanes per row put more ADJACENT rows in one warp, and
    // adjacent rows gather neighbouring x entries, so the sweet spot is ~4 vector iterations
    // (16 entries) per lane -- but never 2 lanes on multi-iteration rows, whose 32-byte row
    // segments would uncoalesce the matrix loads themselves.
    int lanes = 2;
    while (lanes < 32 && lanes * 16 < mean_len) lanes <<= 1;
    if (lanes == 2 && mean_len > 8.0) lanes = 4;
    // tuning hook: B200_CSR_LANES / B200_ELL_LANES = 2,4,8,16,32 overrides the heuristic
    const int v = opt_or(ctx, hook, 0);
    if (v == 2 || v == 4 || v == 8 || v == 16 || v == 32) lanes = v;
    return lanes;
}

// groups loaded per lane and round trip (row_dot_vec); tuning hook <env> = 1|2|4.
// Measured on B200 (profiles/r1e_variant_sweep.md).  Large launches, sustained 200-step runs (the GPU
// sits at its power cap there; a 10-launch burst shows no difference between the depths):
//   fp32: 4 (CSR 0.176 vs 0.179 ms, ELL 0.167 vs 0.172);  fp64: 2 (CSR 0.261 vs 0.266 at 1 and 0.326
//   at 4 -- 96 registers; ELL 0.252 / 0.255 / 0.289).
// Launches of at most ~2 waves (cant): CSR 2; ELL fp32 4, fp64 1.
int pick_unroll(const b200_ctx *ctx, long long threads, int value_bytes, bool ell, b200_opt hook)
{
    const bool small = threads <= 2ll * ctx->sm_count * 2048;
    int u;
    if (ell) u = value_bytes == 4 ? 4 : (small ? 1 : 2);
    else u = (value_bytes == 4 && !small) ? 4 : 2;
    const int v = opt_or(ctx, hook, 0);
    if (v == 1 || v == 2 || v == 4) u = v;
    return u;
}

}  // namespace

struct b200_csr_plan {
    b200_csr_plan_info info;
    int device;
    int n_split;      // blocks per long row
    int *long_rows;   // device list, n_long_rows entries
    // stream (nnz-split) variant for short rows: info.stream_tiles > 0
    int *tile_lo;     // stream_tiles + 1 entries
    int *carry_row;   // stream_tiles entries
    void *carry_val;  // stream_tiles x 8 bytes (T = float or double)
    int stream_groups; // G of csr_stream_kernel: a tile holds G * 1024 entries
    cudaStream_t owner;  // a stream plan's carry buffers serve ONE queue: the one it was created on
};

extern "C" {

// fills *p; on any failure the caller destroys p (nothing leaks on the early returns)
static int csr_plan_fill(b200_ctx *ctx, const int *ptr, int n_rows, b200_csr_plan *p)
{
    int first_last[2] = {0, 0};
    if (n_rows > 0) {
        B200_CUDA(cudaMemcpyAsync(&first_last[0], ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA(cudaMemcpyAsync(&first_last[1], ptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    b200_csr_plan_info &in = p->info;
    in.stream_tiles = 0;
    in.stream_tile_entries = 0;
    in.n_rows = n_rows;
    in.nnz = (long long)first_last[1] - first_last[0];
    in.mean_len = n_rows > 0 ? (double)in.nnz / n_rows : 0.0;
    in.lanes_per_row = pick_lanes(ctx, in.mean_len, OPT_CSR_LANES);
    // a row is "long" when it would keep its lanes busy for more than 64 vector iterations
    // (and is worth a whole block): it goes to the block-per-row kernel instead
    in.long_threshold = in.lanes_per_row * 4 * 64 > 1024 ? in.lanes_per_row * 4 * 64 : 1024;
    in.min_len = in.max_len = 0;
    in.n_long_rows = 0;
    if (n_rows == 0) return B200_SUCCESS;
    PlanStats init = {0x7fffffff, 0, 0, 0};
    PlanStats *d = reinterpret_cast<PlanStats *>(ctx->scratch);
    B200_CUDA(cudaMemcpyAsync(d, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    int blocks = (int)min((long long)ctx->sm_count * 8, ((long long)n_rows + 255) / 256);
    csr_stats_kernel<<<blocks, 256, 0, ctx->stream>>>(ptr, n_rows, in.long_threshold, d);
    B200_LAUNCH_CHECK();
    PlanStats got;
    B200_CUDA(cudaMemcpyAsync(&got, d, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    in.min_len = got.min_len;
    in.max_len = got.max_len;
    in.n_long_rows = got.n_long;
    if (got.n_long > 0) {
        B200_CUDA(cudaMalloc(&p->long_rows, sizeof(int) * (size_t)got.n_long));
        int *counter = ctx->scratch + 16;
        B200_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
        csr_collect_long_kernel<<<blocks, 256, 0, ctx->stream>>>(ptr, n_rows, in.long_threshold, counter, p->long_rows);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        // one block streams ~64 Ki entries; longer rows are split (capped)
        long long split = ((long long)in.max_len + 65535) / 65536;
        p->n_split = (int)(split < 1 ? 1 : (split > 128 ? 128 : split));
    }
    // short rows everywhere: the nnz-split stream kernel (B200_CSR_STREAM=0|1 overrides)
    // (any maximum row length: rows longer than 32 entries inside a tile get a whole warp, rows
    // longer than a tile are stitched by the carries -- the split is by entries, so skewed
    // power-law inputs stay balanced)
    const bool possible = first_last[0] == 0 && in.nnz > 0 && in.nnz < 0x7fffffffll - 4 * kTile;
    const bool short_rows = in.mean_len <= 16.0;
    const bool skewed = in.mean_len <= 32.0 && (double)in.max_len > 16.0 * in.mean_len;
    bool stream = (short_rows || skewed) && possible;
    if (opt_set(ctx, OPT_CSR_STREAM)) stream = ctx->opt[OPT_CSR_STREAM] != 0 && possible;
    if (stream) {
        // groups per thread (tuning hook B200_CSR_STREAM_G=1|2|4): 1 for stencil-like rows, 2 when the
        // rows are skewed (the gather-bound power-law case)
        int g = skewed ? 2 : 1;
        {
            const int v = opt_or(ctx, OPT_CSR_STREAM_G, 0);
            if (v == 1 || v == 2 || v == 4) g = v;
        }
        p->stream_groups = g;
        const int tile = kTile * g;
        const int n_tiles = (int)((in.nnz + tile - 1) / tile);
        B200_CUDA(cudaMalloc(&p->tile_lo, sizeof(int) * ((size_t)n_tiles + 1)));
        B200_CUDA(cudaMalloc(&p->carry_row, sizeof(int) * (size_t)n_tiles));
        B200_CUDA(cudaMalloc(&p->carry_val, sizeof(double) * (size_t)n_tiles));
        csr_tile_rows_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, ctx->stream>>>(ptr, n_rows, n_tiles, tile, p->tile_lo);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        in.stream_tiles = n_tiles;
        in.stream_tile_entries = tile;
    }
    return B200_SUCCESS;
}

int b200_csr_plan_create(b200_ctx *ctx, const int *ptr, int n_rows, b200_csr_plan **plan)
{
    B200_TRACE("b200 csr plan");
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && plan && n_rows >= 0, "bad argument");
    *plan = nullptr;
    b200_csr_plan *p = new b200_csr_plan();
    p->device = ctx->device;
    p->long_rows = nullptr;
    p->tile_lo = nullptr;
    p->carry_row = nullptr;
    p->carry_val = nullptr;
    p->n_split = 1;
    p->stream_groups = 1;
    p->owner = ctx->stream;
    const int rc = csr_plan_fill(ctx, ptr, n_rows, p);
    if (rc != B200_SUCCESS) {
        b200_csr_plan_destroy(p);
        return rc;
    }
    *plan = p;
    return B200_SUCCESS;
}

int b200_csr_plan_get_info(const b200_csr_plan *plan, b200_csr_plan_info *info)
{
    B200_REQUIRE(plan && info, "null argument");
    *info = plan->info;
    return B200_SUCCESS;
}

int b200_csr_plan_destroy(b200_csr_plan *plan)
{
    if (!plan) return B200_SUCCESS;
    cudaSetDevice(plan->device);
    if (plan->long_rows) cudaFree(plan->long_rows);
    if (plan->tile_lo) cudaFree(plan->tile_lo);
    if (plan->carry_row) cudaFree(plan->carry_row);
    if (plan->carry_val) cudaFree(plan->carry_val);
    delete plan;
    return B200_SUCCESS;
}

}  // extern "C"

namespace {

template <typename T, int LPR>
int launch_csr_lpr(b200_ctx *ctx, const int *ptr, const int *col, const T *data, const T *x, T *y,
                   int n_rows, int long_threshold, bool vec)
{
    unsigned blocks = ceil_div_u((long long)n_rows * LPR, kBlock);
    const bool ovl = ovl_on(ctx, (long long)n_rows * LPR);
    const int u = pick_unroll(ctx, (long long)n_rows * LPR, (int)sizeof(T), false, OPT_CSR_UNROLL);
    if (!vec)
        B200_CUDA(ovl ? b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, false, 1, true>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold)
                               : b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, false, 1, false>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold));
    else if (u == 4)
        B200_CUDA(ovl ? b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, true, 4, true>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold)
                               : b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, true, 4, false>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold));
    else if (u == 2)
        B200_CUDA(ovl ? b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, true, 2, true>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold)
                               : b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, true, 2, false>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold));
    else
        B200_CUDA(ovl ? b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, true, 1, true>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold)
                               : b200_launch(ctx, ovl, csr_vector_kernel<T, LPR, true, 1, false>, dim3(blocks), dim3(kBlock), 0, ptr, col, data, x, y, n_rows, long_threshold));
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int spmv_csr_impl(b200_ctx *ctx, const int *ptr, const int *col, const T *data, const T *x, T *y,
                  int n_rows, const b200_csr_plan *plan)
{
    B200_TRACE("b200 spmv csr");
    B200_ENTER_SPMV(ctx);
    B200_REQUIRE(ptr && x && y && n_rows >= 0, "bad argument");
    if (n_rows == 0) return B200_SUCCESS;
    b200_csr_plan *tmp = nullptr;
    if (!plan) {
        int rc = b200_csr_plan_create(ctx, ptr, n_rows, &tmp);
        if (rc) return rc;
        plan = tmp;
    }
    B200_REQUIRE(plan->info.n_rows == n_rows, "plan was built for a different matrix");
    B200_REQUIRE(plan->info.nnz == 0 || (col && data), "null col/data");
    const bool vec = aligned16(col) && aligned16(data);
    const int thr = plan->info.long_threshold;
    int rc;
    if (plan->info.stream_tiles > 0) {
        if (plan->owner != ctx->stream) {
            if (tmp) b200_csr_plan_destroy(tmp);
            b200_set_error("this CSR plan owns the carry buffers of the nnz-split kernel: use it on the "
                           "context it was created on (make one plan per queue)");
            return B200_ERR_INVALID_VALUE;
        }
        const int n_tiles = plan->info.stream_tiles;
        T *carry_val = static_cast<T *>(plan->carry_val);
#define B200_CSR_STREAM(V, GG)                                                                   \
    csr_stream_kernel<T, V, GG><<<n_tiles, kBlock, 0, ctx->stream>>>(ptr, col, data, x, y, (int)plan->info.nnz, \
                                                                    plan->tile_lo, plan->carry_row, carry_val)
        // the tile size is part of the plan (tile_lo): scalar-load and vector-load variants share it
        if (plan->stream_groups == 4) {
            if (vec) B200_CSR_STREAM(true, 4);
            else B200_CSR_STREAM(false, 4);
        } else if (plan->stream_groups == 2) {
            if (vec) B200_CSR_STREAM(true, 2);
            else B200_CSR_STREAM(false, 2);
        } else {
            if (vec) B200_CSR_STREAM(true, 1);
            else B200_CSR_STREAM(false, 1);
        }
#undef B200_CSR_STREAM
        csr_stream_fixup_kernel<T><<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(y, n_tiles, plan->carry_row, carry_val);
        cudaError_t e = cudaGetLastError();
        rc = e == cudaSuccess ? B200_SUCCESS : b200_cuda_fail(e, "csr_stream_kernel", __FILE__, __LINE__);
        if (tmp) {
            cudaStreamSynchronize(ctx->stream);
            b200_csr_plan_destroy(tmp);
        }
        return rc;
    }
    switch (plan->info.lanes_per_row) {
    case 2: rc = launch_csr_lpr<T, 2>(ctx, ptr, col, data, x, y, n_rows, thr, vec); break;
    case 4: rc = launch_csr_lpr<T, 4>(ctx, ptr, col, data, x, y, n_rows, thr, vec); break;
    case 8: rc = launch_csr_lpr<T, 8>(ctx, ptr, col, data, x, y, n_rows, thr, vec); break;
    case 16: rc = launch_csr_lpr<T, 16>(ctx, ptr, col, data, x, y, n_rows, thr, vec); break;
    default: rc = launch_csr_lpr<T, 32>(ctx, ptr, col, data, x, y, n_rows, thr, vec); break;
    }
    if (rc == B200_SUCCESS && plan->info.n_long_rows > 0) {
        dim3 grid(plan->info.n_long_rows, plan->n_split);
        if (vec)
            csr_long_rows_kernel<T, true><<<grid, kBlock, 0, ctx->stream>>>(ptr, col, data, x, y, plan->long_rows);
        else
            csr_long_rows_kernel<T, false><<<grid, kBlock, 0, ctx->stream>>>(ptr, col, data, x, y, plan->long_rows);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = b200_cuda_fail(e, "csr_long_rows_kernel", __FILE__, __LINE__);
    }
    if (tmp) {
        cudaStreamSynchronize(ctx->stream);
        b200_csr_plan_destroy(tmp);
    }
    return rc;
}

template <typename T, int LPR>
int launch_ell_lpr(b200_ctx *ctx, const T *data, const int *col, const T *x, T *y, int n_rows,
                   int row_size, bool vec)
{
    unsigned blocks = ceil_div_u((long long)n_rows * LPR, kBlock);
    const bool ovl = ovl_on(ctx, (long long)n_rows * LPR);
    const int u = pick_unroll(ctx, (long long)n_rows * LPR, (int)sizeof(T), true, OPT_ELL_UNROLL);
    if (!vec)
        B200_CUDA(ovl ? b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, false, 1, true>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size)
                               : b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, false, 1, false>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size));
    else if (u == 4)
        B200_CUDA(ovl ? b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, true, 4, true>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size)
                               : b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, true, 4, false>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size));
    else if (u == 2)
        B200_CUDA(ovl ? b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, true, 2, true>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size)
                               : b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, true, 2, false>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size));
    else
        B200_CUDA(ovl ? b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, true, 1, true>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size)
                               : b200_launch(ctx, ovl, ell_rowmajor_kernel<T, LPR, true, 1, false>, dim3(blocks), dim3(kBlock), 0, data, col, x, y, n_rows, row_size));
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int spmv_ell_impl(b200_ctx *ctx, const T *data, const int *col, const T *x, T *y, int n_rows,
                  int row_size)
{
    B200_TRACE("b200 spmv ell");
    B200_ENTER_SPMV(ctx);
    B200_REQUIRE(x && y && n_rows >= 0 && row_size >= 0, "bad argument");
    if (n_rows == 0) return B200_SUCCESS;
    B200_REQUIRE(row_size == 0 || (data && col), "null data/indices");
    const bool vec = aligned16(col) && aligned16(data);
    switch (pick_lanes(ctx, (double)row_size, OPT_ELL_LANES)) {
    case 2: return launch_ell_lpr<T, 2>(ctx, data, col, x, y, n_rows, row_size, vec);
    case 4: return launch_ell_lpr<T, 4>(ctx, data, col, x, y, n_rows, row_size, vec);
    case 8: return launch_ell_lpr<T, 8>(ctx, data, col, x, y, n_rows, row_size, vec);
    case 16: return launch_ell_lpr<T, 16>(ctx, data, col, x, y, n_rows, row_size, vec);
    default: return launch_ell_lpr<T, 32>(ctx, data, col, x, y, n_rows, row_size, vec);
    }
}

}  // namespace

extern "C" {

int b200_spmv_csr_f64(b200_ctx *ctx, const int *ptr, const int *col, const double *data,
                      const double *vect, double *output, int n_rows, const b200_csr_plan *plan)
{
    return spmv_csr_impl<double>(ctx, ptr, col, data, vect, output, n_rows, plan);
}

int b200_spmv_csr_f32(b200_ctx *ctx, const int *ptr, const int *col, const float *data,
                      const float *vect, float *output, int n_rows, const b200_csr_plan *plan)
{
    return spmv_csr_impl<float>(ctx, ptr, col, data, vect, output, n_rows, plan);
}

int b200_spmv_ell_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                      double *output, int n_rows, int row_size)
{
    return spmv_ell_impl<double>(ctx, data, indices, vect, output, n_rows, row_size);
}

int b200_spmv_ell_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                      float *output, int n_rows, int row_size)
{
    return spmv_ell_impl<float>(ctx, data, indices, vect, output, n_rows, row_size);
}

}  // extern "C"
