// spmv_cmrs_coo.cu -- CMRS and COO SpMV for sm_100a.
//
// Both formats are "entries + a row key per entry" and share one structure, chosen from the B200
// measurements in profiles/ (the x gather is bound by L1 wavefronts, so a warp should work on
// ADJACENT rows, and every per-entry instruction counts):
//
//   * a warp owns one contiguous run of entries (CMRS: one strip of `height` rows; COO: 512
//     entries) and splits it into 8 contiguous spans, one per 4-lane sub-warp.  A sub-warp reads
//     its span with 128-bit loads (64 contiguous bytes per instruction and array), so a warp
//     instruction covers 8 adjacent pieces of the matrix -- typically 8 adjacent rows, whose x
//     entries are neighbours;
//   * a lane keeps ONE running (row key, partial sum) pair in registers.  A group of four
//     entries whose keys all equal the running key costs 4 FMAs (the common case on row-sorted
//     data); only a key change flushes the pair;
//   * CMRS flushes into `height` per-lane REGISTER accumulators selected by compare (the
//     reference's per-entry read-modify-write of local memory, kernels/Cmrs.cl:18, was its
//     bottleneck); a transposing butterfly of __shfl_xor_sync (7 + 2 shuffles for height 8)
//     folds the 32 x height partials and lanes 0..height-1 store the strip's rows;
//   * COO flushes interior runs with one native atomicAdd and merges the lanes' final runs with a
//     ballot-delimited segmented warp scan, so row-sorted input costs about one atomic per
//     (warp, row) and arbitrary order (the column-major cant.mtx of coo.c:43) degrades to one per
//     entry.  Replaces the CAS-loop atomic per entry of kernels/Coo.cl:4-32.
//
// Any entry order inside a strip / the COO arrays is accepted; order only affects speed.
// Bytes: nnz*(8+V) + (T+1)*4 + Cn*V + R*V (CMRS), nnz*(8+V) + Cn*V + R*V (COO); HBM-bound.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int kSubWarps = 8;      // 4-lane sub-warps per warp
constexpr int kCooPerWarp = 512;  // COO entries per warp (64 per sub-warp, 4 groups per lane)

template <typename T, int HMAX>
__device__ __forceinline__ void cmrs_flush(T (&acc)[HMAX], int key, T sum)
{
#pragma unroll
    for (int h = 0; h < HMAX; ++h) acc[h] += (key == h) ? sum : T(0);
}

// HMAX = accumulator registers per lane (8 or 32); height <= HMAX at run time
// cap > 0 (a plan found strips longer than `cap` entries -- power-law inputs, where a strip holding
// a hub row has 10^5 entries): the main pass (EXTRA = false) walks only the first `cap` entries of
// each strip and stores; the extra pass (EXTRA = true) has one warp per (strip, segment) work item
// of the plan and accumulates with atomics.
// PACKED: `idx` holds (row_in_strip << kPackShift) | column in one word (b200_cmrs_pack) and
// `row_in_strip` is not read: 4 + V bytes per entry instead of 8 + V.
constexpr int kPackShift = 27;  // 5 key bits (height <= 32), columns < 2^27
constexpr int kPackMask = (1 << kPackShift) - 1;

// WPS = warps that share one strip (main pass only).  A cant-sized matrix has one wave of strips
// (7 807 strips = 53 warps per SM), each lane walking ~5 dependent round trips; WPS warps split the
// strip's entries into 8*WPS spans and their `height` row sums meet in shared memory in fixed order
// (deterministic, no atomics) -- the CMRS counterpart of the SELL kernel's WPC, and the "strips staged
// through shared memory" of the north star reduced to what has to be staged: the partial row sums.
template <typename T, int HMAX, bool VEC, int U, bool EXTRA, bool PACKED = false, bool OVL = false, int WPS = 1>
__global__ void __launch_bounds__(kBlock)
cmrs_kernel(const T *__restrict__ data, const int *__restrict__ idx, const int *__restrict__ strip_ptr,
            const int *__restrict__ row_in_strip, const T *__restrict__ x, T *__restrict__ y,
            int n_work, int height, int n_rows, int cap, const int2 *__restrict__ items)
{
    static_assert(!EXTRA || WPS == 1, "extra segments are one warp each");
    static_assert(kWarps % WPS == 0, "the warps of a strip sit in one block");
    __shared__ T red[WPS > 1 ? kWarps : 1][HMAX];
    pdl_launch_dependents();
    bool waited = false;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long work = ((long long)blockIdx.x * kWarps + warp) / WPS;
    const int part = warp % WPS;
    const bool active = work < n_work;
    if (WPS == 1 && !active) return;  // whole warps leave together
    long long strip = active ? work : 0;
    int seg = 0;
    if (EXTRA) {
        const int2 it = items[work];
        strip = it.x;
        seg = it.y;
    }
    T acc[HMAX];
#pragma unroll
    for (int h = 0; h < HMAX; ++h) acc[h] = 0;

    int s = 0, e = 0;
    if (active) {
        s = __ldg(strip_ptr + strip);
        e = __ldg(strip_ptr + strip + 1);
    }
    if (cap > 0) {
        s = min(s + seg * cap, e);
        e = min(s + cap, e);
    }
    constexpr int kSpans = kSubWarps * WPS;
    const int span = ((e - s + kSpans - 1) / kSpans + 3) & ~3;  // multiple of 4 entries
    const int ss = min(s + (part * kSubWarps + (lane >> 2)) * span, e), ee = min(ss + span, e);
    int cur = -1;
    T sum = 0;
    if (VEC) {
        // U groups per lane are loaded (3U independent 128-bit loads), then gathered (4U independent
        // loads), then folded: two memory round trips per 16*U entries of the sub-warp
        const int g_end = ss < ee ? (ee + 3) >> 2 : 0;
        for (int g0 = (ss >> 2) + (lane & 3); g0 < g_end; g0 += 4 * U) {
            IVec4 c[U], r[U];
            Vec4<T> v[U];
            T xv[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = (g0 + 4 * u) << 2;
                if (g0 + 4 * u < g_end) {
                    c[u].load(idx + j);
                    if (!PACKED) r[u].load(row_in_strip + j);
                    v[u].load(data + j);
                    if (PACKED) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            r[u].v[k] = (int)((unsigned)c[u].v[k] >> kPackShift);
                            c[u].v[k] &= kPackMask;
                        }
                    }
                }
            }
            pdl_wait_once<OVL>(waited);  // x may still be being written by the previous launch (common.cuh)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = (g0 + 4 * u) << 2;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    xv[u][k] = (g0 + 4 * u < g_end && j + k >= ss && j + k < ee) ? ld_xo<OVL>(x, c[u].v[k]) : T(0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = (g0 + 4 * u) << 2;
                if (g0 + 4 * u >= g_end) break;
                const bool whole = j >= ss && j + 3 < ee;
                if (whole && r[u].v[0] == cur && r[u].v[1] == cur && r[u].v[2] == cur && r[u].v[3] == cur) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) sum += v[u].v[k] * xv[u][k];
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (j + k >= ss && j + k < ee) {
                            if (r[u].v[k] != cur) {
                                cmrs_flush<T, HMAX>(acc, cur, sum);
                                cur = r[u].v[k];
                                sum = 0;
                            }
                            sum += v[u].v[k] * xv[u][k];
                        }
                    }
                }
            }
        }
    } else {
        pdl_wait_once<OVL>(waited);
        for (int j = ss + (lane & 3); j < ee; j += 4) {
            int c = ld_stream(idx + j);
            const int r = PACKED ? (int)((unsigned)c >> kPackShift) : ld_stream(row_in_strip + j);
            if (PACKED) c &= kPackMask;
            if (r != cur) {
                cmrs_flush<T, HMAX>(acc, cur, sum);
                cur = r;
                sum = 0;
            }
            sum += ld_stream(data + j) * ld_xo<OVL>(x, c);
        }
    }
    cmrs_flush<T, HMAX>(acc, cur, sum);
    pdl_wait_once<OVL>(waited);  // lanes without entries never waited: y must not be written early either

    // transposing butterfly: after the HMAX-halving steps lane L holds the partial of row
    // (L % HMAX) over the lanes of its HMAX-group; plain xor steps finish the sum.
#pragma unroll
    for (int w = HMAX / 2; w >= 1; w >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int h = 0; h < w; ++h) {
            const T send = upper ? acc[h] : acc[h + w];
            const T keep = upper ? acc[h + w] : acc[h];
            acc[h] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    T total = acc[0];
#pragma unroll
    for (int off = HMAX; off < 32; off <<= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
    const long long row = strip * height + lane;
    if (WPS == 1) {
        if (lane < height && row < n_rows) {
            if (EXTRA) atomicAdd(y + row, total);
            else y[row] = total;
        }
    } else {
        if (lane < HMAX) red[warp][lane] = total;
        __syncthreads();
        if (part == 0 && active && lane < height && row < n_rows) {
            T t = red[warp][lane];
#pragma unroll
            for (int w = 1; w < WPS; ++w) t += red[warp + w][lane];
            y[row] = t;
        }
    }
}

__global__ void cmrs_pack_kernel(const int *__restrict__ idx, const int *__restrict__ row_in_strip,
                                 long long nnz, int *__restrict__ packed)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (long long)gridDim.x * blockDim.x)
        packed[i] = (int)(((unsigned)row_in_strip[i] << kPackShift) | (unsigned)idx[i]);
}

constexpr int kCmrsCap = 8192;  // entries per (strip, segment) work item

__global__ void cmrs_long_items_kernel(const int *__restrict__ strip_ptr, int n_strips, int cap,
                                       int *counter, int2 *items)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_strips) return;
    const int len = strip_ptr[t + 1] - strip_ptr[t];
    const int extra = (len + cap - 1) / cap - 1;
    if (extra <= 0) return;
    const int at = atomicAdd(counter, extra);
    if (items)
        for (int k = 0; k < extra; ++k) items[at + k] = make_int2((int)t, k + 1);
}

// y = 0 as a kernel instead of a memset node, so that with launch overlap on the COO kernel can be a
// programmatic dependent of it (and it of whatever ran before): memset nodes take no part in that
template <typename T>
__global__ void zero_ovl_kernel(T *__restrict__ y, int n)
{
    pdl_launch_dependents();
    bool waited = false;
    pdl_wait_once<true>(waited);  // y may still be being read / written by the previous launch
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = T(0);
}

// merges the lanes' final (row key, partial sum) pairs: a warp-level segmented inclusive scan whose
// run heads are found with one ballot; the last lane of every run adds it to y
template <typename T>
__device__ __forceinline__ void merge_final_runs(int cur, T sum, T *__restrict__ y, int lane)
{
    const int prev = __shfl_up_sync(0xffffffffu, cur, 1);
    const bool head = lane == 0 || prev != cur;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T t = __shfl_up_sync(0xffffffffu, sum, off);
        if (lane - off >= start) sum += t;
    }
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    if (tail && cur >= 0) atomicAdd(y + cur, sum);
}

// ---- CMRS on skewed (power-law) inputs: nnz-split -------------------------------------------
// One warp per strip is hopeless when strip lengths span 8 ... 10^5 entries (R-MAT scale 24: mean 134,
// hub strips 40 000): most warps live for one round trip, a few for hundreds.  Here a warp owns a tile
// of kCooPerWarp consecutive ENTRIES, whatever strips they belong to, exactly like the COO kernel; the
// absolute row of an entry is strip * height + row_in_strip, and the strip is found from strip_ptr:
// the plan stores the strip of every tile's first entry (one binary search per tile, done once), a lane
// bisects strip_ptr between its tile's first and last strip for its first entry and then walks forward.
// Rows are accumulated like COO rows: register runs, one atomic per (lane, row run), final runs merged
// by the segmented warp scan.  y is zero-filled first.
__global__ void cmrs_tile_strips_kernel(const int *__restrict__ strip_ptr, int n_strips, int n_tiles,
                                        int *__restrict__ tile_strip)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) {
        tile_strip[t] = n_strips - 1;
        return;
    }
    const int e0 = strip_ptr[0] + t * kCooPerWarp;
    int lo = 0, hi = n_strips - 1;  // largest s with strip_ptr[s] <= e0
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (strip_ptr[mid] <= e0) lo = mid;
        else hi = mid - 1;
    }
    tile_strip[t] = lo;
}

template <typename T, int U, bool PACKED>
__global__ void __launch_bounds__(kBlock)
cmrs_stream_kernel(const T *__restrict__ data, const int *__restrict__ idx, const int *__restrict__ strip_ptr,
                   const int *__restrict__ row_in_strip, const T *__restrict__ x, T *__restrict__ y,
                   int first, int last, int height, const int *__restrict__ tile_strip)
{
    const int lane = threadIdx.x & 31;
    const long long tile = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const long long e0 = (long long)first + tile * kCooPerWarp;
    if (e0 >= last) return;  // whole warps leave together
    const long long ss = e0 + (lane >> 2) * (kCooPerWarp / kSubWarps);
    const long long ee = min(ss + kCooPerWarp / kSubWarps, (long long)last);
    const long long j_first = ss + ((lane & 3) << 2);
    // strip of this lane's first entry: largest s in [s_lo, s_hi] with strip_ptr[s] <= j_first
    int strip = __ldg(tile_strip + tile), s_hi = __ldg(tile_strip + tile + 1);
    while (strip < s_hi) {
        const int mid = (strip + s_hi + 1) >> 1;
        if (__ldg(strip_ptr + mid) <= j_first) strip = mid;
        else s_hi = mid - 1;
    }
    long long next = __ldg(strip_ptr + strip + 1);  // first entry of the following strip
    int base = strip * height;
    int cur = -1;
    T sum = 0;
    constexpr int kGroups = kCooPerWarp / kSubWarps / 16;
#pragma unroll
    for (int batch = 0; batch < kGroups / U; ++batch) {
        const long long j0 = j_first + 16 * U * batch;
        IVec4 r[U], c[U];
        Vec4<T> v[U];
        T xv[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = j0 + 16 * u;
            c[u].zero();
            r[u].zero();
            v[u].zero();
            if (j < ee) {
                c[u].load(idx + j);
                if (!PACKED) r[u].load(row_in_strip + j);
                v[u].load(data + j);
                if (PACKED) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        r[u].v[k] = (int)((unsigned)c[u].v[k] >> kPackShift);
                        c[u].v[k] &= kPackMask;
                    }
                }
            }
        }
        int hold = 0;
        if (U > 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) hold |= hold_bits(c[u], v[u]) & r[u].v[0];
            hold >>= 31;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) xv[u][k] = (j0 + 16 * u + k < ee) ? ld_x(x, c[u].v[k] + hold) : T(0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = j0 + 16 * u;
            if (j >= ee) break;
            const int key0 = base + r[u].v[0];
            if (j + 3 < ee && j + 3 < next && key0 == cur && r[u].v[1] == r[u].v[0] && r[u].v[2] == r[u].v[0] &&
                r[u].v[3] == r[u].v[0]) {
#pragma unroll
                for (int k = 0; k < 4; ++k) sum += v[u].v[k] * xv[u][k];
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (j + k < ee) {
                        while (j + k >= next) {  // entered the next (non-empty) strip
                            ++strip;
                            next = __ldg(strip_ptr + strip + 1);
                            base = strip * height;
                        }
                        const int key = base + r[u].v[k];
                        if (key != cur) {
                            if (cur >= 0) atomicAdd(y + cur, sum);
                            cur = key;
                            sum = 0;
                        }
                        sum += v[u].v[k] * xv[u][k];
                    }
                }
            }
        }
    }
    merge_final_runs<T>(cur, sum, y, lane);
}

template <typename T, bool VEC, int U, bool OVL = false>
__global__ void __launch_bounds__(kBlock)
coo_kernel(const int *__restrict__ row, const int *__restrict__ col, const T *__restrict__ data,
           const T *__restrict__ x, T *__restrict__ y, int nnz)
{
    pdl_launch_dependents();
    bool waited = false;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const long long ss = warp * kCooPerWarp + (lane >> 2) * (kCooPerWarp / kSubWarps);
    const long long ee = min(ss + kCooPerWarp / kSubWarps, (long long)nnz);
    int cur = -1;
    T sum = 0;
    if (VEC) {
        // U of the lane's four groups are loaded together (3U independent 128-bit loads), then
        // gathered (4U independent loads), then folded
        constexpr int kGroups = kCooPerWarp / kSubWarps / 16;
#pragma unroll
        for (int batch = 0; batch < kGroups / U; ++batch) {
            const long long j0 = ss + ((lane & 3) << 2) + 16 * U * batch;
            IVec4 r[U], c[U];
            Vec4<T> v[U];
            T xv[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long j = j0 + 16 * u;
                c[u].zero();
                r[u].zero();
                v[u].zero();
                if (j < ee) {
                    r[u].load(row + j);
                    c[u].load(col + j);
                    v[u].load(data + j);
                }
            }
            // 0 at run time; makes the gathers wait for ALL loads of the batch, row keys included
            // (common.cuh: batch_hold) -- otherwise ptxas sinks most of them below the first gathers
            int hold = 0;
            if (U > 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) hold |= hold_bits(c[u], v[u]) & r[u].v[0];
                hold >>= 31;
            }
            pdl_wait_once<OVL>(waited);  // y is being zeroed, x maybe still written (common.cuh)
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) xv[u][k] = (j0 + 16 * u + k < ee) ? ld_xo<OVL>(x, c[u].v[k] + hold) : T(0);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long j = j0 + 16 * u;
                if (j >= ee) break;
                if (j + 3 < ee && r[u].v[0] == cur && r[u].v[1] == cur && r[u].v[2] == cur && r[u].v[3] == cur) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) sum += v[u].v[k] * xv[u][k];
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (j + k < ee) {
                            if (r[u].v[k] != cur) {
                                if (cur >= 0) atomicAdd(y + cur, sum);
                                cur = r[u].v[k];
                                sum = 0;
                            }
                            sum += v[u].v[k] * xv[u][k];
                        }
                    }
                }
            }
        }
    } else {
        pdl_wait_once<OVL>(waited);
        for (long long j = ss + (lane & 3); j < ee; j += 4) {
            const int r = ld_stream(row + j);
            if (r != cur) {
                if (cur >= 0) atomicAdd(y + cur, sum);
                cur = r;
                sum = 0;
            }
            sum += ld_stream(data + j) * ld_xo<OVL>(x, ld_stream(col + j));
        }
    }
    pdl_wait_once<OVL>(waited);
    merge_final_runs<T>(cur, sum, y, lane);
}

}  // namespace

struct b200_cmrs_plan {
    int device;
    int n_strips;
    int n_items;  // extra (strip, segment) work items
    int cap;
    int2 *items;  // device
    // nnz-split variant for skewed strip lengths (cmrs_stream_kernel): n_tiles > 0
    int first, last;  // strip_ptr[0], strip_ptr[n_strips]
    int max_len;      // longest strip
    int n_tiles;
    int *tile_strip;  // n_tiles + 1 entries
};

namespace {

template <typename T, bool PACKED>
int spmv_cmrs_impl(b200_ctx *ctx, const T *data, const int *idx, const int *strip_ptr,
                   const int *row_in_strip, const T *x, T *y, int n_strips, int height, int n_rows,
                   const b200_cmrs_plan *plan)
{
    B200_TRACE("b200 spmv cmrs");
    B200_ENTER_SPMV(ctx);
    B200_REQUIRE(strip_ptr && x && y && n_strips >= 0 && n_rows >= 0, "bad argument");
    if (height < 1 || height > 32) {
        b200_set_error("CMRS height must be in 1..32, got %d", height);
        return B200_ERR_UNSUPPORTED;
    }
    B200_REQUIRE(!plan || plan->n_strips == n_strips, "plan was built for a different matrix");
    if (n_strips == 0) return B200_SUCCESS;
    const bool vec = aligned16(data) && aligned16(idx) && (PACKED || aligned16(row_in_strip));
    const int n_items = plan ? plan->n_items : 0;
    const int cap = n_items > 0 ? plan->cap : 0;
    const int2 *items = n_items > 0 ? plan->items : nullptr;
    // U = groups loaded per lane and round trip; tuning hook B200_CMRS_U=1|2.  Measured on B200
    // (profiles/r1e_variant_sweep.md), sustained 200-step runs: fp32 2 (0.259 vs 0.284 ms), fp64 1
    // (0.371 vs 0.444: 90 registers); launches of at most ~2 waves (cant): 1
    const bool small = (long long)n_strips * 32 <= 2ll * ctx->sm_count * 2048;
    int u = (sizeof(T) == 4 && !small) ? 2 : 1;
    if (opt_set(ctx, OPT_CMRS_U)) u = ctx->opt[OPT_CMRS_U] == 2 ? 2 : 1;
    // skewed strip lengths (power-law inputs): the nnz-split kernel, if the plan prepared its tiles
    // (tuning hook B200_CMRS_STREAM=0|1 is read when the plan is made)
    const bool ovl = ovl_on(ctx, (long long)n_strips * 32);
    if (plan && plan->n_tiles > 0 && vec) {
        B200_CUDA(cudaMemsetAsync(y, 0, sizeof(T) * (size_t)n_rows, ctx->stream));
        const unsigned blocks = ceil_div_u(plan->n_tiles, kWarps);
        if (u == 2)
            cmrs_stream_kernel<T, 2, PACKED><<<blocks, kBlock, 0, ctx->stream>>>(
                data, idx, strip_ptr, row_in_strip, x, y, plan->first, plan->last, height, plan->tile_strip);
        else
            cmrs_stream_kernel<T, 1, PACKED><<<blocks, kBlock, 0, ctx->stream>>>(
                data, idx, strip_ptr, row_in_strip, x, y, plan->first, plan->last, height, plan->tile_strip);
        B200_LAUNCH_CHECK();
        return B200_SUCCESS;
    }
    // warps per strip (tuning hook B200_CMRS_WPS=1|2|4): matrices of about one wave of strips (cant)
    // split every strip over 2 or 4 warps; strips the plan splits by entries anyway stay one warp each
    int wps = 1;
    while (n_items == 0 && vec && height <= 8 && wps < 4 && (long long)n_strips * wps < (long long)ctx->sm_count * 64) wps <<= 1;
    {
        const int v = opt_or(ctx, OPT_CMRS_WPS, 0);
        if (n_items == 0 && vec && height <= 8 && (v == 1 || v == 2 || v == 4)) wps = v;
    }
#define B200_CMRS_LAUNCH3(H, V, UU, W)                                                                      \
    do {                                                                                                    \
        B200_CUDA(ovl                                                                              \
                      ? b200_launch(ctx, ovl, cmrs_kernel<T, H, V, UU, false, PACKED, true, W>,                  \
                                    dim3(ceil_div_u((long long)n_strips * W, kWarps)), dim3(kBlock), 0, data, idx, strip_ptr, \
                                    row_in_strip, x, y, n_strips, height, n_rows, cap, nullptr)             \
                      : b200_launch(ctx, ovl, cmrs_kernel<T, H, V, UU, false, PACKED, false, W>,                 \
                                    dim3(ceil_div_u((long long)n_strips * W, kWarps)), dim3(kBlock), 0, data, idx, strip_ptr, \
                                    row_in_strip, x, y, n_strips, height, n_rows, cap, nullptr));           \
    } while (0)
#define B200_CMRS_LAUNCH2(H, V, UU)                                                                         \
    do {                                                                                                    \
        B200_CMRS_LAUNCH3(H, V, UU, 1);                                                                     \
        if (n_items > 0)                                                                                    \
            cmrs_kernel<T, H, V, UU, true, PACKED><<<ceil_div_u(n_items, kWarps), kBlock, 0, ctx->stream>>>( \
                data, idx, strip_ptr, row_in_strip, x, y, n_items, height, n_rows, cap, items);             \
    } while (0)
#define B200_CMRS_LAUNCH(H, V)             \
    do {                                   \
        if (u == 2) B200_CMRS_LAUNCH2(H, V, 2); \
        else B200_CMRS_LAUNCH2(H, V, 1);   \
    } while (0)
    if (wps > 1) {  // height <= 8, aligned arrays, no extra items
        if (wps == 4) {
            if (u == 2) B200_CMRS_LAUNCH3(8, true, 2, 4);
            else B200_CMRS_LAUNCH3(8, true, 1, 4);
        } else {
            if (u == 2) B200_CMRS_LAUNCH3(8, true, 2, 2);
            else B200_CMRS_LAUNCH3(8, true, 1, 2);
        }
    } else if (height <= 8) {
        if (vec) B200_CMRS_LAUNCH(8, true);
        else B200_CMRS_LAUNCH(8, false);
    } else {
        if (vec) B200_CMRS_LAUNCH(32, true);
        else B200_CMRS_LAUNCH(32, false);
    }
#undef B200_CMRS_LAUNCH
#undef B200_CMRS_LAUNCH2
#undef B200_CMRS_LAUNCH3
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int spmv_coo_impl(b200_ctx *ctx, const int *row, const int *col, const T *data, const T *x, T *y,
                  int nnz, int n_rows)
{
    B200_TRACE("b200 spmv coo");
    B200_ENTER_SPMV(ctx);
    B200_REQUIRE(x && y && nnz >= 0 && n_rows >= 0, "bad argument");
    B200_REQUIRE(nnz == 0 || (row && col && data), "null row/col/data");
    const bool ovl = ovl_on(ctx, (long long)nnz / 16);
    if (ovl && n_rows > 0) {
        const unsigned zb = (unsigned)min((long long)ctx->sm_count * 8, ((long long)n_rows + 255) / 256);
        B200_CUDA(b200_launch(ctx, ovl, zero_ovl_kernel<T>, dim3(zb), dim3(256), 0, y, n_rows));
    } else {
        B200_CUDA(cudaMemsetAsync(y, 0, sizeof(T) * (size_t)n_rows, ctx->stream));
    }
    if (nnz == 0) return B200_SUCCESS;
    const bool vec = aligned16(row) && aligned16(col) && aligned16(data);
    unsigned blocks = ceil_div_u(((long long)nnz + kCooPerWarp - 1) / kCooPerWarp, kWarps);
    // U = groups loaded per lane and round trip (tuning hook B200_COO_U=1|2|4).  Measured on B200
    // (profiles/r1e_variant_sweep.md): 2, except fp32 on launches of many waves (237 vs 246 us)
    int u = (sizeof(T) == 4 && (long long)blocks > 8ll * ctx->sm_count) ? 4 : 2;
    if (opt_set(ctx, OPT_COO_U)) u = ctx->opt[OPT_COO_U] == 4 ? 4 : (ctx->opt[OPT_COO_U] == 1 ? 1 : 2);
#define B200_COO_LAUNCH(V, UU)                                                                            \
    B200_CUDA(ovl ? b200_launch(ctx, ovl, coo_kernel<T, V, UU, true>, dim3(blocks), dim3(kBlock), 0, row, col, data, x, y, nnz) \
                           : b200_launch(ctx, ovl, coo_kernel<T, V, UU, false>, dim3(blocks), dim3(kBlock), 0, row, col, data, x, y, nnz))
    if (!vec) B200_COO_LAUNCH(false, 1);
    else if (u == 4) B200_COO_LAUNCH(true, 4);
    else if (u == 2) B200_COO_LAUNCH(true, 2);
    else B200_COO_LAUNCH(true, 1);
#undef B200_COO_LAUNCH
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // namespace

extern "C" {

__global__ void cmrs_max_len_kernel(const int *__restrict__ strip_ptr, int n_strips, int *out)
{
    int hi = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_strips; t += (long long)gridDim.x * blockDim.x)
        hi = max(hi, strip_ptr[t + 1] - strip_ptr[t]);
    for (int off = 16; off > 0; off >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out, hi);
}

static int cmrs_plan_fill(b200_ctx *ctx, const int *strip_ptr, int n_strips, b200_cmrs_plan *p)
{
    if (n_strips == 0) return B200_SUCCESS;
    int *counter = ctx->scratch + 224, *maxlen = ctx->scratch + 225;
    const unsigned blocks = ceil_div_u(n_strips, 256);
    int count = 0, ends[2] = {0, 0};
    B200_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), ctx->stream));
    cmrs_long_items_kernel<<<blocks, 256, 0, ctx->stream>>>(strip_ptr, n_strips, p->cap, counter, nullptr);
    cmrs_max_len_kernel<<<(unsigned)min((long long)blocks, (long long)ctx->sm_count * 8), 256, 0, ctx->stream>>>(strip_ptr, n_strips, maxlen);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(&count, counter, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(&p->max_len, maxlen, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(&ends[0], strip_ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(&ends[1], strip_ptr + n_strips, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    p->first = ends[0];
    p->last = ends[1];
    if (count > 0) {
        B200_CUDA(cudaMalloc(&p->items, sizeof(int2) * (size_t)count));
        B200_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
        cmrs_long_items_kernel<<<blocks, 256, 0, ctx->stream>>>(strip_ptr, n_strips, p->cap, counter, p->items);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    p->n_items = count;
    // skewed strip lengths -> the nnz-split kernel (B200_CMRS_STREAM=0|1 overrides).  "Skewed": the
    // longest strip is more than 32 mean strips long (a FEM / stencil matrix stays far below)
    const long long nnz = (long long)ends[1] - ends[0];
    const double mean = (double)nnz / n_strips;
    const bool possible = nnz > 0 && (ends[0] & 3) == 0 && nnz < 0x7fffffffll - kCooPerWarp;
    bool stream = possible && (double)p->max_len > 32.0 * mean;
    if (opt_set(ctx, OPT_CMRS_STREAM)) stream = possible && ctx->opt[OPT_CMRS_STREAM] != 0;
    if (stream) {
        const int n_tiles = (int)((nnz + kCooPerWarp - 1) / kCooPerWarp);
        B200_CUDA(cudaMalloc(&p->tile_strip, sizeof(int) * ((size_t)n_tiles + 1)));
        cmrs_tile_strips_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, ctx->stream>>>(strip_ptr, n_strips, n_tiles, p->tile_strip);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        p->n_tiles = n_tiles;
    }
    return B200_SUCCESS;
}

int b200_cmrs_plan_create(b200_ctx *ctx, const int *strip_ptr, int n_strips, b200_cmrs_plan **plan)
{
    B200_ENTER(ctx);
    B200_REQUIRE(strip_ptr && plan && n_strips >= 0, "bad argument");
    *plan = nullptr;
    b200_cmrs_plan *p = new b200_cmrs_plan();
    p->device = ctx->device;
    p->n_strips = n_strips;
    p->n_items = 0;
    p->cap = kCmrsCap;
    p->items = nullptr;
    p->first = p->last = p->max_len = p->n_tiles = 0;
    p->tile_strip = nullptr;
    const int rc = cmrs_plan_fill(ctx, strip_ptr, n_strips, p);
    if (rc != B200_SUCCESS) {
        b200_cmrs_plan_destroy(p);
        return rc;
    }
    *plan = p;
    return B200_SUCCESS;
}

int b200_cmrs_plan_extra_items(const b200_cmrs_plan *plan, int *n_items)
{
    B200_REQUIRE(plan && n_items, "null argument");
    *n_items = plan->n_items;
    return B200_SUCCESS;
}

int b200_cmrs_plan_stream_tiles(const b200_cmrs_plan *plan, int *n_tiles)
{
    B200_REQUIRE(plan && n_tiles, "null argument");
    *n_tiles = plan->n_tiles;
    return B200_SUCCESS;
}

int b200_cmrs_plan_destroy(b200_cmrs_plan *plan)
{
    if (!plan) return B200_SUCCESS;
    cudaSetDevice(plan->device);
    if (plan->items) cudaFree(plan->items);
    if (plan->tile_strip) cudaFree(plan->tile_strip);
    delete plan;
    return B200_SUCCESS;
}

int b200_spmv_cmrs_f64(b200_ctx *ctx, const double *data, const int *indices, const int *strip_ptr,
                       const int *row_in_strip, const double *vect, double *output, int n_strips,
                       int height, int n_rows, const b200_cmrs_plan *plan)
{
    return spmv_cmrs_impl<double, false>(ctx, data, indices, strip_ptr, row_in_strip, vect, output, n_strips, height, n_rows, plan);
}
int b200_spmv_cmrs_f32(b200_ctx *ctx, const float *data, const int *indices, const int *strip_ptr,
                       const int *row_in_strip, const float *vect, float *output, int n_strips,
                       int height, int n_rows, const b200_cmrs_plan *plan)
{
    return spmv_cmrs_impl<float, false>(ctx, data, indices, strip_ptr, row_in_strip, vect, output, n_strips, height, n_rows, plan);
}
int b200_cmrs_pack(b200_ctx *ctx, const int *indices, const int *row_in_strip, long long nnz, int n_cols,
                   int height, int *packed)
{
    B200_ENTER(ctx);
    B200_REQUIRE(nnz >= 0 && (nnz == 0 || (indices && row_in_strip && packed)), "bad argument");
    if (height < 1 || height > 32 || n_cols > (1 << kPackShift)) {
        b200_set_error("packed CMRS needs height <= 32 and at most 2^%d columns (height %d, %d columns)",
                       kPackShift, height, n_cols);
        return B200_ERR_UNSUPPORTED;
    }
    if (nnz == 0) return B200_SUCCESS;
    const unsigned blocks = (unsigned)min((long long)ctx->sm_count * 16, (nnz + 255) / 256);
    cmrs_pack_kernel<<<blocks, 256, 0, ctx->stream>>>(indices, row_in_strip, nnz, packed);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}
int b200_spmv_cmrs_packed_f64(b200_ctx *ctx, const double *data, const int *packed, const int *strip_ptr,
                              const double *vect, double *output, int n_strips, int height, int n_rows,
                              const b200_cmrs_plan *plan)
{
    return spmv_cmrs_impl<double, true>(ctx, data, packed, strip_ptr, nullptr, vect, output, n_strips, height, n_rows, plan);
}
int b200_spmv_cmrs_packed_f32(b200_ctx *ctx, const float *data, const int *packed, const int *strip_ptr,
                              const float *vect, float *output, int n_strips, int height, int n_rows,
                              const b200_cmrs_plan *plan)
{
    return spmv_cmrs_impl<float, true>(ctx, data, packed, strip_ptr, nullptr, vect, output, n_strips, height, n_rows, plan);
}
int b200_spmv_coo_f64(b200_ctx *ctx, const int *row, const int *col, const double *data,
                      const double *vect, double *output, int nnz, int n_rows)
{
    return spmv_coo_impl<double>(ctx, row, col, data, vect, output, nnz, n_rows);
}
int b200_spmv_coo_f32(b200_ctx *ctx, const int *row, const int *col, const float *data,
                      const float *vect, float *output, int nnz, int n_rows)
{
    return spmv_coo_impl<float>(ctx, row, col, data, vect, output, nnz, n_rows);
}

}  // extern "C"
