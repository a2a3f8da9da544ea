// build_formats.cu -- the five format builds on the GPU.
//
// In the reference every build is a host loop over fscanf'ed triples inlined in main():
// csr.c:72-91, ell.c:68-164, sigma_c.c:71-202, cmrs.c:72-117 (COO is the triples themselves,
// coo.c:75-84).  Here the drivers parse the file once, upload the row-sorted triples and build on
// the device.  Integer arrays are identical to the reference's on its well-defined domain (rows
// sorted, none empty, first row 0); tests/test_gpu_parity.py (against the oracle) and
// tests/test_golden.py (against the recorded reference uploads) check that bit for bit.  The sigma-window sort + permutation of
// SELL-C-sigma is new (the reference has none).
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

inline unsigned grid_for(long long n, int per_block = kBlock)
{
    long long b = (n + per_block - 1) / per_block;
    return (unsigned)(b < 1 ? 1 : b);
}

// ---- sortedness ---------------------------------------------------------------------------
__global__ void check_sorted_kernel(const int *__restrict__ rows, int nnz, int n_rows, int *bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int r = rows[i];
    if (r < 0 || r >= n_rows || (i > 0 && rows[i - 1] > r)) *bad = 1;
}

// ---- CSR ptr: ptr[r] = first entry whose row is >= r ----------------------------------------
__global__ void csr_ptr_kernel(const int *__restrict__ rows, int nnz, int n_rows, int *__restrict__ ptr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    const int prev = i > 0 ? rows[i - 1] : -1;
    const int cur = i < nnz ? rows[i] : n_rows;
    for (int r = prev + 1; r <= cur; ++r) ptr[r] = (int)i;
}

// ---- row statistics -------------------------------------------------------------------------
struct DevStats {
    int max_len, min_len, max_x, min_x;
    unsigned long long sum;
};

__global__ void row_stats_kernel(const int *__restrict__ ptr, int n_rows, DevStats *out)
{
    int hi = 0, lo = 0x7fffffff, hix = 0, lox = 0x7fffffff;
    unsigned long long sum = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (long long)gridDim.x * blockDim.x) {
        const int len = ptr[r + 1] - ptr[r];
        hi = max(hi, len);
        lo = min(lo, len);
        sum += (unsigned long long)len;
        if (r != n_rows - 1) {
            hix = max(hix, len);
            lox = min(lox, len);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, off));
        lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, off));
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&out->max_len, hi);
        atomicMin(&out->min_len, lo);
        atomicMax(&out->max_x, hix);
        atomicMin(&out->min_x, lox);
        atomicAdd(&out->sum, sum);
    }
}

// ---- ELL ------------------------------------------------------------------------------------
// row-major: one warp per row, contiguous reads and writes
template <typename T>
__global__ void ell_fill_rowmajor_kernel(const int *__restrict__ ptr, const int *__restrict__ cols,
                                         const double *__restrict__ vals, int n_rows, int row_size,
                                         int *__restrict__ ell_cols, T *__restrict__ ell_data)
{
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const int s = ptr[row], len = ptr[row + 1] - s;
    const long long base = row * row_size;
    for (int k = lane; k < row_size; k += 32) {
        const bool real = k < len;
        ell_cols[base + k] = real ? cols[s + k] : 0;
        ell_data[base + k] = real ? (T)vals[s + k] : T(0);
    }
}

// column-major: thread per row (writes coalesced), rows n_rows..pitch-1 are zero padding
template <typename T>
__global__ void ell_fill_colmajor_kernel(const int *__restrict__ ptr, const int *__restrict__ cols,
                                         const double *__restrict__ vals, int n_rows, int row_size,
                                         int pitch, int *__restrict__ cm_cols, T *__restrict__ cm_data)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= pitch) return;
    int s = 0, len = 0;
    if (row < n_rows) {
        s = ptr[row];
        len = ptr[row + 1] - s;
    }
    for (int k = 0; k < row_size; ++k) {
        const bool real = k < len;
        cm_cols[(long long)k * pitch + row] = real ? cols[s + k] : 0;
        cm_data[(long long)k * pitch + row] = real ? (T)vals[s + k] : T(0);
    }
}

// ---- SELL-C-sigma ---------------------------------------------------------------------------
// sort keys: (len << 32) | row.  Order: window ascending, then length DESCENDING, then row
// ascending (stable).  Padding keys carry row 0x7fffffff and sort last.
__device__ __forceinline__ bool sigma_before(unsigned long long a, unsigned long long b, int sigma)
{
    const int ra = (int)(a & 0xffffffffu), rb = (int)(b & 0xffffffffu);
    const int wa = ra / sigma, wb = rb / sigma;
    if (wa != wb) return wa < wb;
    const unsigned la = (unsigned)(a >> 32), lb = (unsigned)(b >> 32);
    if (la != lb) return la > lb;
    return ra < rb;
}

__global__ void sigma_keys_kernel(const int *__restrict__ ptr, int n_rows, long long n_pad,
                                  unsigned long long *__restrict__ keys)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i < n_rows)
        keys[i] = ((unsigned long long)(unsigned)(ptr[i + 1] - ptr[i]) << 32) | (unsigned)i;
    else
        keys[i] = 0x7fffffffull;
}

__global__ void bitonic_step_kernel(unsigned long long *__restrict__ keys, long long n_pad,
                                    long long k, long long j, long long limit, int sigma)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const long long l = i ^ j;
    if (l <= i) return;
    const bool ascending = (i & k) == 0 || k == limit;
    const unsigned long long a = keys[i], b = keys[l];
    if (sigma_before(b, a, sigma) == ascending) {
        keys[i] = b;
        keys[l] = a;
    }
}

// whole bitonic network for blocks of `tile` (<= 2048) keys in shared memory, up to level `limit`
__global__ void __launch_bounds__(1024)
bitonic_smem_kernel(unsigned long long *__restrict__ keys, int tile, long long limit, int sigma)
{
    extern __shared__ unsigned long long sk[];
    const long long base = (long long)blockIdx.x * tile;
    for (int t = threadIdx.x; t < tile; t += blockDim.x) sk[t] = keys[base + t];
    __syncthreads();
    for (long long k = 2; k <= limit && k <= tile; k <<= 1) {
        for (int j = (int)(k >> 1); j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < tile; t += blockDim.x) {
                const int l = t ^ j;
                if (l > t) {
                    const bool ascending = ((base + t) & k) == 0 || k == limit;
                    const unsigned long long a = sk[t], b = sk[l];
                    if (sigma_before(b, a, sigma) == ascending) {
                        sk[t] = b;
                        sk[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < tile; t += blockDim.x) keys[base + t] = sk[t];
}

// the j < tile tail of one global level k, fused in shared memory
__global__ void __launch_bounds__(1024)
bitonic_smem_tail_kernel(unsigned long long *__restrict__ keys, int tile, long long k, long long limit,
                         int sigma)
{
    extern __shared__ unsigned long long sk[];
    const long long base = (long long)blockIdx.x * tile;
    for (int t = threadIdx.x; t < tile; t += blockDim.x) sk[t] = keys[base + t];
    __syncthreads();
    for (int j = tile >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < tile; t += blockDim.x) {
            const int l = t ^ j;
            if (l > t) {
                const bool ascending = ((base + t) & k) == 0 || k == limit;
                const unsigned long long a = sk[t], b = sk[l];
                if (sigma_before(b, a, sigma) == ascending) {
                    sk[t] = b;
                    sk[l] = a;
                }
            }
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < tile; t += blockDim.x) keys[base + t] = sk[t];
}

__global__ void keys_to_perm_kernel(const unsigned long long *__restrict__ keys, int n_rows,
                                    int *__restrict__ perm)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) perm[i] = (int)(keys[i] & 0xffffffffu);
}

__global__ void iota_kernel(int *__restrict__ p, int n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

// slice widths: one warp per slice, width = 32 * max(1, longest row of the slice)
__global__ void sell_widths_kernel(const int *__restrict__ ptr, const int *__restrict__ perm,
                                   int n_rows, int n_slices, long long *__restrict__ widths)
{
    const long long slice = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    const long long nr = slice * 32 + lane;
    int len = 0;
    if (nr < n_rows) {
        const int old = perm ? perm[nr] : (int)nr;
        len = ptr[old + 1] - ptr[old];
    }
    for (int off = 16; off > 0; off >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, off));
    if (lane == 0) widths[slice] = 32ll * max(len, 1);
}

// single-block exclusive scan, in place: out[0..n] from in[0..n-1] (out has n+1 entries)
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(long long *__restrict__ a, long long n)
{
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < n ? a[i] : 0;
        long long incl = v;
        for (int off = 1; off < 32; off <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += t;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const long long carry = carry_s;
        const long long before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - v;
        if (i < n) a[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = carry_s;
}

template <typename T>
__global__ void sell_fill_kernel(const int *__restrict__ ptr, const int *__restrict__ cols,
                                 const double *__restrict__ vals, int n_rows, int n_slices,
                                 const int *__restrict__ perm, const long long *__restrict__ slice_ptr,
                                 int *__restrict__ sell_cols, T *__restrict__ sell_data)
{
    const long long slice = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    const long long base = slice_ptr[slice];
    const int width = (int)((slice_ptr[slice + 1] - base) >> 5);
    const long long nr = slice * 32 + lane;
    int s = 0, len = 0;
    if (nr < n_rows) {
        const int old = perm ? perm[nr] : (int)nr;
        s = ptr[old];
        len = ptr[old + 1] - s;
    }
    for (int k = 0; k < width; ++k) {
        const bool real = k < len;
        sell_cols[base + (long long)k * 32 + lane] = real ? cols[s + k] : 0;
        sell_data[base + (long long)k * 32 + lane] = real ? (T)vals[s + k] : T(0);
    }
}

__global__ void narrow_ptr_kernel(const long long *__restrict__ in, int n, int *__restrict__ out, int *bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long v = in[i];
    if (v > 0x7fffffffll) *bad = 1;
    out[i] = (int)v;
}

// ---- CMRS -----------------------------------------------------------------------------------
__global__ void cmrs_strip_ptr_kernel(const int *__restrict__ ptr, int n_rows, int height, int n_strips,
                                      int *__restrict__ strip_ptr)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_strips) return;
    const long long r = t * height;
    strip_ptr[t] = ptr[r < n_rows ? r : n_rows];
}

__global__ void cmrs_row_in_strip_kernel(const int *__restrict__ rows, int nnz, int height,
                                         int *__restrict__ row_in_strip)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) row_in_strip[i] = rows[i] % height;
}

// ---- misc -----------------------------------------------------------------------------------
__global__ void convert_kernel(const double *__restrict__ src, float *__restrict__ dst, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

__global__ void offset_kernel(int *__restrict__ a, long long n, int delta)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        a[i] += delta;
}

template <typename T>
__global__ void ramp_kernel(T *__restrict__ x, int n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (T)i;
}

int read_flag(b200_ctx *ctx, int *flag_dev, int *out)
{
    B200_CUDA(cudaMemcpyAsync(out, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_SUCCESS;
}

template <typename T>
int build_ell_impl(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals, int n_rows,
                   int row_size, int *ell_cols, T *ell_data)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && n_rows >= 0 && row_size >= 0, "bad argument");
    if (n_rows == 0 || row_size == 0) return B200_SUCCESS;
    B200_REQUIRE(ell_cols && ell_data, "null output");
    ell_fill_rowmajor_kernel<T><<<grid_for((long long)n_rows * 32), kBlock, 0, ctx->stream>>>(
        ptr, cols, vals, n_rows, row_size, ell_cols, ell_data);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int build_ellcm_impl(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals, int n_rows,
                     int row_size, int pitch, int *cm_cols, T *cm_data)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && n_rows >= 0 && row_size >= 0, "bad argument");
    B200_REQUIRE(pitch >= n_rows && pitch % 32 == 0, "pitch must be a multiple of 32 and >= n_rows");
    if (pitch == 0 || row_size == 0) return B200_SUCCESS;
    B200_REQUIRE(cm_cols && cm_data, "null output");
    ell_fill_colmajor_kernel<T><<<grid_for(pitch), kBlock, 0, ctx->stream>>>(
        ptr, cols, vals, n_rows, row_size, pitch, cm_cols, cm_data);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int sell_fill_impl(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals, int n_rows,
                   int chunk, const int *perm, const long long *slice_ptr, int *sell_cols, T *sell_data)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && slice_ptr && n_rows >= 0, "bad argument");
    if (chunk != 32) {
        b200_set_error("SELL chunk must be 32, got %d", chunk);
        return B200_ERR_UNSUPPORTED;
    }
    const int n_slices = b200_sell_num_slices(n_rows, chunk);
    if (n_slices == 0) return B200_SUCCESS;
    B200_REQUIRE(sell_cols && sell_data, "null output");
    sell_fill_kernel<T><<<grid_for((long long)n_slices * 32), kBlock, 0, ctx->stream>>>(
        ptr, cols, vals, n_rows, n_slices, perm, slice_ptr, sell_cols, sell_data);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // namespace

extern "C" {

int b200_check_sorted_rows(b200_ctx *ctx, const int *rows, int nnz, int n_rows)
{
    B200_ENTER(ctx);
    B200_REQUIRE(nnz >= 0 && n_rows >= 0 && (nnz == 0 || rows), "bad argument");
    if (nnz == 0) return B200_SUCCESS;
    B200_CUDA(cudaMemsetAsync(ctx->scratch, 0, sizeof(int), ctx->stream));
    check_sorted_kernel<<<grid_for(nnz), kBlock, 0, ctx->stream>>>(rows, nnz, n_rows, ctx->scratch);
    B200_LAUNCH_CHECK();
    int bad = 0;
    int rc = read_flag(ctx, ctx->scratch, &bad);
    if (rc) return rc;
    if (bad) {
        b200_set_error("rows are not sorted ascending within [0, n_rows)");
        return B200_ERR_DOMAIN;
    }
    return B200_SUCCESS;
}

int b200_build_csr_ptr(b200_ctx *ctx, const int *rows, int nnz, int n_rows, int *ptr)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && nnz >= 0 && n_rows >= 0 && (nnz == 0 || rows), "bad argument");
    csr_ptr_kernel<<<grid_for((long long)nnz + 1), kBlock, 0, ctx->stream>>>(rows, nnz, n_rows, ptr);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_row_length_stats(b200_ctx *ctx, const int *ptr, int n_rows, b200_row_stats *stats)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && stats && n_rows >= 0, "bad argument");
    DevStats init = {0, 0x7fffffff, 0, 0x7fffffff, 0ull};
    DevStats *d = reinterpret_cast<DevStats *>(ctx->scratch + 64);
    B200_CUDA(cudaMemcpyAsync(d, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    if (n_rows > 0) {
        int blocks = (int)min((long long)ctx->sm_count * 8, ((long long)n_rows + 255) / 256);
        row_stats_kernel<<<blocks, 256, 0, ctx->stream>>>(ptr, n_rows, d);
        B200_LAUNCH_CHECK();
    }
    DevStats got;
    int last[2] = {0, 0};
    B200_CUDA(cudaMemcpyAsync(&got, d, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_rows > 0)
        B200_CUDA(cudaMemcpyAsync(last, ptr + n_rows - 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    stats->max_len = got.max_len;
    stats->min_len = n_rows > 0 ? got.min_len : 0;
    stats->sum_len = (long long)got.sum;
    stats->last_len = last[1] - last[0];
    stats->max_len_excl_last = got.max_x;
    stats->min_len_excl_last = got.min_x;  // INT_MAX when n_rows <= 1, as ell.c:69 leaves it
    stats->sum_len_excl_last = (long long)got.sum - stats->last_len;
    return B200_SUCCESS;
}

int b200_build_ell_f64(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                       int n_rows, int row_size, int *ell_cols, double *ell_data)
{
    return build_ell_impl<double>(ctx, ptr, cols, vals, n_rows, row_size, ell_cols, ell_data);
}
int b200_build_ell_f32(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                       int n_rows, int row_size, int *ell_cols, float *ell_data)
{
    return build_ell_impl<float>(ctx, ptr, cols, vals, n_rows, row_size, ell_cols, ell_data);
}
int b200_build_ell_colmajor_f64(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                                int n_rows, int row_size, int pitch, int *cm_cols, double *cm_data)
{
    return build_ellcm_impl<double>(ctx, ptr, cols, vals, n_rows, row_size, pitch, cm_cols, cm_data);
}
int b200_build_ell_colmajor_f32(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                                int n_rows, int row_size, int pitch, int *cm_cols, float *cm_data)
{
    return build_ellcm_impl<float>(ctx, ptr, cols, vals, n_rows, row_size, pitch, cm_cols, cm_data);
}

int b200_sell_num_slices(int n_rows, int chunk)
{
    if (chunk <= 0 || n_rows <= 0) return 0;
    return (int)(((long long)n_rows + chunk - 1) / chunk);  // sigma_c.c:74-81
}

int b200_build_sell_ptr(b200_ctx *ctx, const int *ptr, int n_rows, int chunk, int sigma, int *perm,
                        long long *slice_ptr, long long *total)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && slice_ptr && total && n_rows >= 0, "bad argument");
    if (chunk != 32) {
        b200_set_error("SELL chunk must be 32, got %d", chunk);
        return B200_ERR_UNSUPPORTED;
    }
    *total = 0;
    const int n_slices = b200_sell_num_slices(n_rows, chunk);
    if (n_slices == 0) {
        B200_CUDA(cudaMemsetAsync(slice_ptr, 0, sizeof(long long), ctx->stream));
        return B200_SUCCESS;
    }
    if (sigma > 1) {
        B200_REQUIRE(perm, "sigma > 1 needs a perm output");
        long long n_pad = 1;
        while (n_pad < n_rows) n_pad <<= 1;
        const bool pow2 = (sigma & (sigma - 1)) == 0;
        long long limit = pow2 ? (sigma < n_pad ? sigma : n_pad) : n_pad;
        unsigned long long *keys = nullptr;
        B200_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)n_pad));
        sigma_keys_kernel<<<grid_for(n_pad), kBlock, 0, ctx->stream>>>(ptr, n_rows, n_pad, keys);
        const int tile = (int)(n_pad < 2048 ? n_pad : 2048);
        const size_t smem = sizeof(unsigned long long) * (size_t)tile;
        const unsigned tiles = (unsigned)(n_pad / tile);
        if (tile >= 2)
            bitonic_smem_kernel<<<tiles, tile >= 2048 ? 1024 : (tile / 2 < 32 ? 32 : tile / 2), smem, ctx->stream>>>(keys, tile, limit, sigma);
        for (long long k = (long long)tile * 2; k <= limit; k <<= 1) {
            for (long long j = k >> 1; j >= tile; j >>= 1)
                bitonic_step_kernel<<<grid_for(n_pad), kBlock, 0, ctx->stream>>>(keys, n_pad, k, j, limit, sigma);
            bitonic_smem_tail_kernel<<<tiles, 1024, smem, ctx->stream>>>(keys, tile, k, limit, sigma);
        }
        keys_to_perm_kernel<<<grid_for(n_rows), kBlock, 0, ctx->stream>>>(keys, n_rows, perm);
        cudaError_t e = cudaGetLastError();
        cudaError_t es = cudaStreamSynchronize(ctx->stream);
        cudaFree(keys);
        if (e != cudaSuccess) return b200_cuda_fail(e, "sigma sort", __FILE__, __LINE__);
        if (es != cudaSuccess) return b200_cuda_fail(es, "sigma sort", __FILE__, __LINE__);
    } else if (perm) {
        iota_kernel<<<grid_for(n_rows), kBlock, 0, ctx->stream>>>(perm, n_rows);
        B200_LAUNCH_CHECK();
    }
    sell_widths_kernel<<<grid_for((long long)n_slices * 32), kBlock, 0, ctx->stream>>>(
        ptr, sigma > 1 ? perm : nullptr, n_rows, n_slices, slice_ptr);
    B200_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(slice_ptr, n_slices);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(total, slice_ptr + n_slices, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_SUCCESS;
}

int b200_sell_ptr_to_i32(b200_ctx *ctx, const long long *slice_ptr, int n_slices, int *row_indices)
{
    B200_ENTER(ctx);
    B200_REQUIRE(slice_ptr && row_indices && n_slices >= 0, "bad argument");
    B200_CUDA(cudaMemsetAsync(ctx->scratch, 0, sizeof(int), ctx->stream));
    narrow_ptr_kernel<<<grid_for((long long)n_slices + 1), kBlock, 0, ctx->stream>>>(
        slice_ptr, n_slices + 1, row_indices, ctx->scratch);
    B200_LAUNCH_CHECK();
    int bad = 0;
    int rc = read_flag(ctx, ctx->scratch, &bad);
    if (rc) return rc;
    if (bad) {
        b200_set_error("padded SELL size exceeds 2^31-1: use the 64-bit slice pointers");
        return B200_ERR_DOMAIN;
    }
    return B200_SUCCESS;
}

int b200_build_sell_fill_f64(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                             int n_rows, int chunk, const int *perm, const long long *slice_ptr,
                             int *sell_cols, double *sell_data)
{
    return sell_fill_impl<double>(ctx, ptr, cols, vals, n_rows, chunk, perm, slice_ptr, sell_cols, sell_data);
}
int b200_build_sell_fill_f32(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                             int n_rows, int chunk, const int *perm, const long long *slice_ptr,
                             int *sell_cols, float *sell_data)
{
    return sell_fill_impl<float>(ctx, ptr, cols, vals, n_rows, chunk, perm, slice_ptr, sell_cols, sell_data);
}

int b200_cmrs_num_strips(int n_rows, int height)
{
    if (height <= 0 || n_rows <= 0) return 0;
    return (int)(((long long)n_rows + height - 1) / height);  // cmrs.c:72
}

int b200_build_cmrs(b200_ctx *ctx, const int *rows, const int *ptr, int nnz, int n_rows, int height,
                    int *strip_ptr, int *row_in_strip)
{
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && strip_ptr && nnz >= 0 && n_rows >= 0, "bad argument");
    B200_REQUIRE(height >= 1, "height must be positive");
    B200_REQUIRE(nnz == 0 || (rows && row_in_strip), "null rows/row_in_strip");
    const int n_strips = b200_cmrs_num_strips(n_rows, height);
    cmrs_strip_ptr_kernel<<<grid_for((long long)n_strips + 1), kBlock, 0, ctx->stream>>>(
        ptr, n_rows, height, n_strips, strip_ptr);
    B200_LAUNCH_CHECK();
    if (nnz > 0) {
        cmrs_row_in_strip_kernel<<<grid_for(nnz), kBlock, 0, ctx->stream>>>(rows, nnz, height, row_in_strip);
        B200_LAUNCH_CHECK();
    }
    return B200_SUCCESS;
}

int b200_convert_f64_to_f32(b200_ctx *ctx, const double *src, float *dst, long long n)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "bad argument");
    if (n == 0) return B200_SUCCESS;
    long long blocks = (n + kBlock - 1) / kBlock;
    if (blocks > (long long)ctx->sm_count * 32) blocks = (long long)ctx->sm_count * 32;
    convert_kernel<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(src, dst, n);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_offset_i32(b200_ctx *ctx, int *a, long long n, int delta)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && (n == 0 || a), "bad argument");
    if (n == 0 || delta == 0) return B200_SUCCESS;
    long long blocks = (n + kBlock - 1) / kBlock;
    if (blocks > (long long)ctx->sm_count * 32) blocks = (long long)ctx->sm_count * 32;
    offset_kernel<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(a, n, delta);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_fill_ramp_f64(b200_ctx *ctx, double *x, int n)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && (n == 0 || x), "bad argument");
    if (n == 0) return B200_SUCCESS;
    ramp_kernel<double><<<grid_for(n), kBlock, 0, ctx->stream>>>(x, n);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_fill_ramp_f32(b200_ctx *ctx, float *x, int n)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && (n == 0 || x), "bad argument");
    if (n == 0) return B200_SUCCESS;
    ramp_kernel<float><<<grid_for(n), kBlock, 0, ctx->stream>>>(x, n);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // extern "C"
