// packed_formats.cu -- fewer bytes per entry, format advice, and conversions back to CSR.
//
// Every SpMV kernel of this library runs at the HBM roofline on FEM-like matrices, so the only lever
// left is the byte count.  The reference's formats spend a 32-bit column index per entry (cl_int
// indices, sigma_c.c:40).  In a banded / FEM matrix the columns of one 32-row chunk span a few thousand
// columns at most, so a chunk can store 16-bit offsets from its smallest column:
//   SELL-32 with 16-bit column deltas ("sell16"): delta16[P] + chunk_base[S]: 2 + V bytes per entry
//   instead of 4 + V (fp32: 6 instead of 8).  A DERIVED device layout, like column-major ELL and packed
//   CMRS: the reference arrays stay the bit-exact build product and unpack(delta16, chunk_base) ==
//   indices is tested bit for bit.
// b200_format_advice turns the row statistics the builders compute anyway into per-format byte counts
// and a recommendation; b200_*_to_csr convert ELL / SELL / CMRS arrays back to CSR on the device, so any
// format converts to any other through CSR and the existing builders.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

inline unsigned grid_for(long long n, int per_block = kBlock)
{
    long long b = (n + per_block - 1) / per_block;
    return (unsigned)(b < 1 ? 1 : b);
}

// ---- sell16: pack ---------------------------------------------------------------------------
// one warp per chunk: smallest / largest column over the chunk's real entries (padding slots are
// (column 0, value 0) in the format itself and get delta 0), then the deltas.  Columns are taken
// modulo n_cols: a chunk whose band wraps around the matrix edge (periodic stencils, the bench's banded
// matrix) has its base in the upper half and deltas that run through column 0.
template <typename T>
__global__ void sell_pack16_kernel(const T *__restrict__ data, const int *__restrict__ idx,
                                   const int *__restrict__ slice_ptr, int n_slices, int n_cols,
                                   int *__restrict__ chunk_base, unsigned short *__restrict__ d16, int *__restrict__ too_wide)
{
    const long long slice = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    const long long b = slice_ptr[slice], e = slice_ptr[slice + 1];
    const int half = n_cols / 2;
    int lo = 0x7fffffff, hi = -1;        // over all real entries
    int lo_up = 0x7fffffff, hi_low = -1;  // smallest column >= half, largest column < half
    for (long long j = b + lane; j < e; j += 32) {
        const int c = idx[j];
        if (c != 0 || data[j] != T(0)) {
            lo = min(lo, c);
            hi = max(hi, c);
            if (c >= half) lo_up = min(lo_up, c);
            else hi_low = max(hi_low, c);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        lo_up = min(lo_up, __shfl_xor_sync(0xffffffffu, lo_up, off));
        hi_low = max(hi_low, __shfl_xor_sync(0xffffffffu, hi_low, off));
    }
    int base = lo;
    if (hi < 0) {
        base = 0;  // a chunk of padding only
    } else if (hi - lo > 0xffff) {
        // does the chunk fit when it starts in the upper half and runs through column 0?
        if (hi_low >= 0 && lo_up != 0x7fffffff && (long long)hi_low + n_cols - lo_up <= 0xffff) base = lo_up;
        else if (lane == 0) atomicExch(too_wide, 1);
    }
    if (lane == 0) chunk_base[slice] = base;
    for (long long j = b + lane; j < e; j += 32) {
        const int c = idx[j];
        const bool real = c != 0 || data[j] != T(0);
        const long long d = c >= base ? c - base : (long long)c + n_cols - base;
        d16[j] = (unsigned short)(real ? (d > 0xffff ? 0xffff : d) : 0);
    }
}

// ---- sell16: SpMV -----------------------------------------------------------------------------
// The SELL-32 kernel of spmv_sell_ell.cu (warp = chunk, lane L owns rows 4(L%8)..+3 of column
// 4*it + L/8, U groups per round trip) with the index stream halved: four 16-bit deltas arrive as one
// 64-bit load (256 contiguous bytes per warp instruction); column = (chunk base + delta) mod n_cols.
template <typename T, int U>
__global__ void __launch_bounds__(kBlock)
sell32_d16_kernel(const T *__restrict__ data, const unsigned short *__restrict__ d16, const int *__restrict__ chunk_base,
                  const T *__restrict__ x, T *__restrict__ y, const int *__restrict__ slice_ptr, int n_slices, int n_out,
                  int n_cols)
{
    const int lane = threadIdx.x & 31;
    const long long slice = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (slice >= n_slices) return;  // whole warps leave together
    const long long cb = slice_ptr[slice];
    const long long n_groups = ((long long)slice_ptr[slice + 1] - cb) >> 2;
    const int base = __ldg(chunk_base + slice);
    const int wrap_at = n_cols - base;  // deltas >= wrap_at have wrapped past the last column
    const unsigned short *ip = d16 + cb;
    const T *dp = data + cb;
    T acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    auto col = [&](unsigned d, int hold) {
        const int dd = (int)d + hold;
        return dd >= wrap_at ? dd - wrap_at : dd + base;
    };
    for (long long g0 = lane; g0 < n_groups; g0 += 32 * U) {
        uint2 c[U];
        Vec4<T> v[U];
        T xv[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long g = g0 + 32 * u;
            c[u] = make_uint2(0u, 0u);
            v[u].zero();
            if (g < n_groups) {
                c[u] = __ldcs(reinterpret_cast<const uint2 *>(ip + (g << 2)));
                v[u].load(dp + (g << 2));
            }
        }
        // 0 at run time; orders the gathers after ALL loads of the batch (common.cuh: batch_hold)
        int hold = 0;
        if (U > 1) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                hold |= (int)(c[u].x >> 1) & (sizeof(T) == 8 ? __double2hiint((double)v[u].v[0]) : __float_as_int((float)v[u].v[0]));
            hold >>= 31;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xv[u][0] = ld_x(x, col(c[u].x & 0xffffu, hold));
            xv[u][1] = ld_x(x, col(c[u].x >> 16, hold));
            xv[u][2] = ld_x(x, col(c[u].y & 0xffffu, hold));
            xv[u][3] = ld_x(x, col(c[u].y >> 16, hold));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc0 += v[u].v[0] * xv[u][0];
            acc1 += v[u].v[1] * xv[u][1];
            acc2 += v[u].v[2] * xv[u][2];
            acc3 += v[u].v[3] * xv[u][3];
        }
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
            if (r + k < n_out) y[r + k] = a[k];
    }
}

// ---- conversions back to CSR --------------------------------------------------------------------
// row lengths of a padded layout = its real entries ((column 0, value 0) slots are padding)
template <typename T>
__global__ void ell_row_len_kernel(const T *__restrict__ data, const int *__restrict__ idx, int n_rows, int row_size,
                                   int *__restrict__ len)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int n = 0;
    for (int k = 0; k < row_size; ++k) {
        const long long j = r * row_size + k;
        n += (idx[j] != 0 || data[j] != T(0));
    }
    len[r] = n;
}

template <typename T>
__global__ void sell_row_len_kernel(const T *__restrict__ data, const int *__restrict__ idx,
                                    const int *__restrict__ slice_ptr, int n_rows, int *__restrict__ len)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const long long s = r >> 5, b = slice_ptr[s], w = (slice_ptr[s + 1] - b) >> 5;
    int n = 0;
    for (long long k = 0; k < w; ++k) {
        const long long j = b + (r & 31) + 32 * k;
        n += (idx[j] != 0 || data[j] != T(0));
    }
    len[r] = n;
}

// in-place exclusive scan of len[0..n) into ptr[0..n] by one block (row counts fit int: nnz < 2^31)
__global__ void __launch_bounds__(1024) scan_i32_kernel(int *__restrict__ a, int n)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? a[i] : 0;
        int s = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, off);
            if ((threadIdx.x & 31) >= off) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, off);
                if (threadIdx.x >= off) w += t;
            }
            warp_tot[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry + (threadIdx.x >= 32 ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + s - v;
        __syncthreads();
        if (i < n) a[i] = before;
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = carry;
}

template <typename T>
__global__ void ell_to_csr_fill_kernel(const T *__restrict__ data, const int *__restrict__ idx, int n_rows, int row_size,
                                       const int *__restrict__ ptr, int *__restrict__ cols, T *__restrict__ vals)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int at = ptr[r];
    for (int k = 0; k < row_size; ++k) {
        const long long j = r * row_size + k;
        if (idx[j] != 0 || data[j] != T(0)) {
            cols[at] = idx[j];
            vals[at] = data[j];
            ++at;
        }
    }
}

template <typename T>
__global__ void sell_to_csr_fill_kernel(const T *__restrict__ data, const int *__restrict__ idx,
                                        const int *__restrict__ slice_ptr, int n_rows, const int *__restrict__ ptr,
                                        int *__restrict__ cols, T *__restrict__ vals)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const long long s = r >> 5, b = slice_ptr[s], w = (slice_ptr[s + 1] - b) >> 5;
    int at = ptr[r];
    for (long long k = 0; k < w; ++k) {
        const long long j = b + (r & 31) + 32 * k;
        if (idx[j] != 0 || data[j] != T(0)) {
            cols[at] = idx[j];
            vals[at] = data[j];
            ++at;
        }
    }
}

// CMRS -> CSR: the entries are already in CSR order when rows are sorted inside a strip (the builder's
// output); the row pointer is the histogram of strip * height + row_in_strip
__global__ void cmrs_row_count_kernel(const int *__restrict__ strip_ptr, const int *__restrict__ row_in_strip,
                                      int n_strips, int height, int n_rows, int *__restrict__ len, int *__restrict__ unsorted)
{
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_strips) return;
    const int b = strip_ptr[t], e = strip_ptr[t + 1];
    for (int j = b + lane; j < e; j += 32) {
        const int r = row_in_strip[j];
        const long long row = t * height + r;
        if (row < n_rows) atomicAdd(len + row, 1);
        if (j + 1 < e && row_in_strip[j + 1] < r) atomicExch(unsorted, 1);
    }
}

template <typename T>
int sell_pack16_impl(b200_ctx *ctx, const T *data, const int *indices, const int *row_indices, int n_slices, int n_cols,
                     int *chunk_base, unsigned short *delta16)
{
    B200_TRACE("b200 sell pack16");
    B200_ENTER(ctx);
    B200_REQUIRE(row_indices && n_slices >= 0 && n_cols >= 1 && (n_slices == 0 || (data && indices && chunk_base && delta16)),
                 "bad argument");
    if (n_slices == 0) return B200_SUCCESS;
    int *flag = ctx->scratch + 320;
    B200_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    sell_pack16_kernel<T><<<grid_for((long long)n_slices * 32), kBlock, 0, ctx->stream>>>(data, indices, row_indices, n_slices,
                                                                                         n_cols, chunk_base, delta16, flag);
    B200_LAUNCH_CHECK();
    int wide = 0;
    B200_CUDA(cudaMemcpyAsync(&wide, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (wide) {
        b200_set_error("a 32-row chunk spans more than 65536 columns (even modulo n_cols): 16-bit column deltas cannot hold this matrix");
        return B200_ERR_UNSUPPORTED;
    }
    return B200_SUCCESS;
}

template <typename T>
int spmv_sell16_impl(b200_ctx *ctx, const T *data, const unsigned short *delta16, const int *chunk_base, const T *x, T *y,
                     const int *row_indices, int chunk, int n_slices, int n_out, int n_cols)
{
    B200_TRACE("b200 spmv sell16");
    B200_ENTER(ctx);
    B200_REQUIRE(x && y && row_indices && n_slices >= 0 && n_out >= 0 && n_cols >= 1, "bad argument");
    if (chunk != 32) {
        b200_set_error("SELL chunk must be 32 (warp-aligned), got %d", chunk);
        return B200_ERR_UNSUPPORTED;
    }
    B200_REQUIRE((long long)n_out <= (long long)n_slices * 32, "n_out exceeds n_slices*32");
    if (n_slices == 0) return B200_SUCCESS;
    B200_REQUIRE(data && delta16 && chunk_base, "null array");
    B200_REQUIRE(aligned16(data) && (reinterpret_cast<uintptr_t>(delta16) & 7) == 0, "sell16 arrays must be 16 / 8-byte aligned");
    const unsigned blocks = grid_for((long long)n_slices * 32);
    // same batch depths as the 32-bit kernel (profiles/r1e_variant_sweep.md); hook B200_SELL_UNROLL
    int u = 4;
    {
        const int v = opt_or(ctx, OPT_SELL_UNROLL, 0);
        if (v == 1 || v == 2 || v == 4) u = v;
    }
    if (u == 4) sell32_d16_kernel<T, 4><<<blocks, kBlock, 0, ctx->stream>>>(data, delta16, chunk_base, x, y, row_indices, n_slices, n_out, n_cols);
    else if (u == 2) sell32_d16_kernel<T, 2><<<blocks, kBlock, 0, ctx->stream>>>(data, delta16, chunk_base, x, y, row_indices, n_slices, n_out, n_cols);
    else sell32_d16_kernel<T, 1><<<blocks, kBlock, 0, ctx->stream>>>(data, delta16, chunk_base, x, y, row_indices, n_slices, n_out, n_cols);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int ell_to_csr_impl(b200_ctx *ctx, const T *data, const int *indices, int n_rows, int row_size, int *ptr, int *cols, T *vals,
                    long long *nnz)
{
    B200_TRACE("b200 ell -> csr");
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && n_rows >= 0 && row_size >= 0 && (n_rows == 0 || row_size == 0 || (data && indices)), "bad argument");
    if (n_rows > 0) {
        ell_row_len_kernel<T><<<grid_for(n_rows), kBlock, 0, ctx->stream>>>(data, indices, n_rows, row_size, ptr);
        B200_LAUNCH_CHECK();
    }
    scan_i32_kernel<<<1, 1024, 0, ctx->stream>>>(ptr, n_rows);
    B200_LAUNCH_CHECK();
    if (cols && vals && n_rows > 0) {
        ell_to_csr_fill_kernel<T><<<grid_for(n_rows), kBlock, 0, ctx->stream>>>(data, indices, n_rows, row_size, ptr, cols, vals);
        B200_LAUNCH_CHECK();
    }
    if (nnz) {
        int last = 0;
        B200_CUDA(cudaMemcpyAsync(&last, ptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *nnz = last;
    }
    return B200_SUCCESS;
}

template <typename T>
int sell_to_csr_impl(b200_ctx *ctx, const T *data, const int *indices, const int *row_indices, int n_rows, int *ptr, int *cols,
                     T *vals, long long *nnz)
{
    B200_TRACE("b200 sell -> csr");
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && row_indices && n_rows >= 0 && (n_rows == 0 || (data && indices)), "bad argument");
    if (n_rows > 0) {
        sell_row_len_kernel<T><<<grid_for(n_rows), kBlock, 0, ctx->stream>>>(data, indices, row_indices, n_rows, ptr);
        B200_LAUNCH_CHECK();
    }
    scan_i32_kernel<<<1, 1024, 0, ctx->stream>>>(ptr, n_rows);
    B200_LAUNCH_CHECK();
    if (cols && vals && n_rows > 0) {
        sell_to_csr_fill_kernel<T><<<grid_for(n_rows), kBlock, 0, ctx->stream>>>(data, indices, row_indices, n_rows, ptr, cols, vals);
        B200_LAUNCH_CHECK();
    }
    if (nnz) {
        int last = 0;
        B200_CUDA(cudaMemcpyAsync(&last, ptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *nnz = last;
    }
    return B200_SUCCESS;
}

}  // namespace

extern "C" {

int b200_sell_pack16_f64(b200_ctx *ctx, const double *data, const int *indices, const int *row_indices, int n_slices,
                         int n_cols, int *chunk_base, unsigned short *delta16)
{
    return sell_pack16_impl<double>(ctx, data, indices, row_indices, n_slices, n_cols, chunk_base, delta16);
}
int b200_sell_pack16_f32(b200_ctx *ctx, const float *data, const int *indices, const int *row_indices, int n_slices,
                         int n_cols, int *chunk_base, unsigned short *delta16)
{
    return sell_pack16_impl<float>(ctx, data, indices, row_indices, n_slices, n_cols, chunk_base, delta16);
}
int b200_spmv_sell16_f64(b200_ctx *ctx, const double *data, const unsigned short *delta16, const int *chunk_base,
                         const double *vect, double *output, const int *row_indices, int chunk, int n_slices, int n_out,
                         int n_cols)
{
    return spmv_sell16_impl<double>(ctx, data, delta16, chunk_base, vect, output, row_indices, chunk, n_slices, n_out, n_cols);
}
int b200_spmv_sell16_f32(b200_ctx *ctx, const float *data, const unsigned short *delta16, const int *chunk_base,
                         const float *vect, float *output, const int *row_indices, int chunk, int n_slices, int n_out,
                         int n_cols)
{
    return spmv_sell16_impl<float>(ctx, data, delta16, chunk_base, vect, output, row_indices, chunk, n_slices, n_out, n_cols);
}

int b200_ell_to_csr_f64(b200_ctx *ctx, const double *data, const int *indices, int n_rows, int row_size, int *ptr, int *cols,
                        double *vals, long long *nnz)
{
    return ell_to_csr_impl<double>(ctx, data, indices, n_rows, row_size, ptr, cols, vals, nnz);
}
int b200_ell_to_csr_f32(b200_ctx *ctx, const float *data, const int *indices, int n_rows, int row_size, int *ptr, int *cols,
                        float *vals, long long *nnz)
{
    return ell_to_csr_impl<float>(ctx, data, indices, n_rows, row_size, ptr, cols, vals, nnz);
}
int b200_sell_to_csr_f64(b200_ctx *ctx, const double *data, const int *indices, const int *row_indices, int n_rows, int *ptr,
                         int *cols, double *vals, long long *nnz)
{
    return sell_to_csr_impl<double>(ctx, data, indices, row_indices, n_rows, ptr, cols, vals, nnz);
}
int b200_sell_to_csr_f32(b200_ctx *ctx, const float *data, const int *indices, const int *row_indices, int n_rows, int *ptr,
                         int *cols, float *vals, long long *nnz)
{
    return sell_to_csr_impl<float>(ctx, data, indices, row_indices, n_rows, ptr, cols, vals, nnz);
}

int b200_cmrs_to_csr_ptr(b200_ctx *ctx, const int *strip_ptr, const int *row_in_strip, int n_strips, int height, int n_rows,
                         int *ptr)
{
    B200_TRACE("b200 cmrs -> csr");
    B200_ENTER(ctx);
    B200_REQUIRE(strip_ptr && ptr && n_strips >= 0 && height >= 1 && n_rows >= 0, "bad argument");
    B200_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int) * ((size_t)n_rows + 1), ctx->stream));
    int *flag = ctx->scratch + 321;
    B200_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    if (n_strips > 0) {
        B200_REQUIRE(row_in_strip, "null row_in_strip");
        cmrs_row_count_kernel<<<grid_for((long long)n_strips * 32), kBlock, 0, ctx->stream>>>(strip_ptr, row_in_strip, n_strips,
                                                                                            height, n_rows, ptr, flag);
        B200_LAUNCH_CHECK();
    }
    scan_i32_kernel<<<1, 1024, 0, ctx->stream>>>(ptr, n_rows);
    B200_LAUNCH_CHECK();
    int unsorted = 0;
    B200_CUDA(cudaMemcpyAsync(&unsorted, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (unsorted) {
        b200_set_error("CMRS entries are not sorted by row inside a strip: indices/data are not in CSR order");
        return B200_ERR_DOMAIN;
    }
    return B200_SUCCESS;
}

// ---- format advice ------------------------------------------------------------------------------
int b200_format_advice(b200_ctx *ctx, const int *ptr, int n_rows, int n_cols, int value_bytes, b200_format_advice_t *out)
{
    B200_TRACE("b200 format advice");
    B200_ENTER(ctx);
    B200_REQUIRE(ptr && out && n_rows >= 0 && n_cols >= 0 && (value_bytes == 4 || value_bytes == 8), "bad argument");
    memset(out, 0, sizeof *out);
    b200_row_stats st;
    int rc = b200_row_length_stats(ctx, ptr, n_rows, &st);
    if (rc) return rc;
    const long long nnz = st.sum_len;
    const long long V = value_bytes, I = 4;
    const long long xy = (long long)n_cols * V + (long long)n_rows * V;
    // padded size of SELL-32 at sigma = 1 and with rows sorted globally (a lower bound for any sigma)
    long long p1 = 0, psorted = 0;
    const int n_slices = b200_sell_num_slices(n_rows, 32);
    if (n_rows > 0) {
        long long *sp = nullptr;
        int *perm = nullptr;
        B200_CUDA(cudaMalloc(&sp, sizeof(long long) * ((size_t)n_slices + 1)));
        cudaError_t e = cudaMalloc(&perm, sizeof(int) * (size_t)n_rows);
        if (e != cudaSuccess) {
            cudaFree(sp);
            return b200_cuda_fail(e, "cudaMalloc(perm)", __FILE__, __LINE__);
        }
        rc = b200_build_sell_ptr(ctx, ptr, n_rows, 32, 1, nullptr, sp, &p1);
        if (rc == B200_SUCCESS) rc = b200_build_sell_ptr(ctx, ptr, n_rows, 32, 65536, perm, sp, &psorted);
        cudaFree(sp);
        cudaFree(perm);
        if (rc) return rc;
    }
    const int T = b200_cmrs_num_strips(n_rows, 8);
    out->n_rows = n_rows;
    out->nnz = nnz;
    out->min_len = st.min_len;
    out->max_len = st.max_len;
    out->mean_len = n_rows > 0 ? (double)nnz / n_rows : 0.0;
    out->sell_padding = nnz > 0 ? (double)p1 / (double)nnz : 1.0;
    out->sell_padding_sigma65536 = nnz > 0 ? (double)psorted / (double)nnz : 1.0;
    out->bytes[B200_FORMAT_COO] = nnz * (2 * I + V) + xy;
    out->bytes[B200_FORMAT_CSR] = nnz * (I + V) + ((long long)n_rows + 1) * I + xy;
    out->bytes[B200_FORMAT_ELL] = (long long)n_rows * st.max_len * (I + V) + xy;
    out->bytes[B200_FORMAT_SELL] = p1 * (I + V) + ((long long)n_slices + 1) * I + xy;
    out->bytes[B200_FORMAT_CMRS] = nnz * (2 * I + V) + ((long long)T + 1) * I + xy;
    out->bytes_sell_sigma65536 = psorted * (I + V) + ((long long)n_slices + 1) * 8 + (long long)n_rows * I + xy;
    out->bytes_sell16 = p1 * (2 + V) + ((long long)n_slices + 1) * I + (long long)n_slices * I + xy;
    // Every kernel streams its bytes at 0.93-1.0 of the HBM peak on regular matrices (DESIGN.md section
    // 4), so fewest bytes wins there.  Skewed (power-law) matrices are gather-bound, not byte-bound: the
    // measured order on R-MAT is sigma-sorted SELL >= COO > CSR (profiles/), as long as sorting brings
    // the padding under ~1.3; ELL is never right for them.
    const bool skewed = out->mean_len > 0 && (double)st.max_len > 16.0 * out->mean_len;
    out->skewed = skewed;
    if (skewed) {
        if (out->sell_padding_sigma65536 <= 1.3) {
            out->recommended = B200_FORMAT_SELL;
            out->recommended_sigma = 65536;
            snprintf(out->reason, sizeof out->reason,
                     "skewed rows (max %d vs mean %.1f): gather-bound; SELL-32 with sigma = 65536 pads only %.2fx",
                     st.max_len, out->mean_len, out->sell_padding_sigma65536);
        } else {
            out->recommended = B200_FORMAT_CSR;
            out->recommended_sigma = 1;
            snprintf(out->reason, sizeof out->reason,
                     "skewed rows (max %d vs mean %.1f) and SELL pads %.2fx even when sorted: CSR (nnz-split kernel)",
                     st.max_len, out->mean_len, out->sell_padding_sigma65536);
        }
        return B200_SUCCESS;
    }
    int best = B200_FORMAT_CSR;
    for (int f = 0; f < 5; ++f)
        if (out->bytes[f] < out->bytes[best]) best = f;
    // within 2 % of CSR, SELL wins: its chunk is a warp's worth of aligned 128-bit loads with no ragged
    // row ends (measured 1.00 vs 0.93-0.97 of the peak on the banded matrix)
    if (out->bytes[B200_FORMAT_SELL] <= out->bytes[best] + out->bytes[best] / 50) best = B200_FORMAT_SELL;
    out->recommended = best;
    out->recommended_sigma = 1;
    snprintf(out->reason, sizeof out->reason, "regular rows (%d..%d, mean %.1f): fewest bytes per SpMV; SELL padding %.3fx",
             st.min_len, st.max_len, out->mean_len, out->sell_padding);
    return B200_SUCCESS;
}

}  // extern "C"
