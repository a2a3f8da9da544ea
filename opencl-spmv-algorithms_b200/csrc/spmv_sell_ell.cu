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
// Column-major ELL is the same idea with one "chunk" spanning the whole matrix: thread t owns rows
// 4t..4t+3 and walks the K columns (optionally split over KS thread-slices when the matrix has
// too few rows to fill 148 SMs).  Bytes: P*(4+V) + (S+1)*P_bytes [+R*4 perm] + Cn*V + R*V for
// SELL, R*K*(4+V) + Cn*V + R*V for ELL; both HBM-bound.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;

template <typename T, typename P>
__global__ void __launch_bounds__(kBlock)
sell32_kernel(const T *__restrict__ data, const int *__restrict__ idx, const T *__restrict__ x,
              T *__restrict__ y, const P *__restrict__ slice_ptr, int n_slices, int n_out,
              const int *__restrict__ perm)
{
    const int lane = threadIdx.x & 31;
    const long long slice = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (slice >= n_slices) return;  // whole warps leave together
    const long long base = slice_ptr[slice];
    const long long n_groups = ((long long)slice_ptr[slice + 1] - base) >> 2;  // 8 per column
    const int *ip = idx + base;
    const T *dp = data + base;
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
    if (lane < 8) {
        const long long r = slice * 32 + lane * 4;
        const T a[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r + k < n_out) y[perm ? perm[r + k] : r + k] = a[k];
    }
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

template <typename T, typename P>
int spmv_sell_impl(b200_ctx *ctx, const T *data, const int *idx, const T *x, T *y, const P *slice_ptr,
                   int chunk, int n_slices, int n_out, const int *perm)
{
    B200_ENTER(ctx);
    B200_REQUIRE(x && y && slice_ptr && n_slices >= 0 && n_out >= 0, "bad argument");
    if (chunk != 32) {
        b200_set_error("SELL chunk must be 32 (warp-aligned), got %d", chunk);
        return B200_ERR_UNSUPPORTED;
    }
    B200_REQUIRE((long long)n_out <= (long long)n_slices * 32, "n_out exceeds n_slices*32");
    if (n_slices == 0) return B200_SUCCESS;
    unsigned blocks = ceil_div_u((long long)n_slices * 32, kBlock);
    if (aligned16(data) && aligned16(idx))
        sell32_kernel<T, P><<<blocks, kBlock, 0, ctx->stream>>>(data, idx, x, y, slice_ptr, n_slices, n_out, perm);
    else
        sell32_scalar_kernel<T, P><<<blocks, kBlock, 0, ctx->stream>>>(data, idx, x, y, slice_ptr, n_slices, n_out, perm);
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
    if (const char *e = getenv("B200_ELLCM_Q")) q = (KS == 1 && atoi(e) == 2) ? 2 : 1;
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

extern "C" {

int b200_spmv_sell_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                       double *output, const int *row_indices, int chunk, int n_slices, int n_out,
                       const int *perm)
{
    return spmv_sell_impl<double, int>(ctx, data, indices, vect, output, row_indices, chunk, n_slices, n_out, perm);
}
int b200_spmv_sell_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                       float *output, const int *row_indices, int chunk, int n_slices, int n_out,
                       const int *perm)
{
    return spmv_sell_impl<float, int>(ctx, data, indices, vect, output, row_indices, chunk, n_slices, n_out, perm);
}
int b200_spmv_sell64_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                         double *output, const long long *slice_ptr, int chunk, int n_slices,
                         int n_out, const int *perm)
{
    return spmv_sell_impl<double, long long>(ctx, data, indices, vect, output, slice_ptr, chunk, n_slices, n_out, perm);
}
int b200_spmv_sell64_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                         float *output, const long long *slice_ptr, int chunk, int n_slices,
                         int n_out, const int *perm)
{
    return spmv_sell_impl<float, long long>(ctx, data, indices, vect, output, slice_ptr, chunk, n_slices, n_out, perm);
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
