// spmv_cmrs_coo.cu -- CMRS and COO SpMV for sm_100a.
//
// CMRS replaces kernels/Cmrs.cl:1-46 (32-lane group per strip, a read-modify-write of local
// memory per non-zero, 3 barriers per strip).  B200 design: one warp owns one strip (`height`
// consecutive rows = one contiguous run of non-zeros).  Lanes stream the run with 128-bit loads
// of values, column indices and row_in_strip; each lane keeps `height` private accumulators in
// REGISTERS (selected by compare, no indexing), a transposing butterfly of __shfl_xor_sync folds
// the 32 x height partials, and the strip's results are staged in shared memory so that a block
// (8 strips = 64 rows at height 8) writes y as one coalesced run.
//
// COO replaces kernels/Coo.cl:4-32 (one CAS-loop atomic per non-zero).  B200 design: each thread
// takes four consecutive entries (128-bit loads of row, col, value), folds equal-row neighbours
// locally, then a ballot-delimited segmented scan across the warp merges runs that span lanes;
// only the last lane of each run issues a (native fp32/fp64) atomicAdd.  Row-sorted input costs
// about one atomic per (warp, row); arbitrary order degrades gracefully to one per entry.
//
// Bytes: nnz*(8+V) + (T+1)*4 + Cn*V + R*V (CMRS), nnz*(8+V) + Cn*V + R*V (COO); HBM-bound.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

// HMAX = accumulator registers per lane (8 or 32); height <= HMAX at run time
template <typename T, int HMAX, bool VEC>
__global__ void __launch_bounds__(kBlock)
cmrs_kernel(const T *__restrict__ data, const int *__restrict__ idx, const int *__restrict__ strip_ptr,
            const int *__restrict__ row_in_strip, const T *__restrict__ x, T *__restrict__ y,
            int n_strips, int height, int n_rows)
{
    __shared__ T stage[kWarps][HMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long strip = (long long)blockIdx.x * kWarps + warp;
    T acc[HMAX];
#pragma unroll
    for (int h = 0; h < HMAX; ++h) acc[h] = 0;
    if (strip < n_strips) {
        const int s = __ldg(strip_ptr + strip), e = __ldg(strip_ptr + strip + 1);
        if (VEC) {
            const int g_end = (e + 3) >> 2;
#pragma unroll 2
            for (int g = (s >> 2) + lane; g < g_end; g += 32) {
                const int j = g << 2;
                IVec4 c, r;
                Vec4<T> v;
                c.load(idx + j);
                r.load(row_in_strip + j);
                v.load(data + j);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (j + k >= s && j + k < e) {
                        const T p = v.v[k] * ld_x(x, c.v[k]);
#pragma unroll
                        for (int h = 0; h < HMAX; ++h) acc[h] += (r.v[k] == h) ? p : T(0);
                    }
                }
            }
        } else {
            for (int j = s + lane; j < e; j += 32) {
                const T p = ld_stream(data + j) * ld_x(x, ld_stream(idx + j));
                const int r = ld_stream(row_in_strip + j);
#pragma unroll
                for (int h = 0; h < HMAX; ++h) acc[h] += (r == h) ? p : T(0);
            }
        }
    }
    // transposing butterfly: after the HMAX-halving steps lane L holds the partial of row
    // (L % HMAX) over the lanes congruent to it; plain xor steps finish the sum.
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
    // lane L (< HMAX) now owns row bitrev-free index: row = L's bits interpreted directly
    if (lane < HMAX) stage[warp][lane] = total;
    __syncthreads();
    // coalesced store of the block's kWarps*height rows
    const long long row0 = (long long)blockIdx.x * kWarps * height;
    for (int t = threadIdx.x; t < kWarps * height; t += kBlock) {
        const int w = t / height, h = t - w * height;
        const long long row = row0 + t;
        if ((long long)blockIdx.x * kWarps + w < n_strips && row < n_rows) y[row] = stage[w][h];
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
coo_kernel(const int *__restrict__ row, const int *__restrict__ col, const T *__restrict__ data,
           const T *__restrict__ x, T *__restrict__ y, int nnz, bool vec)
{
    const int lane = threadIdx.x & 31;
    const long long j0 = ((long long)blockIdx.x * kBlock + threadIdx.x) * 4;
    int r[4], c[4];
    T v[4];
    if (vec && j0 < nnz) {
        IVec4 rr, cc;
        Vec4<T> vv;
        rr.load(row + j0);
        cc.load(col + j0);
        vv.load(data + j0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = rr.v[k];
            c[k] = cc.v[k];
            v[k] = vv.v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool in = j0 + k < nnz;
            r[k] = in ? ld_stream(row + j0 + k) : -1;
            c[k] = in ? ld_stream(col + j0 + k) : 0;
            v[k] = in ? ld_stream(data + j0 + k) : T(0);
        }
    }
    // thread-local fold; completed interior runs go straight to memory
    int cur = -1;
    T sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (j0 + k >= nnz) break;
        const T p = v[k] * ld_x(x, c[k]);
        if (r[k] == cur) {
            sum += p;
        } else {
            if (cur >= 0) atomicAdd(y + cur, sum);
            cur = r[k];
            sum = p;
        }
    }
    // warp-level segmented inclusive scan over (cur, sum); heads delimit runs of equal rows
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

template <typename T>
int spmv_cmrs_impl(b200_ctx *ctx, const T *data, const int *idx, const int *strip_ptr,
                   const int *row_in_strip, const T *x, T *y, int n_strips, int height, int n_rows)
{
    B200_ENTER(ctx);
    B200_REQUIRE(strip_ptr && x && y && n_strips >= 0 && n_rows >= 0, "bad argument");
    if (height < 1 || height > 32) {
        b200_set_error("CMRS height must be in 1..32, got %d", height);
        return B200_ERR_UNSUPPORTED;
    }
    if (n_strips == 0) return B200_SUCCESS;
    const bool vec = aligned16(data) && aligned16(idx) && aligned16(row_in_strip);
    unsigned blocks = ceil_div_u(n_strips, kWarps);
#define B200_CMRS_LAUNCH(H, V)                                                                   \
    cmrs_kernel<T, H, V><<<blocks, kBlock, 0, ctx->stream>>>(data, idx, strip_ptr, row_in_strip, \
                                                             x, y, n_strips, height, n_rows)
    if (height <= 8) {
        if (vec) B200_CMRS_LAUNCH(8, true);
        else B200_CMRS_LAUNCH(8, false);
    } else {
        if (vec) B200_CMRS_LAUNCH(32, true);
        else B200_CMRS_LAUNCH(32, false);
    }
#undef B200_CMRS_LAUNCH
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <typename T>
int spmv_coo_impl(b200_ctx *ctx, const int *row, const int *col, const T *data, const T *x, T *y,
                  int nnz, int n_rows)
{
    B200_ENTER(ctx);
    B200_REQUIRE(x && y && nnz >= 0 && n_rows >= 0, "bad argument");
    B200_REQUIRE(nnz == 0 || (row && col && data), "null row/col/data");
    B200_CUDA(cudaMemsetAsync(y, 0, sizeof(T) * (size_t)n_rows, ctx->stream));
    if (nnz == 0) return B200_SUCCESS;
    const bool vec = aligned16(row) && aligned16(col) && aligned16(data);
    unsigned blocks = ceil_div_u(((long long)nnz + 3) / 4, kBlock);
    coo_kernel<T><<<blocks, kBlock, 0, ctx->stream>>>(row, col, data, x, y, nnz, vec);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // namespace

extern "C" {

int b200_spmv_cmrs_f64(b200_ctx *ctx, const double *data, const int *indices, const int *strip_ptr,
                       const int *row_in_strip, const double *vect, double *output, int n_strips,
                       int height, int n_rows)
{
    return spmv_cmrs_impl<double>(ctx, data, indices, strip_ptr, row_in_strip, vect, output, n_strips, height, n_rows);
}
int b200_spmv_cmrs_f32(b200_ctx *ctx, const float *data, const int *indices, const int *strip_ptr,
                       const int *row_in_strip, const float *vect, float *output, int n_strips,
                       int height, int n_rows)
{
    return spmv_cmrs_impl<float>(ctx, data, indices, strip_ptr, row_in_strip, vect, output, n_strips, height, n_rows);
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
