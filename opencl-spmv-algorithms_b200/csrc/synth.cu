// synth.cu -- synthetic matrices generated on the device (BASELINE.json configs 3 and 5), the
// multi-GPU row partition and the two vector helpers of the iterated (power-iteration) mode.
//
// The reference can only ingest MatrixMarket text (csr.c:77-91), which is > 99 % of its
// wall-clock; 10^8-nnz inputs are generated directly in HBM instead.  Every generator is a pure
// function of (seed, global row, slot) built on a 64-bit integer mix, so any row block of the
// same global matrix can be produced independently on any rank -- and on the host
// (b200_gen_*_host) bit-identically, which is how the CPU baseline sees the same matrix.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// uniform in [0,1) with 53 random bits
__host__ __device__ inline double unit53(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// banded FEM-like pattern: nnz_per_row entries per row in groups of 4 consecutive columns (dof
// blocks), group offsets spread evenly over [-half_band, half_band-4], columns wrap mod n.
__host__ __device__ inline int banded_col(int n, int row, int slot, int nnz_per_row, int half_band)
{
    const int groups = (nnz_per_row + 3) / 4;
    const int g = slot >> 2, j = slot & 3;
    const long long span = 2ll * half_band - 4;
    const long long off = -(long long)half_band + (groups > 1 ? span * g / (groups - 1) : 0) + j;
    long long c = ((long long)row + off) % n;
    if (c < 0) c += n;
    return (int)c;
}

__host__ __device__ inline double banded_val(uint64_t seed, int row, int slot, int nnz_per_row)
{
    const uint64_t h = mix64(seed ^ mix64((uint64_t)row * (uint64_t)nnz_per_row + (uint64_t)slot));
    double v = 2.0 * unit53(h) - 1.0;
    return v == 0.0 ? 0.5 : v;
}

__global__ void banded_kernel(int n, int row_begin, int row_count, int npr, int half_band, uint64_t seed,
                              int *__restrict__ rows, int *__restrict__ cols, double *__restrict__ vals)
{
    const long long total = (long long)row_count * npr;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int r = row_begin + (int)(e / npr), k = (int)(e % npr);
        rows[e] = r;
        cols[e] = banded_col(n, r, k, npr, half_band);
        vals[e] = banded_val(seed, r, k, npr);
    }
}

// 7-point Laplacian on nx*ny*nz (x fastest): diag 6, neighbours -1, columns ascending
__host__ __device__ inline int lap7_count(int nx, int ny, int nz, long long row)
{
    const int x = (int)(row % nx), y = (int)((row / nx) % ny), z = (int)(row / ((long long)nx * ny));
    return 1 + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
}

__global__ void lap7_count_kernel(int nx, int ny, int nz, int row_begin, int row_count,
                                  long long *__restrict__ counts)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < row_count) counts[i] = lap7_count(nx, ny, nz, (long long)row_begin + i);
}

__global__ void lap7_fill_kernel(int nx, int ny, int nz, int row_begin, int row_count,
                                 const long long *__restrict__ start, int *__restrict__ rows,
                                 int *__restrict__ cols, double *__restrict__ vals)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_count) return;
    const long long row = (long long)row_begin + i;
    const int x = (int)(row % nx), y = (int)((row / nx) % ny), z = (int)(row / ((long long)nx * ny));
    const long long sxy = (long long)nx * ny;
    long long at = start[i];
    auto put = [&](long long c, double v) {
        rows[at] = (int)row;
        cols[at] = (int)c;
        vals[at] = v;
        ++at;
    };
    if (z > 0) put(row - sxy, -1.0);
    if (y > 0) put(row - nx, -1.0);
    if (x > 0) put(row - 1, -1.0);
    put(row, 6.0);
    if (x < nx - 1) put(row + 1, -1.0);
    if (y < ny - 1) put(row + nx, -1.0);
    if (z < nz - 1) put(row + sxy, -1.0);
}

__global__ void __launch_bounds__(1024) scan_counts_kernel(long long *__restrict__ a, long long n)
{
    // single-block exclusive scan (same scheme as build_formats.cu; a has n+1 slots)
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
            warp_tot[lane] = w;
        }
        __syncthreads();
        const long long carry = carry_s;
        if (i < n) a[i] = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = carry_s;
}

template <typename T>
__global__ void uniform_kernel(T *__restrict__ x, long long n, uint64_t seed, T lo, T hi)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] = (T)fma((double)hi - (double)lo, unit53(mix64(seed ^ mix64((uint64_t)i))), (double)lo);
}

// y *= (invert_sqrt ? 1/sqrt(*s) : *s)
__global__ void scale_kernel(double *__restrict__ y, long long n, const double *__restrict__ s, int invert_sqrt)
{
    const double f = invert_sqrt ? rsqrt(*s) : *s;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] *= f;
}

__global__ void sumsq_kernel(const double *__restrict__ y, long long n, double *__restrict__ acc)
{
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        s += y[i] * y[i];
    s = subwarp_sum<32>(s);
    __shared__ double part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
        t = subwarp_sum<32>(t);
        if (threadIdx.x == 0) atomicAdd(acc, t);
    }
}

__global__ void minmax_kernel(const int *__restrict__ a, long long n, int *__restrict__ out)
{
    int lo = 0x7fffffff, hi = (int)0x80000000;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = a[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

inline unsigned capped_grid(const b200_ctx *ctx, long long n)
{
    long long b = (n + kBlock - 1) / kBlock;
    long long cap = (long long)ctx->sm_count * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void col_blocks_kernel(const int *__restrict__ a, long long n, int shift, unsigned char *__restrict__ used)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = a[i] >> shift;
        if (!used[b]) used[b] = 1;  // many threads store the same 1: benign
    }
}

}  // namespace

extern "C" {

long long b200_gen_banded_nnz(long long n_global, int row_begin, int row_count, int nnz_per_row)
{
    (void)n_global;
    (void)row_begin;
    return (long long)row_count * nnz_per_row;
}

int b200_gen_banded_coo(b200_ctx *ctx, int n_global, int row_begin, int row_count, int nnz_per_row,
                        int half_band, uint64_t seed, int *rows, int *cols, double *vals)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n_global > 0 && row_begin >= 0 && row_count >= 0 && nnz_per_row > 0, "bad shape");
    B200_REQUIRE((long long)row_begin + row_count <= n_global, "row block outside the matrix");
    B200_REQUIRE(half_band >= 2 * ((nnz_per_row + 3) / 4) && 2ll * half_band <= n_global,
                 "half_band must satisfy 2*groups <= half_band <= n/2 (distinct columns)");
    B200_REQUIRE((long long)row_count * nnz_per_row <= 0x7fffffffll, "block exceeds 2^31-1 entries");
    if (row_count == 0) return B200_SUCCESS;
    B200_REQUIRE(rows && cols && vals, "null output");
    banded_kernel<<<capped_grid(ctx, (long long)row_count * nnz_per_row), kBlock, 0, ctx->stream>>>(
        n_global, row_begin, row_count, nnz_per_row, half_band, seed, rows, cols, vals);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_gen_banded_coo_host(int n_global, int row_begin, int row_count, int nnz_per_row,
                             int half_band, uint64_t seed, int *rows, int *cols, double *vals)
{
    B200_REQUIRE(n_global > 0 && row_begin >= 0 && row_count >= 0 && nnz_per_row > 0, "bad shape");
    B200_REQUIRE(rows && cols && vals, "null output");
    const long long total = (long long)row_count * nnz_per_row;
#pragma omp parallel for
    for (long long e = 0; e < total; ++e) {
        const int r = row_begin + (int)(e / nnz_per_row), k = (int)(e % nnz_per_row);
        rows[e] = r;
        cols[e] = banded_col(n_global, r, k, nnz_per_row, half_band);
        vals[e] = banded_val(seed, r, k, nnz_per_row);
    }
    return B200_SUCCESS;
}

long long b200_gen_laplace7_nnz(int nx, int ny, int nz, int row_begin, int row_count)
{
    long long total = 0;
    for (long long r = row_begin; r < (long long)row_begin + row_count; ++r)
        total += lap7_count(nx, ny, nz, r);
    return total;
}

int b200_gen_laplace7_coo(b200_ctx *ctx, int nx, int ny, int nz, int row_begin, int row_count,
                          int *rows, int *cols, double *vals)
{
    B200_ENTER(ctx);
    B200_REQUIRE(nx > 0 && ny > 0 && nz > 0 && row_begin >= 0 && row_count >= 0, "bad shape");
    B200_REQUIRE((long long)nx * ny * nz <= 0x7fffffffll, "grid exceeds 2^31-1 rows");
    B200_REQUIRE((long long)row_begin + row_count <= (long long)nx * ny * nz, "row block outside the grid");
    if (row_count == 0) return B200_SUCCESS;
    B200_REQUIRE(rows && cols && vals, "null output");
    long long *start = nullptr;
    B200_CUDA(cudaMalloc(&start, sizeof(long long) * ((size_t)row_count + 1)));
    const unsigned blocks = (unsigned)(((long long)row_count + kBlock - 1) / kBlock);
    lap7_count_kernel<<<blocks, kBlock, 0, ctx->stream>>>(nx, ny, nz, row_begin, row_count, start);
    scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(start, row_count);
    lap7_fill_kernel<<<blocks, kBlock, 0, ctx->stream>>>(nx, ny, nz, row_begin, row_count, start, rows, cols, vals);
    cudaError_t e = cudaGetLastError();
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    cudaFree(start);
    if (e != cudaSuccess) return b200_cuda_fail(e, "laplace7 generator", __FILE__, __LINE__);
    if (es != cudaSuccess) return b200_cuda_fail(es, "laplace7 generator", __FILE__, __LINE__);
    return B200_SUCCESS;
}

int b200_gen_uniform_f64(b200_ctx *ctx, double *x, long long n, uint64_t seed, double lo, double hi)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && (n == 0 || x), "bad argument");
    if (n == 0) return B200_SUCCESS;
    uniform_kernel<double><<<capped_grid(ctx, n), kBlock, 0, ctx->stream>>>(x, n, seed, lo, hi);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_gen_uniform_f32(b200_ctx *ctx, float *x, long long n, uint64_t seed, float lo, float hi)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && (n == 0 || x), "bad argument");
    if (n == 0) return B200_SUCCESS;
    uniform_kernel<float><<<capped_grid(ctx, n), kBlock, 0, ctx->stream>>>(x, n, seed, lo, hi);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_gen_uniform_f64_host(double *x, long long n, uint64_t seed, double lo, double hi)
{
    B200_REQUIRE(n >= 0 && (n == 0 || x), "bad argument");
#pragma omp parallel for
    for (long long i = 0; i < n; ++i)
        x[i] = fma(hi - lo, unit53(mix64(seed ^ mix64((uint64_t)i))), lo);
    return B200_SUCCESS;
}

// nnz-balanced contiguous row blocks with aligned cut points
int b200_partition_rows(const int *ptr_host, int n_rows, int n_parts, int align, int *cuts)
{
    B200_REQUIRE(ptr_host && cuts && n_rows >= 0 && n_parts >= 1 && align >= 1, "bad argument");
    const long long base = ptr_host[0], nnz = (long long)ptr_host[n_rows] - base;
    cuts[0] = 0;
    for (int p = 1; p < n_parts; ++p) {
        const long long target = base + nnz * p / n_parts;
        // first row whose start is >= target
        int lo = 0, hi = n_rows;
        while (lo < hi) {
            const int mid = lo + (hi - lo) / 2;
            if (ptr_host[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        long long cut = ((long long)lo + align / 2) / align * align;  // nearest multiple
        if (cut > n_rows) cut = n_rows;
        if (cut < cuts[p - 1]) cut = cuts[p - 1];
        cuts[p] = (int)cut;
    }
    cuts[n_parts] = n_rows;
    return B200_SUCCESS;
}

int b200_scale_f64(b200_ctx *ctx, double *y, long long n, const double *scale_device, int invert_sqrt)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && scale_device && (n == 0 || y), "bad argument");
    if (n == 0) return B200_SUCCESS;
    scale_kernel<<<capped_grid(ctx, n), kBlock, 0, ctx->stream>>>(y, n, scale_device, invert_sqrt);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_sumsq_f64(b200_ctx *ctx, const double *y, long long n, double *acc_device)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && acc_device && (n == 0 || y), "bad argument");
    if (n == 0) return B200_SUCCESS;
    sumsq_kernel<<<capped_grid(ctx, n), kBlock, 0, ctx->stream>>>(y, n, acc_device);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_used_column_blocks(b200_ctx *ctx, const int *cols, long long nnz, int n_cols, int block_log2,
                            unsigned char *used_host)
{
    B200_ENTER(ctx);
    B200_REQUIRE(nnz >= 0 && n_cols >= 0 && block_log2 >= 0 && block_log2 < 31 && used_host && (nnz == 0 || cols), "bad argument");
    const size_t n_blocks = ((size_t)n_cols + ((size_t)1 << block_log2) - 1) >> block_log2;
    memset(used_host, 0, n_blocks);
    if (nnz == 0 || n_blocks == 0) return B200_SUCCESS;
    unsigned char *d = nullptr;
    B200_CUDA(cudaMalloc(&d, n_blocks));
    cudaError_t e = cudaMemsetAsync(d, 0, n_blocks, ctx->stream);
    if (e == cudaSuccess) {
        col_blocks_kernel<<<capped_grid(ctx, nnz), kBlock, 0, ctx->stream>>>(cols, nnz, block_log2, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(used_host, d, n_blocks, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return b200_cuda_fail(e, "b200_used_column_blocks", __FILE__, __LINE__);
    return B200_SUCCESS;
}

int b200_minmax_i32(b200_ctx *ctx, const int *a, long long n, int *min_out, int *max_out)
{
    B200_ENTER(ctx);
    B200_REQUIRE(n >= 0 && min_out && max_out && (n == 0 || a), "bad argument");
    *min_out = 0x7fffffff;
    *max_out = (int)0x80000000;
    if (n == 0) return B200_SUCCESS;
    int *d = ctx->scratch + 256;
    const int init[2] = {0x7fffffff, (int)0x80000000};
    B200_CUDA(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    minmax_kernel<<<capped_grid(ctx, n), kBlock, 0, ctx->stream>>>(a, n, d);
    B200_LAUNCH_CHECK();
    int got[2];
    B200_CUDA(cudaMemcpyAsync(got, d, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *min_out = got[0];
    *max_out = got[1];
    return B200_SUCCESS;
}

}  // extern "C"
