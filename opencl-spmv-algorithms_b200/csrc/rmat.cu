// rmat.cu -- device generator for the power-law workload (BASELINE.json configs[3]).
//
// R-MAT: n = 2^scale rows/columns, n*edge_factor edges; every edge picks one quadrant per bit
// level with probabilities (a, b, c, 1-a-b-c).  One diagonal entry per row is added so that no
// row is empty (the reference's builders are only defined without empty rows),
// duplicates are removed and the triples are sorted by (row, col).  Each edge is a pure function
// of (seed, edge index), so any rank can generate exactly its own row block [row_begin,
// row_begin+row_count) of the same global matrix.
//
// This is input generation, not the SpMV hot path: the sort and the duplicate removal use CUB
// (cub::DeviceRadixSort / cub::DeviceSelect, header-only, shipped with the CUDA toolkit).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

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

struct RmatParams {
    int scale;
    unsigned ta, tb, tc;  // 16-bit cumulative thresholds of a, a+b, a+b+c
    uint64_t seed;
};

__device__ __forceinline__ void rmat_edge(const RmatParams &p, long long e, unsigned &row, unsigned &col)
{
    unsigned r = 0, c = 0;
    uint64_t h = 0;
    for (int level = 0; level < p.scale; ++level) {
        if ((level & 3) == 0) h = mix64(p.seed ^ mix64((uint64_t)e * 8 + (uint64_t)(level >> 2)));
        const unsigned u = (unsigned)(h & 0xffffu);
        h >>= 16;
        const unsigned quad = u < p.ta ? 0u : (u < p.tb ? 1u : (u < p.tc ? 2u : 3u));
        r = (r << 1) | (quad >> 1);
        c = (c << 1) | (quad & 1u);
    }
    row = r;
    col = c;
}

// pass 1 (keys == nullptr): count the edges of the row block; pass 2: append their keys
__global__ void rmat_edges_kernel(RmatParams p, long long n_edges, unsigned row_begin, unsigned row_end,
                                  unsigned long long *counter, unsigned long long *keys,
                                  unsigned long long capacity)
{
    unsigned long long local = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges;
         e += (long long)gridDim.x * blockDim.x) {
        unsigned r, c;
        rmat_edge(p, e, r, c);
        if (r >= row_begin && r < row_end) {
            if (keys) {
                const unsigned long long at = atomicAdd(counter, 1ull);
                if (at < capacity) keys[at] = ((unsigned long long)r << 32) | c;
            } else {
                ++local;
            }
        }
    }
    if (!keys) {
        for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, local);
    }
}

__global__ void rmat_diagonal_kernel(unsigned row_begin, unsigned row_count, unsigned long long *keys)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < row_count) {
        const unsigned long long r = row_begin + (unsigned)i;
        keys[i] = (r << 32) | r;
    }
}

__global__ void rmat_unpack_kernel(const unsigned long long *__restrict__ keys, long long n, uint64_t seed,
                                   int *__restrict__ rows, int *__restrict__ cols, double *__restrict__ vals)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        rows[i] = (int)(k >> 32);
        cols[i] = (int)(k & 0xffffffffu);
        const uint64_t h = mix64(seed ^ mix64(k + 0x51ed270b7f4a7c15ull));
        double v = 2.0 * ((double)(h >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
        vals[i] = v == 0.0 ? 0.5 : v;
    }
}

int make_params(int scale, double a, double b, double c, uint64_t seed, RmatParams *p)
{
    B200_REQUIRE(scale >= 1 && scale <= 30, "scale must be in 1..30");
    B200_REQUIRE(a > 0 && b >= 0 && c >= 0 && a + b + c < 1.0, "bad quadrant probabilities");
    p->scale = scale;
    p->ta = (unsigned)(a * 65536.0);
    p->tb = (unsigned)((a + b) * 65536.0);
    p->tc = (unsigned)((a + b + c) * 65536.0);
    p->seed = seed;
    return B200_SUCCESS;
}

unsigned edge_grid(const b200_ctx *ctx, long long n)
{
    long long b = (n + kBlock - 1) / kBlock, cap = (long long)ctx->sm_count * 32;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int b200_gen_rmat_count(b200_ctx *ctx, int scale, int edge_factor, double a, double b, double c,
                        uint64_t seed, int row_begin, int row_count, long long *n_candidates)
{
    B200_ENTER(ctx);
    RmatParams p;
    int rc = make_params(scale, a, b, c, seed, &p);
    if (rc) return rc;
    B200_REQUIRE(n_candidates && edge_factor >= 1 && row_begin >= 0 && row_count >= 0, "bad argument");
    B200_REQUIRE((long long)row_begin + row_count <= (1ll << scale), "row block outside the matrix");
    const long long n_edges = (1ll << scale) * edge_factor;
    unsigned long long *counter = reinterpret_cast<unsigned long long *>(ctx->scratch + 128);
    B200_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
    rmat_edges_kernel<<<edge_grid(ctx, n_edges), kBlock, 0, ctx->stream>>>(
        p, n_edges, (unsigned)row_begin, (unsigned)(row_begin + row_count), counter, nullptr, 0);
    B200_LAUNCH_CHECK();
    unsigned long long got = 0;
    B200_CUDA(cudaMemcpyAsync(&got, counter, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_candidates = (long long)got + row_count;  // + one diagonal entry per row
    return B200_SUCCESS;
}

int b200_gen_rmat_coo(b200_ctx *ctx, int scale, int edge_factor, double a, double b, double c,
                      uint64_t seed, int row_begin, int row_count, long long capacity, int *rows,
                      int *cols, double *vals, long long *nnz_out)
{
    B200_ENTER(ctx);
    RmatParams p;
    int rc = make_params(scale, a, b, c, seed, &p);
    if (rc) return rc;
    B200_REQUIRE(nnz_out && rows && cols && vals && edge_factor >= 1 && row_begin >= 0 && row_count >= 0,
                 "bad argument");
    B200_REQUIRE((long long)row_begin + row_count <= (1ll << scale), "row block outside the matrix");
    B200_REQUIRE(capacity >= row_count, "capacity smaller than the diagonal");
    *nnz_out = 0;
    const long long n_edges = (1ll << scale) * edge_factor;
    unsigned long long *keys = nullptr, *sorted = nullptr, *n_unique = nullptr;
    void *temp = nullptr;
    size_t temp_sort = 0, temp_sel = 0;
    cudaError_t e = cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)capacity);
    if (e == cudaSuccess) e = cudaMalloc(&sorted, sizeof(unsigned long long) * (size_t)capacity);
    if (e == cudaSuccess) e = cudaMalloc(&n_unique, sizeof(unsigned long long));
    unsigned long long *counter = reinterpret_cast<unsigned long long *>(ctx->scratch + 128);
    unsigned long long got = 0;
    long long n_keys = 0;
    if (e == cudaSuccess) {
        // diagonal first (slots 0..row_count-1), edges appended after it
        const unsigned long long first = (unsigned long long)row_count;
        e = cudaMemcpyAsync(counter, &first, sizeof first, cudaMemcpyHostToDevice, ctx->stream);
        if (row_count > 0)
            rmat_diagonal_kernel<<<(row_count + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(
                (unsigned)row_begin, (unsigned)row_count, keys);
        rmat_edges_kernel<<<edge_grid(ctx, n_edges), kBlock, 0, ctx->stream>>>(
            p, n_edges, (unsigned)row_begin, (unsigned)(row_begin + row_count), counter, keys,
            (unsigned long long)capacity);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&got, counter, sizeof got, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e == cudaSuccess && (long long)got > capacity) {
        cudaFree(keys);
        cudaFree(sorted);
        cudaFree(n_unique);
        b200_set_error("capacity %lld too small: the row block has %llu candidate entries "
                       "(call b200_gen_rmat_count first)", capacity, got);
        return B200_ERR_INVALID_VALUE;
    }
    n_keys = (long long)got;
    if (e == cudaSuccess && n_keys > 0x7fffffffll) {
        cudaFree(keys);
        cudaFree(sorted);
        cudaFree(n_unique);
        b200_set_error("row block has more than 2^31-1 entries; use more, smaller row blocks");
        return B200_ERR_UNSUPPORTED;
    }
    if (e == cudaSuccess) {
        cub::DeviceRadixSort::SortKeys(nullptr, temp_sort, keys, sorted, (int)n_keys, 0, 32 + scale, ctx->stream);
        cub::DeviceSelect::Unique(nullptr, temp_sel, sorted, keys, n_unique, (int)n_keys, ctx->stream);
        e = cudaMalloc(&temp, temp_sort > temp_sel ? temp_sort : temp_sel);
    }
    if (e == cudaSuccess)
        e = cub::DeviceRadixSort::SortKeys(temp, temp_sort, keys, sorted, (int)n_keys, 0, 32 + scale, ctx->stream);
    if (e == cudaSuccess)
        e = cub::DeviceSelect::Unique(temp, temp_sel, sorted, keys, n_unique, (int)n_keys, ctx->stream);
    unsigned long long uniq = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&uniq, n_unique, sizeof uniq, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && uniq > 0) {
        rmat_unpack_kernel<<<edge_grid(ctx, (long long)uniq), kBlock, 0, ctx->stream>>>(
            keys, (long long)uniq, seed, rows, cols, vals);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(temp);
    cudaFree(keys);
    cudaFree(sorted);
    cudaFree(n_unique);
    if (e != cudaSuccess) return b200_cuda_fail(e, "R-MAT generator", __FILE__, __LINE__);
    *nnz_out = (long long)uniq;
    return B200_SUCCESS;
}

}  // extern "C"
