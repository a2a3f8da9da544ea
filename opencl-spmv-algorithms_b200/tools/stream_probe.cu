// stream_probe.cu -- what a plain streaming read reaches at a given launch size (the floor a cant-sized SpMV
// kernel can be compared with: DESIGN.md section 5).
//
//   ./stream_probe [--mb 35,53,70,280,1090] [--copies 7] [--reps 28]
//
// For every size: `copies` independent buffers (rotation keeps every launch cold, as bench.py --workload cant does),
// `reps` back-to-back launches recorded into ONE CUDA graph, timed with events around a replay.  Kernels:
//   read_once     one block per 256 x U 16-byte groups, all U loads of a thread issued before the first is used,
//                 one 4-byte store per block (the shape of the SpMV kernels: short-lived blocks, several waves);
//   read_persist  2-8 resident blocks per SM, grid-stride, the NEXT batch of U loads issued before the current one is
//                 summed (software-pipelined).
// Prints one JSON line: per size and kernel the microseconds per launch and GB/s.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e__));          \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

template <int U>
__global__ void __launch_bounds__(256) read_once(const int4 *__restrict__ a, long long n16, int *__restrict__ out)
{
    const long long base = (long long)blockIdx.x * 256 * U + threadIdx.x;
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long i = base + (long long)u * 256;
        v[u] = i < n16 ? __ldcs(a + i) : make_int4(0, 0, 0, 0);
    }
    int s = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) s += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0 && s == 0x5a5a5a5a) out[blockIdx.x] = s;  // never true for zero-filled buffers
}

template <int U>
__global__ void __launch_bounds__(256) read_persist(const int4 *__restrict__ a, long long n16, int *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * 256 * U;
    long long base = (long long)blockIdx.x * 256 * U + threadIdx.x;
    int4 v[U], w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long i = base + (long long)u * 256;
        v[u] = i < n16 ? __ldcs(a + i) : make_int4(0, 0, 0, 0);
    }
    int s = 0;
    for (; base < n16; base += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + stride + (long long)u * 256;
            w[u] = i < n16 ? __ldcs(a + i) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) s += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = w[u];
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0 && s == 0x5a5a5a5a) out[blockIdx.x] = s;
}

// ---- the same stream with the gather of an SpMV (--gather): what the x gather costs by itself -------------
// idx[j] = a column within +-half_band of row j / npr (the banded bench matrix in spirit), val[j] = 1.  A thread
// loads U groups of four (index, value) pairs with 128-bit loads, gathers x through the read-only path (or not:
// GATHER = false multiplies by a constant instead), and adds everything up; one 4-byte store per warp.
__global__ void fill_banded(int *idx, float *val, long long n, int npr, int half_band, int n_cols)
{
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const long long row = j / npr;
        unsigned h = (unsigned)(j * 2654435761u) ^ (unsigned)(j >> 17);
        h ^= h >> 13;
        h *= 0x5bd1e995u;
        h ^= h >> 15;
        long long c = row - half_band + (long long)(h % (2u * half_band + 1u));
        c = c < 0 ? c + n_cols : (c >= n_cols ? c - n_cols : c);
        idx[j] = (int)c;
        val[j] = 1.0f;
    }
}

template <int U, bool GATHER>
__global__ void __launch_bounds__(256) spmv_like(const int4 *__restrict__ idx, const float4 *__restrict__ val,
                                                 const float *__restrict__ x, long long n4, float *__restrict__ out)
{
    const long long base = (long long)blockIdx.x * 256 * U + threadIdx.x;
    int4 c[U];
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long i = base + (long long)u * 256;
        c[u] = i < n4 ? __ldcs(idx + i) : make_int4(0, 0, 0, 0);
        v[u] = i < n4 ? __ldcs(val + i) : make_float4(0, 0, 0, 0);
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (GATHER)
            acc += v[u].x * __ldg(x + c[u].x) + v[u].y * __ldg(x + c[u].y) + v[u].z * __ldg(x + c[u].z) + v[u].w * __ldg(x + c[u].w);
        else
            acc += v[u].x * (float)c[u].x + v[u].y * (float)c[u].y + v[u].z * (float)c[u].z + v[u].w * (float)c[u].w;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) out[((long long)blockIdx.x * 256 + threadIdx.x) >> 5] = acc;
}

static int gather_probe()
{
    const long long rows = 2097152, npr = 64, n = rows * npr, n4 = n / 4;
    int *idx = nullptr;
    float *val = nullptr, *x = nullptr, *out = nullptr;
    CK(cudaMalloc(&idx, n * 4));
    CK(cudaMalloc(&val, n * 4));
    CK(cudaMalloc(&x, rows * 4));
    CK(cudaMalloc(&out, (n4 / 32 + 1024) * 4));
    CK(cudaMemset(x, 0, rows * 4));
    fill_banded<<<2368, 256>>>(idx, val, n, (int)npr, 2000, (int)rows);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("{\"gather_probe\": {\"entries\": %lld, \"bytes_streamed\": %lld, \"variants\": {", n, n * 8);
    bool first = true;
    for (int variant = 0; variant < 6; ++variant) {
        const int u = variant % 3 == 0 ? 1 : (variant % 3 == 1 ? 2 : 4);
        const bool gather = variant >= 3;
        const unsigned grid = (unsigned)((n4 + 256ll * u - 1) / (256ll * u));
        auto launch = [&]() {
            const int4 *ci = reinterpret_cast<const int4 *>(idx);
            const float4 *cv = reinterpret_cast<const float4 *>(val);
            if (gather) {
                if (u == 1) spmv_like<1, true><<<grid, 256>>>(ci, cv, x, n4, out);
                else if (u == 2) spmv_like<2, true><<<grid, 256>>>(ci, cv, x, n4, out);
                else spmv_like<4, true><<<grid, 256>>>(ci, cv, x, n4, out);
            } else {
                if (u == 1) spmv_like<1, false><<<grid, 256>>>(ci, cv, x, n4, out);
                else if (u == 2) spmv_like<2, false><<<grid, 256>>>(ci, cv, x, n4, out);
                else spmv_like<4, false><<<grid, 256>>>(ci, cv, x, n4, out);
            }
        };
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 20; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / 20;
        printf("%s\"%s_u%d\": {\"us\": %.2f, \"gbs\": %.0f}", first ? "" : ", ", gather ? "gather" : "no_gather", u, us,
               (n * 8.0 + rows * 4.0 + n / 32.0) / us * 1e-3);
        first = false;
    }
    printf("}}}\n");
    return 0;
}

int main(int argc, char **argv)
{
    for (int i = 1; i < argc; ++i)
        if (!strcmp(argv[i], "--gather")) return gather_probe();
    std::vector<double> mbs = {35.4, 53.2, 70.2, 280, 1090};
    int copies = 7, reps = 28;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--mb") && i + 1 < argc) {
            mbs.clear();
            char *tok = strtok(argv[++i], ",");
            while (tok) {
                mbs.push_back(atof(tok));
                tok = strtok(nullptr, ",");
            }
        } else if (!strcmp(argv[i], "--copies") && i + 1 < argc) {
            copies = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--reps") && i + 1 < argc) {
            reps = atoi(argv[++i]);
        }
    }
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    int *out = nullptr;
    CK(cudaMalloc(&out, sizeof(int) << 22));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("{\"sms\": %d, \"copies\": %d, \"launches_per_graph\": %d, \"sizes\": [", sms, copies, reps);
    bool first_size = true;
    for (double mb : mbs) {
        const long long n16 = (long long)(mb * 1e6 / 16);
        std::vector<int4 *> buf(copies, nullptr);
        for (int c = 0; c < copies; ++c) {
            CK(cudaMalloc(&buf[c], (size_t)n16 * 16));
            CK(cudaMemsetAsync(buf[c], 0, (size_t)n16 * 16, st));
        }
        struct Variant {
            std::string name;
            int kind, u, per_sm;
        };
        std::vector<Variant> vs = {{"read_once_u2", 0, 2, 0},     {"read_once_u4", 0, 4, 0},     {"read_once_u8", 0, 8, 0},
                                   {"read_persist_u4_2perSM", 1, 4, 2}, {"read_persist_u4_4perSM", 1, 4, 4},
                                   {"read_persist_u8_2perSM", 1, 8, 2}, {"read_persist_u4_8perSM", 1, 4, 8}};
        printf("%s{\"mb\": %.1f, \"variants\": {", first_size ? "" : ", ", mb);
        first_size = false;
        bool first_v = true;
        for (const Variant &v : vs) {
            cudaGraph_t g;
            cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int r = 0; r < reps; ++r) {
                const int4 *a = buf[r % copies];
                if (v.kind == 0) {
                    const unsigned grid = (unsigned)((n16 + 256ll * v.u - 1) / (256ll * v.u));
                    if (v.u == 2) read_once<2><<<grid, 256, 0, st>>>(a, n16, out);
                    else if (v.u == 4) read_once<4><<<grid, 256, 0, st>>>(a, n16, out);
                    else read_once<8><<<grid, 256, 0, st>>>(a, n16, out);
                } else {
                    const unsigned grid = (unsigned)(sms * v.per_sm);
                    if (v.u == 4) read_persist<4><<<grid, 256, 0, st>>>(a, n16, out);
                    else read_persist<8><<<grid, 256, 0, st>>>(a, n16, out);
                }
            }
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaStreamSynchronize(st));
            float best = 1e30f;
            for (int t = 0; t < 5; ++t) {
                CK(cudaEventRecord(e0, st));
                CK(cudaGraphLaunch(ge, st));
                CK(cudaEventRecord(e1, st));
                CK(cudaStreamSynchronize(st));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                best = std::min(best, ms);
            }
            const double us = best * 1e3 / reps;
            printf("%s\"%s\": {\"us\": %.2f, \"gbs\": %.0f}", first_v ? "" : ", ", v.name.c_str(), us, n16 * 16.0 / us * 1e-3);
            first_v = false;
            CK(cudaGraphExecDestroy(ge));
            CK(cudaGraphDestroy(g));
        }
        printf("}}");
        for (int c = 0; c < copies; ++c) CK(cudaFree(buf[c]));
    }
    printf("]}\n");
    return 0;
}
