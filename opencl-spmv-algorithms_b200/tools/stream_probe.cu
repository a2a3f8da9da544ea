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

int main(int argc, char **argv)
{
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
