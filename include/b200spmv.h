/*
 * b200spmv.h -- C ABI of libb200spmv.so, the B200 (sm_100a) SpMV library.
 *
 * This header is the drop-in boundary.  The reference (sgartkink/opencl-spmv-algorithms) has no
 * plugin API: its seam is the set of OpenCL runtime calls inlined in each driver's main() plus
 * three helpers in inc/helper_functions.h.  Every entry point below names the reference call(s)
 * it replaces (file:line relative to the reference root).  Conventions kept from the reference:
 *
 *   - every call returns an int status, 0 == success (cl_int / CL_SUCCESS, e.g. csr.c:117);
 *   - the caller owns host memory; the library owns device memory behind plain device pointers
 *     that the caller releases explicitly (csr.c:260-263);
 *   - one context = one device + one in-order stream (clCreateContext +
 *     clCreateCommandQueueWithProperties(props = 0), csr.c:107,115); uploads are asynchronous and
 *     completed by b200_sync (clEnqueueWriteBuffer(CL_FALSE) + clFinish, csr.c:183-193); downloads
 *     block (clEnqueueReadBuffer(CL_TRUE), csr.c:220);
 *   - kernels are compiled ahead of time into the library: there is no runtime source read
 *     (read_source_from_cl_file / clCreateProgramWithSource / clBuildProgram / clCreateKernel,
 *     csr.c:136-159, have no counterpart and the kernels/ directory dependency disappears).
 *
 * Signatures use only plain pointers and sizes.  All "const T *" array arguments of the launch,
 * build and generator entry points are DEVICE pointers obtained from b200_malloc (or any CUDA
 * allocation, e.g. a torch tensor's data_ptr()).  There is no CPU fallback anywhere: without a
 * CUDA device every entry point that needs one fails with B200_ERR_NO_DEVICE / B200_ERR_CUDA.
 */
#ifndef B200SPMV_H
#define B200SPMV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SPMV_VERSION 100 /* 1.0.0 */

/* status codes; the drivers map them to the reference's ReturnCode (inc/enums.h:4-11):
 * B200_ERR_NO_DEVICE -> OpenCLDeviceError(1), every other failure -> OpenCLProgramError(2). */
enum {
    B200_SUCCESS = 0,
    B200_ERR_NO_DEVICE = 1,
    B200_ERR_CUDA = 2,
    B200_ERR_INVALID_VALUE = 3,
    B200_ERR_OUT_OF_MEMORY = 4,
    B200_ERR_UNSUPPORTED = 5,
    B200_ERR_DOMAIN = 6, /* input outside the builder's domain (e.g. rows not sorted) */
    B200_ERR_COMM = 7    /* NCCL reported an error (synchronously, or asynchronously: b200_comm_check) */
};

/* the reference's five formats, in the order of its README */
enum { B200_FORMAT_COO = 0, B200_FORMAT_CSR = 1, B200_FORMAT_ELL = 2, B200_FORMAT_SELL = 3, B200_FORMAT_CMRS = 4 };

typedef struct b200_ctx b200_ctx;     /* cl_context + cl_command_queue */
typedef struct b200_event b200_event; /* cl_event, but with device timestamps */

const char *b200_status_string(int status);
const char *b200_last_error(void); /* detail of the last failure on this thread */
int b200_version(void);

/* ---- devices: replaces get_device_ids (inc/helper_functions.h:76-129) ---- */
int b200_get_device_count(int *count);
int b200_device_name(int device, char *buf, size_t buf_len);
int b200_device_sm_count(int device, int *sm_count);

/* ---- context / queue: clCreateContext + clCreateCommandQueueWithProperties (csr.c:107-121),
 *      clFlush/clReleaseCommandQueue/clReleaseContext (csr.c:273-277) ---- */
int b200_ctx_create(int device, b200_ctx **ctx);
/* same, but enqueue on an existing CUDA stream (a cudaStream_t passed as void*, 0 = default) */
int b200_ctx_create_on_stream(int device, void *cuda_stream, b200_ctx **ctx);
int b200_ctx_destroy(b200_ctx *ctx);
int b200_ctx_device(const b200_ctx *ctx, int *device);
/* keep `bytes` at `dptr` (normally x) resident in L2 for later launches (access-policy window on this
 * context's queue + a persisting-L2 set-aside, which is DEVICE-wide: up to 79 of the 126 MB); bytes = 0
 * clears the window and gives the set-aside back (so does b200_ctx_destroy) -- a stale one slows every
 * other kernel on the device.  New: the OpenCL reference has no equivalent. */
int b200_ctx_set_l2_persist(b200_ctx *ctx, const void *dptr, size_t bytes);

/* Launch overlap (new; programmatic dependent launch).  With overlap on, an SpMV launch (CSR / ELL /
 * SELL / CMRS / COO, launches of at most a few waves) that DIRECTLY follows another SpMV launch on this
 * context may start before its predecessor has finished: it streams its row pointers and first batch of
 * indices / values at once and reads `vect` / writes `output` only after everything queued before it has
 * completed.  Results are unchanged; the ~2 us of launch ramp and drain between consecutive small
 * launches (a cant-sized SpMV lasts ~12 us) overlap.  It is safe by construction: any other call that
 * enters the context -- an upload, a build, a pack, a memset, an event, a graph launch -- makes the next
 * SpMV launch fully ordered again, so matrix arrays are never read before they are written.
 * Default: ON for contexts that own their queue (b200_ctx_create: only library calls enqueue there),
 * OFF for b200_ctx_create_on_stream (foreign kernels on that stream could write a matrix array between
 * two SpMV launches without the library knowing); enable = 0 gives the strict in-order queue of the
 * reference (csr.c:115). */
int b200_ctx_set_launch_overlap(b200_ctx *ctx, int enable);

/* Tuning hooks (new).  Every kernel-selection heuristic can be overridden per context: `name` is one
 * of the B200_* hooks listed in DESIGN.md section 4 ("B200_CSR_LANES", "B200_SELL_WPC", ...), `value`
 * its decimal value, NULL or "" = back to automatic.  The environment variables of the same names are
 * read exactly once, by b200_ctx_create; launches never call getenv.  Unknown name ->
 * B200_ERR_INVALID_VALUE. */
int b200_ctx_set_option(b200_ctx *ctx, const char *name, const char *value);

/* ---- buffers: clCreateBuffer (csr.c:123-127), clReleaseMemObject (csr.c:260-263),
 *      clEnqueueWriteBuffer (csr.c:183-186), clEnqueueReadBuffer (csr.c:220), clFinish ---- */
/* b200_malloc over-allocates by B200_MALLOC_PAD bytes: the kernels fetch indices and values as
 * 16-byte-ALIGNED groups of four entries, so the last group of an array can reach up to 24 bytes
 * past its end (fp64 values, nnz % 4 == 1).  Matrix arrays that do not come from b200_malloc (a torch
 * tensor, a cudaMalloc of the exact size) must carry B200_MALLOC_PAD readable bytes after their last
 * entry as well; vectors (vect / output) need none. */
#define B200_MALLOC_PAD 32
int b200_malloc(b200_ctx *ctx, size_t bytes, void **dptr);
int b200_free(b200_ctx *ctx, void *dptr);
int b200_memcpy_h2d_async(b200_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);
int b200_memcpy_d2h(b200_ctx *ctx, void *dst_host, const void *src_device, size_t bytes);
int b200_memcpy_d2h_async(b200_ctx *ctx, void *dst_host, const void *src_device, size_t bytes);
int b200_memcpy_d2d_async(b200_ctx *ctx, void *dst_device, const void *src_device, size_t bytes);
int b200_memset_async(b200_ctx *ctx, void *dst_device, int byte_value, size_t bytes);
int b200_sync(b200_ctx *ctx);
int b200_host_alloc_pinned(size_t bytes, void **hptr);
int b200_host_free_pinned(void *hptr);

/* ---- timing: the reference brackets the launch with clock_gettime + clWaitForEvents
 *      (csr.c:198-206); these are device timestamps on the context's stream ---- */
int b200_event_create(b200_ctx *ctx, b200_event **ev);
int b200_event_record(b200_ctx *ctx, b200_event *ev);
int b200_event_elapsed_ms(b200_event *start, b200_event *stop, float *ms); /* waits for `stop` */
int b200_event_destroy(b200_event *ev);

/* ---- launch graphs (new).  The reference issues ONE clEnqueueNDRangeKernel per program run
 *      (csr.c:201) and never loops; callers that do loop -- the iterated mode, a sweep over several
 *      matrices -- pay more host time per launch than a cant-sized SpMV lasts on the device (~10 us).
 *      Every launch / async copy / memset issued on `ctx` between begin and end is recorded into a
 *      CUDA graph instead of being executed (the nearest OpenCL notion is cl_khr_command_buffer);
 *      b200_graph_launch replays the whole sequence, in order, with a single call.  Calls that
 *      synchronise or allocate (SpMV entry points with plan == NULL, *_plan_create, b200_malloc,
 *      b200_memcpy_d2h, b200_sync) are not allowed while recording and fail with B200_ERR_CUDA. ---- */
typedef struct b200_graph b200_graph;
int b200_graph_begin(b200_ctx *ctx);
int b200_graph_end(b200_ctx *ctx, b200_graph **graph);
int b200_graph_launch(b200_ctx *ctx, b200_graph *graph);
int b200_graph_destroy(b200_graph *graph);

/* =====================================================================================
 * SpMV launches.  One entry per format x dtype; argument order = the order the reference
 * driver sets with clSetKernelArg, followed by what the OpenCL kernel took from its launch
 * shape or left undefined.  All launches are asynchronous on the context's stream.
 * ===================================================================================== */

/* CSR: kernel csr(ptr,col,data,vect,output,N) kernels/Csr.cl:1, args csr.c:170-175.
 * `plan` carries the row-length statistics that pick the lanes-per-row variant and the
 * long-row list; NULL = analyse `ptr` on the fly (one extra pass + a small sync). */
typedef struct b200_csr_plan b200_csr_plan;
typedef struct {
    int n_rows;
    long long nnz;
    int min_len, max_len;
    double mean_len;
    int lanes_per_row; /* 2,4,8,16,32: the vector-kernel variant chosen */
    int long_threshold; /* rows longer than this go to the block-per-row kernel */
    int n_long_rows;
    int stream_tiles; /* > 0: short rows everywhere -> the nnz-split stream kernel is used instead,
                         with this many tiles (a plan then owns carry buffers: it serves the context
                         it was created on) */
    int stream_tile_entries; /* entries per tile: 1024 x groups per thread (1, 2 or 4) */
} b200_csr_plan_info;
int b200_csr_plan_create(b200_ctx *ctx, const int *ptr, int n_rows, b200_csr_plan **plan);
int b200_csr_plan_get_info(const b200_csr_plan *plan, b200_csr_plan_info *info);
int b200_csr_plan_destroy(b200_csr_plan *plan);
int b200_spmv_csr_f64(b200_ctx *ctx, const int *ptr, const int *col, const double *data,
                      const double *vect, double *output, int n_rows, const b200_csr_plan *plan);
int b200_spmv_csr_f32(b200_ctx *ctx, const int *ptr, const int *col, const float *data,
                      const float *vect, float *output, int n_rows, const b200_csr_plan *plan);

/* COO: kernel coo(row,col,data,vect,output,N=nnz) kernels/Coo.cl:24, args coo.c:163-168.
 * Entries may be in any order.  `output` (n_rows elements) is zero-filled by the call (the
 * reference never zeroes it: coo.c:120, quirk q1), then accumulated with a segmented warp
 * reduction and one atomic per (warp, row-run). */
int b200_spmv_coo_f64(b200_ctx *ctx, const int *row, const int *col, const double *data,
                      const double *vect, double *output, int nnz, int n_rows);
int b200_spmv_coo_f32(b200_ctx *ctx, const int *row, const int *col, const float *data,
                      const float *vect, float *output, int nnz, int n_rows);

/* ELL, the reference's ROW-MAJOR arrays: kernel ell(data,indices,vect,output,N,row_size,local)
 * kernels/Ell.cl:1, args ell.c:242-248 (the __local scratch argument disappears). */
int b200_spmv_ell_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                      double *output, int n_rows, int row_size);
int b200_spmv_ell_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                      float *output, int n_rows, int row_size);
/* ELL, COLUMN-MAJOR device layout (element k of row r at k*pitch + r, pitch % 32 == 0,
 * pitch >= n_rows): the transpose of the same K-padded matrix, built by b200_build_ell_colmajor
 * or b200_ell_transpose.  This is the coalesced thread-per-row kernel. */
int b200_spmv_ellcm_f64(b200_ctx *ctx, const double *data_cm, const int *indices_cm,
                        const double *vect, double *output, int n_rows, int row_size, int pitch);
int b200_spmv_ellcm_f32(b200_ctx *ctx, const float *data_cm, const int *indices_cm,
                        const float *vect, float *output, int n_rows, int row_size, int pitch);

/* SELL-C (C = 32): kernel sigma_c(data,indices,vect,output,row_indices,C) kernels/Sigma_C.cl:1,
 * args sigma_c.c:280-285.  The reference launches one 32-lane group per slice and writes
 * n_slices*32 outputs including the padding rows of the last slice (Sigma_C.cl:17,
 * sigma_c.c:212); `n_out` says how many leading outputs to write (n_slices*32 for the
 * reference's padded buffer, n_rows otherwise).  `perm` (new; NULL = identity) is the
 * sigma-window permutation perm[new_row] = old_row: row new_row's result goes to
 * output[perm[new_row]].  chunk must be 32 (warp-aligned chunks).
 * `plan` (new; NULL = one warp walks each chunk whole, which is what FEM-like inputs want) lists
 * the chunks wider than 256 columns -- power-law inputs, where one hub row makes a chunk 10^5
 * columns wide -- so that their columns are split over many warps (atomics for the extra pieces).
 * The plan also counts the chunks wider than 8 columns: a matrix with none (a stencil) in a launch of
 * at least 4 chunks per resident warp is served by the persistent software-pipelined kernel
 * (B200_SELL_PIPE = 0 turns that off, 2 | 3 | 4 force it at that many blocks per SM). */
typedef struct b200_sell_plan b200_sell_plan;
int b200_sell_plan_create(b200_ctx *ctx, const int *row_indices, int n_slices, b200_sell_plan **plan);
int b200_sell64_plan_create(b200_ctx *ctx, const long long *slice_ptr, int n_slices, b200_sell_plan **plan);
int b200_sell_plan_extra_items(const b200_sell_plan *plan, int *n_items);
int b200_sell_plan_destroy(b200_sell_plan *plan);
int b200_spmv_sell_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                       double *output, const int *row_indices, int chunk, int n_slices, int n_out,
                       const int *perm, const b200_sell_plan *plan);
int b200_spmv_sell_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                       float *output, const int *row_indices, int chunk, int n_slices, int n_out,
                       const int *perm, const b200_sell_plan *plan);
/* same with 64-bit chunk pointers (the reference's cl_int row_indices overflows when the padded
 * size exceeds 2^31, sigma_c.c:41,118-119) */
int b200_spmv_sell64_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                         double *output, const long long *slice_ptr, int chunk, int n_slices,
                         int n_out, const int *perm, const b200_sell_plan *plan);
int b200_spmv_sell64_f32(b200_ctx *ctx, const float *data, const int *indices, const float *vect,
                         float *output, const long long *slice_ptr, int chunk, int n_slices,
                         int n_out, const int *perm, const b200_sell_plan *plan);

/* CMRS: kernel cmrs(data,indices,strip_ptr,row_in_strip,vect,output,N=strips,height,local)
 * kernels/Cmrs.cl:1, args cmrs.c:197-205.  `n_rows` bounds the store of the last strip (the
 * reference writes height outputs per strip unconditionally, Cmrs.cl:38-41, quirk q5).
 * height <= 32. */
typedef struct b200_cmrs_plan b200_cmrs_plan; /* strips longer than 8192 entries (power-law hubs) are
                                               * split over several warps; NULL = never split.  When the
                                               * strip lengths are skewed (longest > 32 x mean) the plan
                                               * also prepares the nnz-split kernel: warps own 512-entry
                                               * tiles instead of strips, rows accumulate like COO rows
                                               * (atomics at run ends); *n_tiles > 0 says so */
int b200_cmrs_plan_create(b200_ctx *ctx, const int *strip_ptr, int n_strips, b200_cmrs_plan **plan);
int b200_cmrs_plan_extra_items(const b200_cmrs_plan *plan, int *n_items);
int b200_cmrs_plan_stream_tiles(const b200_cmrs_plan *plan, int *n_tiles);
int b200_cmrs_plan_destroy(b200_cmrs_plan *plan);
int b200_spmv_cmrs_f64(b200_ctx *ctx, const double *data, const int *indices, const int *strip_ptr,
                       const int *row_in_strip, const double *vect, double *output, int n_strips,
                       int height, int n_rows, const b200_cmrs_plan *plan);
int b200_spmv_cmrs_f32(b200_ctx *ctx, const float *data, const int *indices, const int *strip_ptr,
                       const int *row_in_strip, const float *vect, float *output, int n_strips,
                       int height, int n_rows, const b200_cmrs_plan *plan);

/* CMRS, packed device layout (new): packed[i] = (row_in_strip[i] << 27) | indices[i], so
 * the kernel streams 4 + V bytes per entry instead of the 8 + V of the reference's two index arrays
 * (cmrs.c:42-45).  The reference arrays stay the bit-exact build product; this is a derived layout
 * like column-major ELL.  Needs height <= 32 and n_cols <= 2^27 (else B200_ERR_UNSUPPORTED). */
int b200_cmrs_pack(b200_ctx *ctx, const int *indices, const int *row_in_strip, long long nnz, int n_cols,
                   int height, int *packed);
int b200_spmv_cmrs_packed_f64(b200_ctx *ctx, const double *data, const int *packed, const int *strip_ptr,
                              const double *vect, double *output, int n_strips, int height, int n_rows,
                              const b200_cmrs_plan *plan);
int b200_spmv_cmrs_packed_f32(b200_ctx *ctx, const float *data, const int *packed, const int *strip_ptr,
                              const float *vect, float *output, int n_strips, int height, int n_rows,
                              const b200_cmrs_plan *plan);

/* =====================================================================================
 * Format builds on the GPU (new; in the reference they are host loops inlined in each main()).
 * Input: device COO triples as the drivers parse them ("%d %d %lg", 1-based -> 0-based,
 * coo.c:79-84), SORTED BY ROW for everything except COO itself.  On the reference's
 * well-defined domain (rows sorted, none empty, first row 0) every integer array is identical
 * to the reference's; empty rows are handled correctly instead of shifting (quirk q4).
 * ===================================================================================== */

/* checks sortedness; B200_ERR_DOMAIN if rows are not non-decreasing or out of range */
int b200_check_sorted_rows(b200_ctx *ctx, const int *rows, int nnz, int n_rows);
/* CSR row pointer (csr.c:72-91): ptr has n_rows+1 entries */
int b200_build_csr_ptr(b200_ctx *ctx, const int *rows, int nnz, int n_rows, int *ptr);
/* row-length statistics from ptr (host outputs).  *_excl_last reproduce what ell.c:68-104 prints
 * (the reference never counts the last row) */
typedef struct {
    int max_len, min_len;
    long long sum_len;
    int max_len_excl_last, min_len_excl_last;
    long long sum_len_excl_last;
    int last_len;
} b200_row_stats;
int b200_row_length_stats(b200_ctx *ctx, const int *ptr, int n_rows, b200_row_stats *stats);
/* ELL (ell.c:118-164): row-major, padding = (col 0, value 0) */
int b200_build_ell_f64(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                       int n_rows, int row_size, int *ell_cols, double *ell_data);
int b200_build_ell_f32(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                       int n_rows, int row_size, int *ell_cols, float *ell_data);
int b200_build_ell_colmajor_f64(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                                int n_rows, int row_size, int pitch, int *cm_cols,
                                double *cm_data);
int b200_build_ell_colmajor_f32(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                                int n_rows, int row_size, int pitch, int *cm_cols, float *cm_data);
/* SELL-C-sigma (sigma_c.c:71-202 at sigma <= 1).  Two steps so the caller can allocate:
 *  1. b200_build_sell_ptr: perm (n_rows ints; identity for sigma <= 1; may be NULL when
 *     sigma <= 1), slice_ptr (n_slices+1 int64) and the padded size *total (host output);
 *  2. b200_build_sell_fill_*: column-major-in-slice fill, padding = (col 0, value 0).
 * b200_sell_ptr_to_i32 narrows slice_ptr to the reference's cl_int row_indices
 * (B200_ERR_DOMAIN on overflow). */
int b200_sell_num_slices(int n_rows, int chunk);
int b200_build_sell_ptr(b200_ctx *ctx, const int *ptr, int n_rows, int chunk, int sigma, int *perm,
                        long long *slice_ptr, long long *total);
int b200_sell_ptr_to_i32(b200_ctx *ctx, const long long *slice_ptr, int n_slices,
                         int *row_indices);
int b200_build_sell_fill_f64(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                             int n_rows, int chunk, const int *perm, const long long *slice_ptr,
                             int *sell_cols, double *sell_data);
int b200_build_sell_fill_f32(b200_ctx *ctx, const int *ptr, const int *cols, const double *vals,
                             int n_rows, int chunk, const int *perm, const long long *slice_ptr,
                             int *sell_cols, float *sell_data);
/* CMRS (cmrs.c:72-117): strip_ptr has ceil(n_rows/height)+1 entries, row_in_strip has nnz */
int b200_cmrs_num_strips(int n_rows, int height);
int b200_build_cmrs(b200_ctx *ctx, const int *rows, const int *ptr, int nnz, int n_rows,
                    int height, int *strip_ptr, int *row_in_strip);
/* value conversion for the fp32 path (the reference is fp64-only) */
int b200_convert_f64_to_f32(b200_ctx *ctx, const double *src, float *dst, long long n);
/* a[i] += delta: rebases the row indices (or a row pointer) of a row block to shard-local */
int b200_offset_i32(b200_ctx *ctx, int *a, long long n, int delta);
/* x[i] = i, the reference's input vector (csr.c:95-99) */
int b200_fill_ramp_f64(b200_ctx *ctx, double *x, int n);
int b200_fill_ramp_f32(b200_ctx *ctx, float *x, int n);

/* =====================================================================================
 * Synthetic matrices generated on the device (BASELINE.json configs 3-5; the reference can only
 * ingest MatrixMarket text).  Each fills rows [row_begin, row_begin+row_count) of the global
 * matrix as row-sorted COO triples with global column indices; *_nnz gives the exact count.
 * ===================================================================================== */
long long b200_gen_banded_nnz(long long n_global, int row_begin, int row_count, int nnz_per_row);
int b200_gen_banded_coo(b200_ctx *ctx, int n_global, int row_begin, int row_count, int nnz_per_row,
                        int half_band, uint64_t seed, int *rows, int *cols, double *vals);
long long b200_gen_laplace7_nnz(int nx, int ny, int nz, int row_begin, int row_count);
int b200_gen_laplace7_coo(b200_ctx *ctx, int nx, int ny, int nz, int row_begin, int row_count,
                          int *rows, int *cols, double *vals);
/* R-MAT power-law matrix, n = 2^scale: n*edge_factor edges with quadrant probabilities
 * (a, b, c, 1-a-b-c), one diagonal entry per row added (no empty rows), duplicates removed.
 * b200_gen_rmat_count gives an upper bound (*n_candidates, before duplicate removal) for the
 * output capacity of the row block; b200_gen_rmat_coo writes *nnz_out entries. */
int b200_gen_rmat_count(b200_ctx *ctx, int scale, int edge_factor, double a, double b, double c,
                        uint64_t seed, int row_begin, int row_count, long long *n_candidates);
int b200_gen_rmat_coo(b200_ctx *ctx, int scale, int edge_factor, double a, double b, double c,
                      uint64_t seed, int row_begin, int row_count, long long capacity, int *rows,
                      int *cols, double *vals, long long *nnz_out);
int b200_gen_uniform_f64(b200_ctx *ctx, double *x, long long n, uint64_t seed, double lo, double hi);
int b200_gen_uniform_f32(b200_ctx *ctx, float *x, long long n, uint64_t seed, float lo, float hi);
/* host twins of the generators' per-element functions (same integer hash; bit-identical), so a
 * CPU-only process can rebuild any row block of the same matrix */
int b200_gen_banded_coo_host(int n_global, int row_begin, int row_count, int nnz_per_row,
                             int half_band, uint64_t seed, int *rows, int *cols, double *vals);
int b200_gen_uniform_f64_host(double *x, long long n, uint64_t seed, double lo, double hi);

/* =====================================================================================
 * Multi-GPU row partition (new): contiguous, nnz-balanced row blocks whose cut points are
 * multiples of `align` (lcm of the SELL chunk, CMRS height and sigma window) so that every
 * shard's format arrays are slices of the global build.  Host function over a host ptr[].
 * cuts has n_parts+1 entries: part p owns rows [cuts[p], cuts[p+1]).
 * ===================================================================================== */
int b200_partition_rows(const int *ptr_host, int n_rows, int n_parts, int align, int *cuts);
/* Fused SpMV + exchange for the iterated mode (new; replaces SpMV -> scale -> all-gather):
 *   y = (A_r . x) / sqrt(sum(scale_sumsq[0..32)))     (scale_sumsq: device, NULL = no scaling)
 * is stored at [dst_offset, dst_offset + n_rows) of EVERY buffer in dst[0..n_dst) -- dst is a HOST
 * array of device pointers: this rank's own next-x buffer plus the peers' buffers mapped with
 * b200_ipc_open_handle, so the stores travel over NVLink/NVSwitch while the SpMV is still running.
 * The same kernel accumulates ||y||^2 of its rows into the B200_SUMSQ_SLOTS partial sums of
 * sumsq_out (device, zeroed by the caller; NULL = skip), one atomic per block.  SELL-32, reference
 * chunk pointers, no permutation.  The caller orders steps with one small all-reduce of those
 * slots (which it needs anyway for the norm): no other launch is needed per step.
 * A warp stores its chunk's 32 rows as one contiguous 256-byte write per destination; dst[i] only
 * needs the natural 8-byte alignment of a double array (checked: B200_ERR_INVALID_VALUE).
 * Blocks of at least 4 chunks per resident warp run as a persistent grid whose warps keep the next
 * chunk's loads in flight while the current one is gathered and stored (B200_BCAST_U = 1: one chunk
 * per warp always; 2 | 3 | 4: the persistent kernel at that many blocks per SM); ||y||^2 then costs one
 * atomic per resident block.  Same y bit for bit either way. */
#define B200_SUMSQ_SLOTS 32
int b200_spmv_sell_bcast_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                             const int *row_indices, int chunk, int n_slices, int n_rows,
                             const double *scale_sumsq, double *sumsq_out, double *const *dst,
                             int n_dst, long long dst_offset);
/* Same kernel, halo-limited: destination d receives only rows [dst_row_lo[d], dst_row_hi[d]) of this
 * rank's block (local row numbers; HOST arrays of n_dst ints; lo >= hi = nothing).  The caller passes,
 * per destination, the rows that destination's matrix block actually reads as columns
 * (b200_minmax_i32 over its column indices, exchanged once): the rank's own buffer gets the whole
 * block, a neighbour the halo, everyone else nothing.  NULL, NULL = every row to every destination
 * (b200_spmv_sell_bcast_f64).  SURVEY 8f.3. */
int b200_spmv_sell_halo_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                            const int *row_indices, int chunk, int n_slices, int n_rows,
                            const double *scale_sumsq, double *sumsq_out, double *const *dst, int n_dst,
                            long long dst_offset, const int *dst_row_lo, const int *dst_row_hi);
/* Same kernel with NO collective call per step at all ("ring" of flags over peer memory).  Every rank
 * allocates one zero-filled sync block of B200_SYNC_BLOCK_BYTES with b200_malloc and maps the others'
 * (b200_ipc_*); sync_blocks[r] (HOST array, n_dst entries, indexed by RANK like dst[]) is rank r's
 * block as seen from this process.  Step k is two launches on the context's queue: the fused SpMV +
 * halo-store kernel (1/||x|| from the sums folded at the end of step k-1; k = 0: no scaling), and a
 * one-warp kernel that publishes this rank's partial sums of ||y||^2 to every rank, fences system-wide,
 * releases "step k done" flags in every rank's block, waits until every rank's flag has arrived here
 * and folds the ranks' sums for step k+1.  The flags are the barrier that orders the peer writes of the
 * double-buffered x: work queued after this call sees every rank's step-k rows.  All ranks must call
 * it with the same consecutive step numbers, concurrently (a rank that never arrives is reported as
 * B200_ERR_CUDA by the next b200_sync after a 2 s wait, not a hang).  ||y||^2 of step k = sum over
 * ranks r and slots s of the doubles at byte offset 8*(16 + ((k&1)*16 + r)*32 + s) of any rank's block. */
#define B200_SYNC_BLOCK_BYTES 16384
int b200_spmv_sell_ring_f64(b200_ctx *ctx, const double *data, const int *indices, const double *vect,
                            const int *row_indices, int chunk, int n_slices, int n_rows, double *const *dst,
                            int n_dst, long long dst_offset, const int *dst_row_lo, const int *dst_row_hi,
                            void *const *sync_blocks, int my_rank, unsigned long long step);
/* min and max of a device int array (host outputs): the column range a row block reads */
int b200_minmax_i32(b200_ctx *ctx, const int *a, long long n, int *min_out, int *max_out);
/* which blocks of 2^block_log2 consecutive columns a row block reads (host output, one byte per column
 * block, ceil(n_cols / 2^block_log2) bytes: 1 = some entry has its column there).  What a caller that
 * keeps x on the host has to upload before an SpMV of this block -- a banded shard reads a window of x
 * (two windows when the band wraps around), not the whole vector. */
int b200_used_column_blocks(b200_ctx *ctx, const int *cols, long long nnz, int n_cols, int block_log2,
                            unsigned char *used_host);
/* CUDA IPC plumbing for the peers' buffers (allocations made with b200_malloc) */
int b200_ipc_get_handle(b200_ctx *ctx, void *dptr, unsigned char handle[64]);
int b200_ipc_open_handle(b200_ctx *ctx, const unsigned char handle[64], void **peer_dptr);
int b200_ipc_close_handle(b200_ctx *ctx, void *peer_dptr);
/* power-iteration helpers for the iterated mode: y *= scale; sum of squares into *acc (device) */
int b200_scale_f64(b200_ctx *ctx, double *y, long long n, const double *scale_device, int invert_sqrt);
int b200_sumsq_f64(b200_ctx *ctx, const double *y, long long n, double *acc_device);


/* =====================================================================================
 * Fewer bytes, format advice, conversions (new).  Derived device layouts and helpers: the
 * reference's arrays (sigma_c.c:40-43 etc.) stay the bit-exact build product.
 * ===================================================================================== */

/* SELL-32 with 16-bit column deltas ("sell16").  indices[j] = (chunk_base[s] + delta16[j]) mod n_cols
 * for the entries j of chunk s; chunk_base[s] = the chunk's smallest column, or, when the chunk's band
 * wraps around the matrix edge (periodic stencils), its smallest column in the upper half; padding
 * slots ((column 0, value 0) in the format itself) get delta 0.  2 + V bytes per entry instead of 4 + V.
 * Needs every chunk's columns to span at most 65536 modulo n_cols (banded / FEM matrices):
 * B200_ERR_UNSUPPORTED otherwise.  sigma = 1 layout with int32 chunk pointers; delta16 has as many
 * entries as indices. */
int b200_sell_pack16_f64(b200_ctx *ctx, const double *data, const int *indices, const int *row_indices, int n_slices,
                         int n_cols, int *chunk_base, unsigned short *delta16);
int b200_sell_pack16_f32(b200_ctx *ctx, const float *data, const int *indices, const int *row_indices, int n_slices,
                         int n_cols, int *chunk_base, unsigned short *delta16);
int b200_spmv_sell16_f64(b200_ctx *ctx, const double *data, const unsigned short *delta16, const int *chunk_base,
                         const double *vect, double *output, const int *row_indices, int chunk, int n_slices,
                         int n_out, int n_cols);
int b200_spmv_sell16_f32(b200_ctx *ctx, const float *data, const unsigned short *delta16, const int *chunk_base,
                         const float *vect, float *output, const int *row_indices, int chunk, int n_slices,
                         int n_out, int n_cols);

/* Back to CSR on the device, so that any format converts to any other through CSR and the builders
 * above.  Padding slots ((column 0, value 0)) are dropped.  ptr has n_rows + 1 entries; cols / vals may
 * be NULL to get the row pointer (and *nnz) first and allocate; nnz may be NULL (no sync then). */
int b200_ell_to_csr_f64(b200_ctx *ctx, const double *data, const int *indices, int n_rows, int row_size, int *ptr,
                        int *cols, double *vals, long long *nnz);
int b200_ell_to_csr_f32(b200_ctx *ctx, const float *data, const int *indices, int n_rows, int row_size, int *ptr,
                        int *cols, float *vals, long long *nnz);
int b200_sell_to_csr_f64(b200_ctx *ctx, const double *data, const int *indices, const int *row_indices, int n_rows,
                         int *ptr, int *cols, double *vals, long long *nnz);
int b200_sell_to_csr_f32(b200_ctx *ctx, const float *data, const int *indices, const int *row_indices, int n_rows,
                         int *ptr, int *cols, float *vals, long long *nnz);
/* CMRS keeps its entries in CSR order (cmrs.c:84-113): only the row pointer has to be rebuilt.
 * B200_ERR_DOMAIN if the rows inside a strip are not sorted. */
int b200_cmrs_to_csr_ptr(b200_ctx *ctx, const int *strip_ptr, const int *row_in_strip, int n_strips, int height,
                         int n_rows, int *ptr);

/* Format advice from the row statistics: algorithmic bytes per SpMV of every format for this matrix
 * (SURVEY 8d formulas, V = value_bytes) and the format to use.  bytes[] is indexed by B200_FORMAT_*. */
typedef struct {
    int n_rows;
    long long nnz;
    int min_len, max_len;
    double mean_len;
    int skewed;                       /* max_len > 16 x mean_len: power-law-like, gather-bound */
    double sell_padding;              /* padded / nnz at sigma = 1 */
    double sell_padding_sigma65536;   /* ... with rows sorted in windows of 65536 */
    long long bytes[5];
    long long bytes_sell_sigma65536;
    long long bytes_sell16;           /* sigma = 1 with 16-bit deltas (if the matrix allows them) */
    int recommended;                  /* B200_FORMAT_* */
    int recommended_sigma;            /* for SELL */
    char reason[160];
} b200_format_advice_t;
int b200_format_advice(b200_ctx *ctx, const int *ptr, int n_rows, int n_cols, int value_bytes,
                       b200_format_advice_t *advice);

/* =====================================================================================
 * Iterated (power-iteration) mode behind the C ABI (new; SURVEY 8b viii, 8e).  The reference's
 * device loop enumerates up to 8 GPUs and breaks after the first (csr.c:12,22-30,279); here one
 * rank = one context = one GPU, ranks being processes (peers mapped with b200_ipc_*) or threads of
 * one process (peers enabled with b200_ctx_enable_peer_access).
 * ===================================================================================== */

/* ---- communicator: NCCL over NVLink, opened at run time (dlopen of libnccl.so.2; a process that
 *      already carries NCCL, e.g. through torch, shares that copy).  Without the library every call
 *      fails with B200_ERR_UNSUPPORTED -- the single-GPU entry points never need it.  Rank 0 makes an
 *      id, hands it to the other ranks by any means, and all ranks create their communicator
 *      concurrently (collective call).  All collectives run on the context's stream. ---- */
typedef struct b200_comm b200_comm;
#define B200_COMM_ID_BYTES 128
int b200_comm_get_unique_id(unsigned char id[B200_COMM_ID_BYTES]);
int b200_comm_create(b200_ctx *ctx, const unsigned char id[B200_COMM_ID_BYTES], int rank, int world,
                     b200_comm **comm);
int b200_comm_destroy(b200_comm *comm);
int b200_comm_info(const b200_comm *comm, int *rank, int *world, int *nccl_version); /* any out may be NULL */
/* asynchronous-error poll (ncclCommGetAsyncError): B200_ERR_COMM once a peer died or a link failed;
 * b200_iterator_run / b200_iterator_norm call it too */
int b200_comm_check(b200_comm *comm);
int b200_comm_allreduce_sum_f64(b200_comm *comm, double *buf_device, long long count); /* in place */
/* in place: rank r's segment is full[r*count_per_rank .. (r+1)*count_per_rank) */
int b200_comm_allgather_f64(b200_comm *comm, double *full_device, long long count_per_rank);
/* the same for any element type: rank r's segment is the bytes [r*bytes_per_rank, (r+1)*bytes_per_rank) of
 * `full_device`.  What it is for: a replicated x that arrives from the HOST (single-shot SpMV over a
 * row-partitioned matrix) crosses the host link once -- every rank uploads 1/N of it -- and reaches the other
 * ranks over NVLink, instead of every rank uploading all of it. */
int b200_comm_allgather_bytes(b200_comm *comm, void *full_device, long long bytes_per_rank);
/* threads-of-one-process ranks: let this context's device store into `peer_device`'s allocations */
int b200_ctx_enable_peer_access(b200_ctx *ctx, int peer_device);

/* ---- NVSwitch multicast block (new): one 2 MiB block of every rank's GPU memory bound to
 *      ONE multicast object; a multimem.st / multimem.red to its address is carried out by the switch on
 *      every GPU's copy.  The iterated mode uses it for the per-step all-reduce of ||y||^2 (one store per
 *      rank into its own slot, slots added in rank order after the barrier: bit-identical on every rank)
 *      and the barrier that orders the halo stores -- one 32-thread kernel per step, no NCCL call
 *      (B200_ITER_FUSED_MCAST).  Needs NVLink multicast support (b200_mcast_supported); every entry point
 *      returns B200_ERR_UNSUPPORTED where the driver, the device or the container's permissions lack it.
 *      Set-up, in this order:  rank 0 b200_mcast_create;  other ranks b200_mcast_import_pid_fd (other
 *      processes: rank 0's pid and the descriptor it got), b200_mcast_import_fd (a descriptor received over
 *      a socket) or b200_mcast_share (threads of rank 0's process);  ALL ranks b200_mcast_add_device;
 *      barrier;  ALL ranks b200_mcast_bind;  barrier;  use;  b200_mcast_destroy. ---- */
typedef struct b200_mcast b200_mcast;
int b200_mcast_supported(b200_ctx *ctx, int *supported);
int b200_mcast_create(b200_ctx *ctx, int world, b200_mcast **mcast, int *export_fd /* may be NULL */);
int b200_mcast_import_fd(b200_ctx *ctx, int world, int fd, b200_mcast **mcast);
int b200_mcast_import_pid_fd(b200_ctx *ctx, int world, int owner_pid, int owner_fd, b200_mcast **mcast);
int b200_mcast_share(b200_ctx *ctx, const b200_mcast *owner, b200_mcast **mcast);
int b200_mcast_add_device(b200_mcast *mcast);
int b200_mcast_bind(b200_mcast *mcast, int rank /* this rank's slot: 0 .. world-1, world <= 16 */);
/* the multicast address (stores / reductions reach every rank's copy) and this rank's own copy */
int b200_mcast_pointers(const b200_mcast *mcast, void **multicast_ptr, void **local_ptr, size_t *bytes);
/* the two 32-slot device arrays of a step: the accumulator the fused SpMV kernel adds ||y_r||^2 into
 * (sumsq_out) and the sums over all ranks that b200_mcast_allreduce_barrier leaves for the next kernel
 * (scale_sumsq) */
int b200_mcast_step_buffers(const b200_mcast *mcast, double **sumsq_accumulator, const double **summed_over_ranks);
/* one launch on the context's queue: all-reduce of the accumulator over the ranks through the switch
 * (multimem.red), arrival signal to every rank, wait for all ranks.  Work queued after it sees every
 * rank's stores of the step.  All ranks must call it the same number of times; a rank that never arrives
 * is reported as B200_ERR_CUDA by the next b200_sync after 2 s, not a hang. */
int b200_mcast_allreduce_barrier(b200_mcast *mcast);
int b200_mcast_destroy(b200_mcast *mcast);

/* which rows of rank `rank`'s block every rank reads as columns (host function).  col_min[d] /
 * col_max[d] = smallest / largest global column index in rank d's matrix block (b200_minmax_i32 over
 * its column array, exchanged once).  Blocks are equal: rank r owns rows [r*rows_per_rank,
 * min((r+1)*rows_per_rank, n_rows_total)).  lo[d], hi[d] (world entries each) = LOCAL row range of this
 * rank's block that destination d needs; the rank itself always gets its whole block. */
int b200_halo_rows(const int *col_min, const int *col_max, int world, int rank, long long rows_per_rank,
                   long long n_rows_total, int *lo, int *hi);

/* ---- iterator: `steps` steps of  y = A_r x / ||x||_2 ;  x <- y  on row-partitioned A ---- */
typedef struct {            /* this rank's row block, fp64, GLOBAL column indices, device arrays */
    int format;             /* B200_FORMAT_CSR or B200_FORMAT_SELL (chunk 32, no permutation) */
    int n_rows;             /* rows of the block */
    int n_slices;           /* SELL: chunks of the block (ceil(n_rows / 32)) */
    const int *ptr;         /* CSR: ptr[n_rows + 1];  SELL: row_indices[n_slices + 1] */
    const int *indices;
    const double *data;
    const b200_csr_plan *csr_plan; /* CSR: optional plan (NULL = analyse on the fly, not recordable) */
} b200_block_f64;
enum {
    B200_ITER_FUSED = 0,    /* SELL kernel stores y straight into the x buffers of the ranks that read
                             * it (b200_spmv_sell_halo_f64) + one 256-byte all-reduce per step */
    B200_ITER_ALLGATHER = 1, /* SpMV -> sum of squares -> all-reduce -> scale -> in-place ncclAllGather */
    B200_ITER_FUSED_MCAST = 2 /* B200_ITER_FUSED with the all-reduce + barrier done through NVSwitch multicast
                               * (b200_mcast_allreduce_barrier) instead of NCCL: two launches per step */
};
typedef struct {
    int mode;               /* B200_ITER_* */
    int world, rank;
    long long rows_per_rank; /* multiple of 32, >= n_rows: rank r's block sits at r*rows_per_rank of x */
    double *const *x[2];    /* x[b][r]: x buffer b (two alternate) of rank r as THIS process sees it, each
                             * world*rows_per_rank doubles; x[0][rank] holds the start
                             * vector.  B200_ITER_ALLGATHER reads only x[b][rank] */
    const int *halo_lo, *halo_hi; /* FUSED: b200_halo_rows output (world entries); NULL, NULL = every row
                             * to every rank */
    int graph_steps;        /* 0: every step is issued as separate launches; G (even): steps are recorded
                             * once into a launch graph of G steps and replayed (first step always direct) */
    b200_mcast *mcast;      /* B200_ITER_FUSED_MCAST: this rank's bound multicast block */
} b200_iter_desc;
typedef struct b200_iterator b200_iterator;
int b200_iterator_create(b200_ctx *ctx, b200_comm *comm /* NULL iff world == 1 or FUSED_MCAST */, const b200_block_f64 *block,
                         const b200_iter_desc *desc, b200_iterator **iterator);
/* enqueue `steps` more steps (asynchronous; all ranks must call it with the same counts) */
int b200_iterator_run(b200_iterator *iterator, int steps);
/* waits; *norm = ||A x_{k-1} / ||x_{k-1}|| ||_2 of the last step issued: the eigenvalue estimate */
int b200_iterator_norm(b200_iterator *iterator, double *norm);
/* steps issued so far, the buffer holding the current vector (own block + received rows; FUSED: not
 * yet divided by the norm), kernels + collectives issued so far; any out may be NULL */
int b200_iterator_state(const b200_iterator *iterator, unsigned long long *steps_done, double **x_current,
                        unsigned long long *launches);
int b200_iterator_destroy(b200_iterator *iterator);

#ifdef __cplusplus
}
#endif
#endif /* B200SPMV_H */
