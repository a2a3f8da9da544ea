/*
 * opencl_b200.c -- libOpenCL_b200.so: an OpenCL-symbol shim over the B200 C ABI.
 *
 * Exports the 21 OpenCL entry points the reference drivers call and forwards them to
 * libb200spmv.so, so that the UNMODIFIED reference programs (coo.c, csr.c, ell.c, sigma_c.c,
 * cmrs.c, compiled against shim/CL/cl.h and linked with this library instead of libOpenCL.so,
 * reference Makefile:5-6) run their SpMV on the B200 kernels.  The reference's own check_result
 * then validates the GPU result in the same process ("result is ok"), next to its own CPU path.
 *
 *   cl_context + cl_command_queue   -> one b200_ctx (device 0, one in-order stream)
 *   cl_mem                          -> device pointer + size (b200_malloc)
 *   clCreateProgramWithSource / clBuildProgram: accepted and ignored -- the kernels are compiled
 *     into libb200spmv.so; the .cl text the driver read from kernels/ is not used
 *   clCreateKernel(name)            -> "coo" | "csr" | "ell" | "sigma_c" | "cmrs"
 *   clSetKernelArg                  -> recorded (cl_mem handles by value, ints by value)
 *   clEnqueueNDRangeKernel          -> b200_spmv_<fmt>_f64 with the recorded arguments; the
 *     OpenCL launch shape (global/local size) is ignored, sizes the OpenCL kernel took from its
 *     launch shape are derived from the buffer sizes
 * Only what those five programs need is implemented; anything else returns CL_INVALID_VALUE.
 */
#include <CL/cl.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200spmv.h"

struct b200cl_platform { int unused; };
struct b200cl_device { int index; };
struct b200cl_context { b200_ctx *ctx; int refs; };
struct b200cl_queue { struct b200cl_context *context; };
struct b200cl_program { int unused; };
struct b200cl_event { int unused; };
struct b200cl_mem {
    struct b200cl_context *context;
    void *dptr;
    size_t size;
};
#define B200CL_MAX_ARGS 12
struct b200cl_kernel {
    char name[32];
    struct b200cl_mem *mem[B200CL_MAX_ARGS];
    int ival[B200CL_MAX_ARGS];
    unsigned set_mask;
    b200_csr_plan *csr_plan; /* row statistics, built at the first launch (like a JIT build) */
    const void *csr_plan_ptr;
};

static struct b200cl_platform the_platform;
static struct b200cl_device the_devices[8];
static struct b200cl_event the_event;

static cl_int to_cl(int status)
{
    switch (status) {
    case B200_SUCCESS: return CL_SUCCESS;
    case B200_ERR_NO_DEVICE: return CL_DEVICE_NOT_FOUND;
    case B200_ERR_INVALID_VALUE: return CL_INVALID_VALUE;
    default: return CL_OUT_OF_RESOURCES;
    }
}

cl_int clGetPlatformIDs(cl_uint num_entries, cl_platform_id *platforms, cl_uint *num_platforms)
{
    if (num_platforms) *num_platforms = 1;
    if (platforms && num_entries >= 1) platforms[0] = &the_platform;
    return CL_SUCCESS;
}

cl_int clGetDeviceIDs(cl_platform_id platform, cl_device_type type, cl_uint num_entries,
                      cl_device_id *devices, cl_uint *num_devices)
{
    int count = 0;
    (void)platform;
    (void)type;
    int status = b200_get_device_count(&count);
    if (count > 8) count = 8;
    if (num_devices) *num_devices = (cl_uint)count;
    if (status != B200_SUCCESS) return to_cl(status);
    for (int i = 0; devices && i < count && (cl_uint)i < num_entries; ++i) {
        the_devices[i].index = i;
        devices[i] = &the_devices[i];
    }
    return CL_SUCCESS;
}

cl_context clCreateContext(const cl_context_properties *props, cl_uint num_devices,
                           const cl_device_id *devices,
                           void (*notify)(const char *, const void *, size_t, void *),
                           void *user_data, cl_int *err)
{
    (void)props;
    (void)notify;
    (void)user_data;
    struct b200cl_context *c = (struct b200cl_context *)calloc(1, sizeof *c);
    int device = (num_devices > 0 && devices && devices[0]) ? devices[0]->index : 0;
    int status = c ? b200_ctx_create(device, &c->ctx) : B200_ERR_OUT_OF_MEMORY;
    if (err) *err = to_cl(status);
    if (status != B200_SUCCESS) {
        fprintf(stderr, "libOpenCL_b200: %s\n", b200_last_error());
        free(c);
        return NULL;
    }
    c->refs = 1;
    return c;
}

cl_command_queue clCreateCommandQueueWithProperties(cl_context ctx, cl_device_id dev,
                                                    const cl_queue_properties *props, cl_int *err)
{
    (void)dev;
    (void)props;
    struct b200cl_queue *q = (struct b200cl_queue *)calloc(1, sizeof *q);
    if (q) q->context = ctx;
    if (err) *err = (q && ctx) ? CL_SUCCESS : CL_INVALID_VALUE;
    return q;
}

cl_mem clCreateBuffer(cl_context ctx, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *err)
{
    (void)flags;
    (void)host_ptr;
    struct b200cl_mem *m = (struct b200cl_mem *)calloc(1, sizeof *m);
    int status = (m && ctx) ? b200_malloc(ctx->ctx, size, &m->dptr) : B200_ERR_INVALID_VALUE;
    if (status == B200_SUCCESS) status = b200_memset_async(ctx->ctx, m->dptr, 0, size);
    if (err) *err = to_cl(status);
    if (status != B200_SUCCESS) {
        free(m);
        return NULL;
    }
    m->context = ctx;
    m->size = size;
    return m;
}

cl_program clCreateProgramWithSource(cl_context ctx, cl_uint count, const char **strings,
                                     const size_t *lengths, cl_int *err)
{
    (void)ctx;
    (void)count;
    (void)strings;
    (void)lengths;
    if (err) *err = CL_SUCCESS;
    return (cl_program)calloc(1, sizeof(struct b200cl_program));
}

cl_int clBuildProgram(cl_program prog, cl_uint num_devices, const cl_device_id *devices,
                      const char *options, void (*notify)(cl_program, void *), void *user_data)
{
    (void)prog;
    (void)num_devices;
    (void)devices;
    (void)options;
    (void)notify;
    (void)user_data;
    return CL_SUCCESS;
}

cl_int clGetProgramBuildInfo(cl_program prog, cl_device_id dev, cl_program_build_info name,
                             size_t size, void *value, size_t *size_ret)
{
    static const char log[] = "kernels are precompiled into libb200spmv.so (sm_100a)";
    (void)prog;
    (void)dev;
    (void)name;
    if (size_ret) *size_ret = sizeof log;
    if (value && size > 0) {
        strncpy((char *)value, log, size);
        ((char *)value)[size - 1] = '\0';
    }
    return CL_SUCCESS;
}

cl_kernel clCreateKernel(cl_program prog, const char *name, cl_int *err)
{
    static const char *known[] = {"coo", "csr", "ell", "sigma_c", "cmrs"};
    (void)prog;
    for (size_t i = 0; name && i < sizeof known / sizeof known[0]; ++i) {
        if (strcmp(name, known[i]) == 0) {
            struct b200cl_kernel *k = (struct b200cl_kernel *)calloc(1, sizeof *k);
            if (k) strcpy(k->name, name);
            if (err) *err = k ? CL_SUCCESS : CL_OUT_OF_RESOURCES;
            return k;
        }
    }
    if (err) *err = CL_INVALID_KERNEL_NAME;
    return NULL;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint index, size_t size, const void *value)
{
    if (!k || index >= B200CL_MAX_ARGS) return CL_INVALID_VALUE;
    if (value == NULL) {
        /* __local scratch (ell.c:248, cmrs.c:205): the CUDA kernels need none */
    } else if (size == sizeof(cl_mem)) {
        memcpy(&k->mem[index], value, sizeof(cl_mem));
    } else if (size == sizeof(int)) {
        memcpy(&k->ival[index], value, sizeof(int));
    } else {
        return CL_INVALID_VALUE;
    }
    k->set_mask |= 1u << index;
    return CL_SUCCESS;
}

cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset,
                            size_t size, const void *ptr, cl_uint n_wait, const cl_event *wait,
                            cl_event *event)
{
    (void)n_wait;
    (void)wait;
    if (!q || !buf || offset + size > buf->size) return CL_INVALID_VALUE;
    b200_ctx *ctx = q->context->ctx;
    int status = b200_memcpy_h2d_async(ctx, (char *)buf->dptr + offset, ptr, size);
    if (status == B200_SUCCESS && blocking) status = b200_sync(ctx);
    if (event) *event = &the_event;
    return to_cl(status);
}

cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset,
                           size_t size, void *ptr, cl_uint n_wait, const cl_event *wait,
                           cl_event *event)
{
    (void)blocking; /* always blocking: the drivers read with CL_TRUE (csr.c:220) */
    (void)n_wait;
    (void)wait;
    if (!q || !buf || offset + size > buf->size) return CL_INVALID_VALUE;
    if (event) *event = &the_event;
    return to_cl(b200_memcpy_d2h(q->context->ctx, ptr, (const char *)buf->dptr + offset, size));
}

#define ARG_MEM(i) (k->mem[i] ? k->mem[i]->dptr : NULL)
#define NEED(mask) if ((k->set_mask & (mask)) != (mask)) return CL_INVALID_KERNEL_ARGS

cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint work_dim,
                              const size_t *global_offset, const size_t *global_size,
                              const size_t *local_size, cl_uint n_wait, const cl_event *wait,
                              cl_event *event)
{
    (void)work_dim;
    (void)global_offset;
    (void)global_size;
    (void)local_size;
    (void)n_wait;
    (void)wait;
    if (!q || !k) return CL_INVALID_VALUE;
    b200_ctx *ctx = q->context->ctx;
    int status;
    if (strcmp(k->name, "csr") == 0) {
        /* csr(ptr, col, data, vect, output, N)                              kernels/Csr.cl:1 */
        NEED(0x3f);
        if (k->csr_plan && k->csr_plan_ptr != ARG_MEM(0)) {
            b200_csr_plan_destroy(k->csr_plan);
            k->csr_plan = NULL;
        }
        status = B200_SUCCESS;
        if (!k->csr_plan) {
            status = b200_csr_plan_create(ctx, (const int *)ARG_MEM(0), k->ival[5], &k->csr_plan);
            k->csr_plan_ptr = ARG_MEM(0);
        }
        if (status == B200_SUCCESS)
            status = b200_spmv_csr_f64(ctx, (const int *)ARG_MEM(0), (const int *)ARG_MEM(1),
                                       (const double *)ARG_MEM(2), (const double *)ARG_MEM(3),
                                       (double *)ARG_MEM(4), k->ival[5], k->csr_plan);
    } else if (strcmp(k->name, "coo") == 0) {
        /* coo(row, col, data, vect, output, N = nnz)                        kernels/Coo.cl:24 */
        NEED(0x3f);
        status = b200_spmv_coo_f64(ctx, (const int *)ARG_MEM(0), (const int *)ARG_MEM(1),
                                   (const double *)ARG_MEM(2), (const double *)ARG_MEM(3),
                                   (double *)ARG_MEM(4), k->ival[5],
                                   (int)(k->mem[4]->size / sizeof(double)));
    } else if (strcmp(k->name, "ell") == 0) {
        /* ell(data, indices, vect, output, N, row_size, local)              kernels/Ell.cl:1 */
        NEED(0x3f);
        status = b200_spmv_ell_f64(ctx, (const double *)ARG_MEM(0), (const int *)ARG_MEM(1),
                                   (const double *)ARG_MEM(2), (double *)ARG_MEM(3), k->ival[4],
                                   k->ival[5]);
    } else if (strcmp(k->name, "sigma_c") == 0) {
        /* sigma_c(data, indices, vect, output, row_indices, C)              kernels/Sigma_C.cl:1 */
        NEED(0x3f);
        const int n_slices = (int)(k->mem[4]->size / sizeof(int)) - 1;
        status = b200_spmv_sell_f64(ctx, (const double *)ARG_MEM(0), (const int *)ARG_MEM(1),
                                    (const double *)ARG_MEM(2), (double *)ARG_MEM(3),
                                    (const int *)ARG_MEM(4), k->ival[5], n_slices,
                                    (int)(k->mem[3]->size / sizeof(double)), NULL, NULL);
    } else if (strcmp(k->name, "cmrs") == 0) {
        /* cmrs(data, indices, strip_ptr, row_in_strip, vect, output, N, height, local)  Cmrs.cl:1 */
        NEED(0xff);
        status = b200_spmv_cmrs_f64(ctx, (const double *)ARG_MEM(0), (const int *)ARG_MEM(1),
                                    (const int *)ARG_MEM(2), (const int *)ARG_MEM(3),
                                    (const double *)ARG_MEM(4), (double *)ARG_MEM(5), k->ival[6],
                                    k->ival[7], (int)(k->mem[5]->size / sizeof(double)), NULL);
    } else {
        return CL_INVALID_KERNEL_NAME;
    }
    if (status != B200_SUCCESS) fprintf(stderr, "libOpenCL_b200: %s\n", b200_last_error());
    if (event) *event = &the_event;
    return to_cl(status);
}

cl_int clWaitForEvents(cl_uint n, const cl_event *events)
{
    (void)n;
    (void)events;
    return CL_SUCCESS; /* completion is observed by the clFinish that always follows (csr.c:202-203) */
}

cl_int clFinish(cl_command_queue q) { return q ? to_cl(b200_sync(q->context->ctx)) : CL_INVALID_VALUE; }
cl_int clFlush(cl_command_queue q) { (void)q; return CL_SUCCESS; }

cl_int clReleaseMemObject(cl_mem m)
{
    if (!m) return CL_INVALID_VALUE;
    int status = b200_free(m->context->ctx, m->dptr);
    free(m);
    return to_cl(status);
}

cl_int clReleaseCommandQueue(cl_command_queue q)
{
    /* sigma_c.c:370-371 calls clFinish(command_queue) AFTER clReleaseCommandQueue(command_queue); a
     * real runtime survives that through reference counting, so the 8-byte queue object is simply
     * kept alive for the life of the process */
    (void)q;
    return CL_SUCCESS;
}
cl_int clReleaseKernel(cl_kernel k)
{
    if (k && k->csr_plan) b200_csr_plan_destroy(k->csr_plan);
    free(k);
    return CL_SUCCESS;
}
cl_int clReleaseProgram(cl_program p) { free(p); return CL_SUCCESS; }

cl_int clReleaseContext(cl_context c)
{
    if (!c) return CL_INVALID_VALUE;
    /* the reference leaks buffer_output in three drivers (csr.c:260-263): destroying the context
     * here is still safe, device memory is reclaimed at process exit */
    b200_ctx_destroy(c->ctx);
    free(c);
    return CL_SUCCESS;
}
