/*
 * oracle/fake_cl/fake_opencl.c -- TEST SCAFFOLDING (see CL/cl.h in this directory).
 *
 * Host-memory fake of the OpenCL runtime calls the reference drivers make.  Purpose: run the
 * UNMODIFIED reference host code and RECORD what it uploads -- i.e. the reference's format
 * arrays, bit for bit, in upload order -- plus the scalar kernel arguments and launch shape.
 *
 * Recording is on when the environment variable FAKECL_DUMP_DIR names a directory:
 *   upload_<n>.bin   payload of the n-th clEnqueueWriteBuffer (n from 0)
 *   manifest.txt     one line per event: "upload n bytes", "arg kernel idx size value",
 *                    "launch kernel global local", "read bytes"
 * Kernel launches compute nothing: device output buffers stay zero-filled, so the reference's
 * "result is wrong" line for its GPU section is expected; its CPU section is real.
 */
#include <CL/cl.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct fake_cl_platform {
    int unused;
};
struct fake_cl_device {
    int unused;
};
struct fake_cl_context {
    int unused;
};
struct fake_cl_queue {
    int unused;
};
struct fake_cl_program {
    int unused;
};
struct fake_cl_event {
    int unused;
};
struct fake_cl_mem {
    size_t size;
    unsigned char *bytes;
};
struct fake_cl_kernel {
    char name[64];
};

static struct fake_cl_platform the_platform;
static struct fake_cl_device the_device;
static struct fake_cl_event the_event;
static int upload_counter = 0;

static FILE *manifest(void)
{
    static FILE *f = NULL;
    static int tried = 0;
    if (!tried) {
        const char *dir = getenv("FAKECL_DUMP_DIR");
        tried = 1;
        if (dir && *dir) {
            char path[4096];
            snprintf(path, sizeof path, "%s/manifest.txt", dir);
            f = fopen(path, "w");
        }
    }
    return f;
}

static void note(const char *fmt, ...)
    __attribute__((format(printf, 1, 2)));

#include <stdarg.h>
static void note(const char *fmt, ...)
{
    FILE *f = manifest();
    va_list ap;
    if (!f) return;
    va_start(ap, fmt);
    vfprintf(f, fmt, ap);
    va_end(ap);
    fflush(f);
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
    (void)platform;
    (void)type;
    if (num_devices) *num_devices = 1;
    if (devices && num_entries >= 1) devices[0] = &the_device;
    return CL_SUCCESS;
}

cl_context clCreateContext(const cl_context_properties *props, cl_uint num_devices,
                           const cl_device_id *devices,
                           void (*notify)(const char *, const void *, size_t, void *),
                           void *user_data, cl_int *err)
{
    (void)props;
    (void)num_devices;
    (void)devices;
    (void)notify;
    (void)user_data;
    if (err) *err = CL_SUCCESS;
    return (cl_context)calloc(1, sizeof(struct fake_cl_context));
}

cl_command_queue clCreateCommandQueueWithProperties(cl_context ctx, cl_device_id dev,
                                                    const cl_queue_properties *props, cl_int *err)
{
    (void)ctx;
    (void)dev;
    (void)props;
    if (err) *err = CL_SUCCESS;
    return (cl_command_queue)calloc(1, sizeof(struct fake_cl_queue));
}

cl_mem clCreateBuffer(cl_context ctx, cl_mem_flags flags, size_t size, void *host_ptr, cl_int *err)
{
    struct fake_cl_mem *m = (struct fake_cl_mem *)calloc(1, sizeof *m);
    (void)ctx;
    (void)flags;
    (void)host_ptr;
    m->size = size;
    m->bytes = (unsigned char *)calloc(size ? size : 1, 1);
    if (err) *err = CL_SUCCESS;
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
    return (cl_program)calloc(1, sizeof(struct fake_cl_program));
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
    (void)prog;
    (void)dev;
    (void)name;
    if (size_ret) *size_ret = 1;
    if (value && size >= 1) ((char *)value)[0] = '\0';
    return CL_SUCCESS;
}

cl_kernel clCreateKernel(cl_program prog, const char *name, cl_int *err)
{
    struct fake_cl_kernel *k = (struct fake_cl_kernel *)calloc(1, sizeof *k);
    (void)prog;
    strncpy(k->name, name ? name : "?", sizeof k->name - 1);
    if (err) *err = CL_SUCCESS;
    return k;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint index, size_t size, const void *value)
{
    if (value && size == sizeof(int)) {
        int v;
        memcpy(&v, value, sizeof v);
        note("arg %s %u %zu %d\n", k->name, index, size, v);
    } else {
        note("arg %s %u %zu -\n", k->name, index, size);
    }
    return CL_SUCCESS;
}

cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset,
                            size_t size, const void *ptr, cl_uint n_wait, const cl_event *wait,
                            cl_event *event)
{
    const char *dir = getenv("FAKECL_DUMP_DIR");
    (void)q;
    (void)blocking;
    (void)n_wait;
    (void)wait;
    if (offset + size > buf->size) return CL_INVALID_VALUE;
    memcpy(buf->bytes + offset, ptr, size);
    if (dir && *dir) {
        char path[4096];
        FILE *f;
        snprintf(path, sizeof path, "%s/upload_%d.bin", dir, upload_counter);
        f = fopen(path, "wb");
        if (f) {
            fwrite(ptr, 1, size, f);
            fclose(f);
        }
        note("upload %d %zu\n", upload_counter, size);
    }
    ++upload_counter;
    if (event) *event = &the_event;
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset,
                           size_t size, void *ptr, cl_uint n_wait, const cl_event *wait,
                           cl_event *event)
{
    (void)q;
    (void)blocking;
    (void)n_wait;
    (void)wait;
    if (offset + size > buf->size) return CL_INVALID_VALUE;
    memcpy(ptr, buf->bytes + offset, size);
    note("read %zu\n", size);
    if (event) *event = &the_event;
    return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint work_dim,
                              const size_t *global_offset, const size_t *global_size,
                              const size_t *local_size, cl_uint n_wait, const cl_event *wait,
                              cl_event *event)
{
    (void)q;
    (void)work_dim;
    (void)global_offset;
    (void)n_wait;
    (void)wait;
    note("launch %s %zu %zu\n", k->name, global_size ? global_size[0] : 0,
         local_size ? local_size[0] : 0);
    if (event) *event = &the_event;
    return CL_SUCCESS;
}

cl_int clWaitForEvents(cl_uint n, const cl_event *events)
{
    (void)n;
    (void)events;
    return CL_SUCCESS;
}

cl_int clFinish(cl_command_queue q)
{
    (void)q;
    return CL_SUCCESS;
}

cl_int clFlush(cl_command_queue q)
{
    (void)q;
    return CL_SUCCESS;
}

cl_int clReleaseMemObject(cl_mem m)
{
    if (m) {
        free(m->bytes);
        free(m);
    }
    return CL_SUCCESS;
}

cl_int clReleaseCommandQueue(cl_command_queue q)
{
    free(q);
    return CL_SUCCESS;
}

cl_int clReleaseKernel(cl_kernel k)
{
    free(k);
    return CL_SUCCESS;
}

cl_int clReleaseProgram(cl_program p)
{
    free(p);
    return CL_SUCCESS;
}

cl_int clReleaseContext(cl_context c)
{
    free(c);
    return CL_SUCCESS;
}
