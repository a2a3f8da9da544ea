/*
 * shim/CL/cl.h -- the subset of the Khronos OpenCL C API that the reference drivers use
 * (21 entry points), declared so that the UNMODIFIED reference sources
 * compile against libOpenCL_b200.so (opencl_b200.c in this directory), which forwards every call
 * to the B200 C ABI (include/b200spmv.h).  The image ships no OpenCL headers; these declarations
 * follow the public OpenCL 3.0 specification.
 */
#ifndef B200_SHIM_CL_H
#define B200_SHIM_CL_H

#include <stddef.h>
#include <stdint.h>

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef double cl_double;
typedef float cl_float;
typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_queue_properties;
typedef intptr_t cl_context_properties;
typedef cl_uint cl_program_build_info;

typedef struct b200cl_platform *cl_platform_id;
typedef struct b200cl_device *cl_device_id;
typedef struct b200cl_context *cl_context;
typedef struct b200cl_queue *cl_command_queue;
typedef struct b200cl_mem *cl_mem;
typedef struct b200cl_program *cl_program;
typedef struct b200cl_kernel *cl_kernel;
typedef struct b200cl_event *cl_event;

#define CL_SUCCESS 0
#define CL_DEVICE_NOT_FOUND (-1)
#define CL_OUT_OF_RESOURCES (-5)
#define CL_INVALID_VALUE (-30)
#define CL_INVALID_KERNEL_NAME (-46)
#define CL_INVALID_KERNEL_ARGS (-52)
#define CL_FALSE 0
#define CL_TRUE 1
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_PROGRAM_BUILD_LOG 0x1183

cl_int clGetPlatformIDs(cl_uint num_entries, cl_platform_id *platforms, cl_uint *num_platforms);
cl_int clGetDeviceIDs(cl_platform_id platform, cl_device_type type, cl_uint num_entries,
                      cl_device_id *devices, cl_uint *num_devices);
cl_context clCreateContext(const cl_context_properties *props, cl_uint num_devices,
                           const cl_device_id *devices,
                           void (*notify)(const char *, const void *, size_t, void *),
                           void *user_data, cl_int *err);
cl_command_queue clCreateCommandQueueWithProperties(cl_context ctx, cl_device_id dev,
                                                    const cl_queue_properties *props, cl_int *err);
cl_mem clCreateBuffer(cl_context ctx, cl_mem_flags flags, size_t size, void *host_ptr,
                      cl_int *err);
cl_program clCreateProgramWithSource(cl_context ctx, cl_uint count, const char **strings,
                                     const size_t *lengths, cl_int *err);
cl_int clBuildProgram(cl_program prog, cl_uint num_devices, const cl_device_id *devices,
                      const char *options, void (*notify)(cl_program, void *), void *user_data);
cl_int clGetProgramBuildInfo(cl_program prog, cl_device_id dev, cl_program_build_info name,
                             size_t size, void *value, size_t *size_ret);
cl_kernel clCreateKernel(cl_program prog, const char *name, cl_int *err);
cl_int clSetKernelArg(cl_kernel k, cl_uint index, size_t size, const void *value);
cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset,
                            size_t size, const void *ptr, cl_uint n_wait, const cl_event *wait,
                            cl_event *event);
cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem buf, cl_bool blocking, size_t offset,
                           size_t size, void *ptr, cl_uint n_wait, const cl_event *wait,
                           cl_event *event);
cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint work_dim,
                              const size_t *global_offset, const size_t *global_size,
                              const size_t *local_size, cl_uint n_wait, const cl_event *wait,
                              cl_event *event);
cl_int clWaitForEvents(cl_uint n, const cl_event *events);
cl_int clFinish(cl_command_queue q);
cl_int clFlush(cl_command_queue q);
cl_int clReleaseMemObject(cl_mem m);
cl_int clReleaseCommandQueue(cl_command_queue q);
cl_int clReleaseKernel(cl_kernel k);
cl_int clReleaseProgram(cl_program p);
cl_int clReleaseContext(cl_context c);

#endif
