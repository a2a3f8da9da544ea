/*
 * helper_functions.h -- shared host helpers of the five drivers.
 *
 * Same names, argument meaning and stdout text as the reference's inc/helper_functions.h for the
 * helpers that survive the move from OpenCL to the B200 C ABI:
 *   read_size_of_matrices_from_file   (reference :134-165)
 *   calculate_and_print_performance   (reference :167-173)   GFLOP/s = 2*nnz / ms * 1e-6
 *   calculate_and_print_speed         (reference :175-182)   "GB" = nnz*8 ... 2*nnz*8
 *   check_result                      (reference :184-236)   serial fp64 COO re-read of the file
 * The OpenCL-only helpers (read_source_from_cl_file, read_build_program_info, get_device_ids)
 * have no counterpart: kernels are compiled into libb200spmv.so and devices are enumerated with
 * b200_get_device_count.  Unlike the reference header, this one only DECLARES; the definitions
 * live in src/driver_common.c.
 */
#ifndef B200_HOST_HELPER_FUNCTIONS_H
#define B200_HOST_HELPER_FUNCTIONS_H

#include <stdbool.h>
#include <stdio.h>

#include "b200spmv.h"
#include "enums.h"
#include "mmio.h"

#define EPSILON 0.000001 /* the reference's absolute tolerance (:11) */
#define DEVICES_DEFAULT_SIZE 8

/* optional, additive command line (no arguments == the reference's behaviour) */
typedef struct {
    const char *matrix; /* --matrix PATH   (default: the reference's hard-coded file) */
    int use_f32;        /* --dtype f32|f64 (default f64: the reference is fp64-only) */
    int sigma;          /* --sigma N       (sigma_c only; default 1 == reference layout) */
    int reps;           /* --reps N        timed launches, the mean is printed (default 1) */
    int device;         /* --device D */
    int no_cpu;         /* --no-cpu        skip the "CPU calculations" block */
    int rowmajor;       /* --rowmajor (default) | --colmajor   (ell only) which ELL kernel/layout runs */
    int expand_symmetric; /* --expand-symmetric  mirror the off-diagonal entries of a symmetric file (default:
                             ignore the banner's symmetry, as the reference does) */
    int cache;          /* --cache  keep "<matrix>.b200cache" (binary triples) next to the matrix */
    /* iterated, multi-GPU mode (csr and sigma_c only; src/driver_iterate.c) */
    int gpus;           /* --gpus N       devices device .. device+N-1, one host thread each (default 1) */
    int iters;          /* --iters K      K power-iteration steps instead of the single SpMV (default 0 = off) */
    int json;           /* --json         one JSON line instead of the text block */
    const char *synthetic; /* --synthetic laplace7:NXxNYxNZ  matrix generated on the devices, no file */
    int sync_mcast;     /* --sync mcast|nccl  (sigma_c) per-step all-reduce + barrier through NVSwitch multicast
                           (b200_mcast_*) instead of NCCL (default nccl) */
} driver_options;

int driver_parse_args(int argc, char **argv, const char *default_matrix, driver_options *opt);

bool read_size_of_matrices_from_file(FILE *file, int *number_of_rows, int *number_of_columns,
                                     int *number_of_nonzeroes);
/* the per-driver "%d %d %lg\n" loop (csr.c:77-83), 1-based -> 0-based; false on a short file */
bool read_entries(FILE *file, int number_of_nonzeroes, int *rows, int *cols, double *data);

/* New, optional: with set_expand_symmetric(1), driver_load_matrix and check_result mirror the
 * off-diagonal entries of files whose banner says symmetric / hermitian (skew-symmetric: negated)
 * and sort the result by (row, column).  expand_symmetric_entries does the work on malloc'ed arrays
 * (replaced on success); symmetry: 0 general (no-op), 1 symmetric, -1 skew.  last_banner_symmetry()
 * = that code for the header read last by read_size_of_matrices_from_file. */
void set_expand_symmetric(int enable);
int last_banner_symmetry(void);
bool expand_symmetric_entries(int number_of_rows, int number_of_columns, int symmetry, int *number_of_nonzeroes,
                              int **rows, int **cols, double **data);

/* New, optional: binary cache of the parsed triples ("<file>.b200cache", validated against the source's
 * size and mtime).  load_triples = header + entries of a file (malloc'ed arrays), through the cache
 * when set_use_cache(1). */
void set_use_cache(int enable);
bool load_triples(const char *filename, int *n_rows, int *n_cols, int *nnz, int **rows, int **cols, double **data);

void calculate_and_print_performance(double ms, int number_of_nonzeroes);
void calculate_and_print_speed(double ms, int number_of_nonzeroes);
void set_value_bytes(int bytes); /* 8 (default) or 4: the element size the speed line uses */

/* Re-reads the file and accumulates data[row] += value * vect[col] serially in fp64, like the
 * reference.  Passes when every row is within EPSILON absolutely (the reference's criterion) or
 * when the relative max-norm error is within the tolerance set here (BASELINE.json: 1e-12 for
 * fp64, 1e-5 for fp32). */
void set_check_tolerance(double relative_max_norm);
bool check_result(const char *filename, double *vect, double *result);

double now_ms(void);

/* the part of the flow every driver shares: device check, fopen + header + entry parse, x = ramp
 * (csr.c:22-28,54-99); returns a ReturnCode */
typedef struct {
    int n_rows, n_cols, nnz;
    int *rows, *cols; /* 0-based, file order */
    double *data, *vect;
} host_matrix;
typedef struct {
    void *rows, *cols, *data64, *vect; /* device: triples as parsed + x in the run's dtype */
} device_triples;
int driver_load_matrix(const driver_options *opt, host_matrix *m);
void driver_free_matrix(host_matrix *m);
int driver_upload_triples(b200_ctx *ctx, const host_matrix *m, int use_f32, device_triples *d);
void driver_free_triples(b200_ctx *ctx, device_triples *d);
int driver_read_output(b200_ctx *ctx, const void *buffer_output, int n, int use_f32, double *output);

/* the iterated mode: K steps of y = A x / ||x|| on opt->gpus devices (src/driver_iterate.c).  format =
 * B200_FORMAT_SELL (fused kernel + halo exchange + all-reduce) or B200_FORMAT_CSR (SpMV + ncclAllGather);
 * m = the loaded matrix, or NULL with --synthetic.  Returns a ReturnCode. */
int driver_run_iterated(const driver_options *opt, const host_matrix *m, int format, const char *driver_name);

/* prints "<what> error <status>" (+ the library's detail line) and yields the exit code */
int report_b200_error(const char *what, int status);

#define B200_TRY(call)                                                 \
    do {                                                               \
        int status__ = (call);                                         \
        if (status__ != B200_SUCCESS) return report_b200_error(#call, status__); \
    } while (0)

#endif
