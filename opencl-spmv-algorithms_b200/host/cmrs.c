/*
 * cmrs -- drop-in for the reference's ./bin/cmrs (cmrs.c): same flow, same stdout, same exit codes.
 *
 *   load databases/cant-sorted.mtx, parse                       cmrs.c:58-70,84-90
 *   FORMAT BUILD on the GPU: strip_ptr + row_in_strip, height 8  cmrs.c:72-117 -> b200_build_cmrs
 *   timed launch, one warp per strip                            cmrs.c:227-240 -> b200_spmv_cmrs_*
 *   read back, check_result, "CPU calculations" block           cmrs.c:250-290,319-345
 */
#include <stdio.h>
#include <stdlib.h>

#include "helper_functions.h"

void compute_using_cpu(double *data, double *vect, int *strip_ptr, int *row_in_strip, int *cols,
                       int strip_ptr_size, int number_of_nonzeroes, int height, double **result);

int main(int argc, char *argv[])
{
    driver_options opt;
    host_matrix m;
    device_triples d;
    if (driver_parse_args(argc, argv, "databases/cant-sorted.mtx", &opt)) return OtherError;
    if (opt.iters > 0) {
        fprintf(stderr, "the iterated mode (--iters / --gpus) is implemented by csr and sigma_c\n");
        return OtherError;
    }
    int rc = driver_load_matrix(&opt, &m);
    if (rc != Success) return rc;
    const int number_of_rows = m.n_rows, number_of_nonzeroes = m.nnz;
    const int height = 8;
    const size_t nnz = (size_t)number_of_nonzeroes, V = opt.use_f32 ? sizeof(float) : sizeof(double);

    b200_ctx *ctx = NULL;
    B200_TRY(b200_ctx_create(opt.device, &ctx));
    rc = driver_upload_triples(ctx, &m, opt.use_f32, &d);
    if (rc != Success) return rc;

    const int number_of_strips = b200_cmrs_num_strips(number_of_rows, height);
    const int strip_ptr_size = number_of_strips + 1;
    void *buffer_ptr, *buffer_strip_ptr, *buffer_row_in_strip, *buffer_data = d.data64, *buffer_output;
    B200_TRY(b200_malloc(ctx, sizeof(int) * ((size_t)number_of_rows + 1), &buffer_ptr));
    B200_TRY(b200_malloc(ctx, sizeof(int) * (size_t)strip_ptr_size, &buffer_strip_ptr));
    B200_TRY(b200_malloc(ctx, sizeof(int) * nnz, &buffer_row_in_strip));
    B200_TRY(b200_malloc(ctx, V * (size_t)number_of_rows, &buffer_output));
    B200_TRY(b200_check_sorted_rows(ctx, (const int *)d.rows, number_of_nonzeroes, number_of_rows));
    B200_TRY(b200_build_csr_ptr(ctx, (const int *)d.rows, number_of_nonzeroes, number_of_rows, (int *)buffer_ptr));
    B200_TRY(b200_build_cmrs(ctx, (const int *)d.rows, (const int *)buffer_ptr, number_of_nonzeroes, number_of_rows,
                             height, (int *)buffer_strip_ptr, (int *)buffer_row_in_strip));
    if (opt.use_f32) {
        B200_TRY(b200_malloc(ctx, sizeof(float) * nnz, &buffer_data));
        B200_TRY(b200_convert_f64_to_f32(ctx, (const double *)d.data64, (float *)buffer_data, number_of_nonzeroes));
    }
    B200_TRY(b200_sync(ctx));

#define LAUNCH()                                                                                        \
    (opt.use_f32 ? b200_spmv_cmrs_f32(ctx, (const float *)buffer_data, (const int *)d.cols,             \
                                      (const int *)buffer_strip_ptr, (const int *)buffer_row_in_strip,  \
                                      (const float *)d.vect, (float *)buffer_output, number_of_strips,  \
                                      height, number_of_rows, plan)                                     \
                 : b200_spmv_cmrs_f64(ctx, (const double *)buffer_data, (const int *)d.cols,            \
                                      (const int *)buffer_strip_ptr, (const int *)buffer_row_in_strip,  \
                                      (const double *)d.vect, (double *)buffer_output, number_of_strips, \
                                      height, number_of_rows, plan))
    /* long-strip work list (empty for FEM-like inputs) */
    b200_cmrs_plan *plan = NULL;
    B200_TRY(b200_cmrs_plan_create(ctx, (const int *)buffer_strip_ptr, number_of_strips, &plan));

    /* run program */
    B200_TRY(LAUNCH());
    B200_TRY(b200_sync(ctx));
    double start = now_ms();
    int error = B200_SUCCESS;
    for (int i = 0; i < opt.reps && error == B200_SUCCESS; ++i) error = LAUNCH();
    if (error == B200_SUCCESS) error = b200_sync(ctx);
    double ms = (now_ms() - start) / opt.reps;
    calculate_and_print_performance(ms, number_of_nonzeroes);
    calculate_and_print_speed(ms, number_of_nonzeroes);
    if (error != B200_SUCCESS) return report_b200_error("b200_spmv_cmrs", error);

    /* read output */
    double *output = (double *)malloc(sizeof(double) * (size_t)number_of_rows + 16);
    rc = driver_read_output(ctx, buffer_output, number_of_rows, opt.use_f32, output);
    if (rc != Success) return rc;
    if (check_result(opt.matrix, m.vect, output) == true) printf("result is ok\n");
    else printf("result is wrong\n");

    /* CPU: over the GPU-built strip arrays */
    if (!opt.no_cpu) {
        int *strip_ptr = (int *)malloc(sizeof(int) * (size_t)strip_ptr_size + 16);
        int *row_in_strip = (int *)malloc(sizeof(int) * nnz + 16);
        double *output_cpu = (double *)calloc((size_t)number_of_strips * height + 1, sizeof(double));
        B200_TRY(b200_memcpy_d2h(ctx, strip_ptr, buffer_strip_ptr, sizeof(int) * (size_t)strip_ptr_size));
        B200_TRY(b200_memcpy_d2h(ctx, row_in_strip, buffer_row_in_strip, sizeof(int) * nnz));
        compute_using_cpu(m.data, m.vect, strip_ptr, row_in_strip, m.cols, strip_ptr_size, number_of_nonzeroes,
                          height, &output_cpu);
        if (check_result(opt.matrix, m.vect, output_cpu) == true) printf("cpu result is ok\n");
        else printf("cpu result is wrong\n");
        free(strip_ptr);
        free(row_in_strip);
        free(output_cpu);
    }

    /* release memory */
    b200_cmrs_plan_destroy(plan);
    if (buffer_data != d.data64) b200_free(ctx, buffer_data);
    b200_free(ctx, buffer_ptr);
    b200_free(ctx, buffer_strip_ptr);
    b200_free(ctx, buffer_row_in_strip);
    b200_free(ctx, buffer_output);
    driver_free_triples(ctx, &d);
    b200_ctx_destroy(ctx);
    driver_free_matrix(&m);
    free(output);
    return Success;
}

void compute_using_cpu(double *data, double *vect, int *strip_ptr, int *row_in_strip, int *cols,
                       int strip_ptr_size, int number_of_nonzeroes, int height, double **result)
{
    double start = now_ms();
#pragma omp parallel for
    for (int strip = 0; strip < strip_ptr_size - 1; ++strip) {
        double *rows_of_strip = *result + (size_t)strip * (size_t)height;
        for (int j = strip_ptr[strip]; j < strip_ptr[strip + 1]; ++j)
            rows_of_strip[row_in_strip[j]] += data[j] * vect[cols[j]];
    }
    double ms = now_ms() - start;
    printf("\nCPU calculations\n");
    calculate_and_print_performance(ms, number_of_nonzeroes);
}
