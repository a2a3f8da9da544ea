/*
 * ell -- drop-in for the reference's ./bin/ell (ell.c): same flow, same stdout, same exit codes.
 *
 *   load databases/cant-sorted.mtx, parse                      ell.c:54-66,73-101
 *   row statistics line "average column length ..."            ell.c:68-104  (the reference never
 *       counts the LAST row in these three numbers; b200_row_length_stats' *_excl_last fields
 *       reproduce that so the line is byte-identical)
 *   FORMAT BUILD on the GPU: K-padded ELL, padding (col 0, 0)   ell.c:118-164 -> b200_build_ell_*
 *   timed launch                                               ell.c:270-280 -> b200_spmv_ellcm_*
 *       default (--rowmajor): the sub-warp-per-row kernel that consumes the reference's row-major
 *       arrays as they are (the fastest ELL kernel measured on B200, profiles/);
 *       --colmajor: the thread-per-row coalesced kernel on the transposed device layout
 *   read back, check_result, "CPU calculations" block          ell.c:290-330,357-383
 */
#include <stdio.h>
#include <stdlib.h>

#include "helper_functions.h"

void compute_using_cpu(double *data, double *vect, int *cols, int number_of_rows, int longest_col,
                       int number_of_nonzeroes, double **result);

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
    const size_t V = opt.use_f32 ? sizeof(float) : sizeof(double);

    b200_ctx *ctx = NULL;
    B200_TRY(b200_ctx_create(opt.device, &ctx));
    rc = driver_upload_triples(ctx, &m, opt.use_f32, &d);
    if (rc != Success) return rc;

    /* statistics pass (device) */
    void *buffer_ptr;
    b200_row_stats st;
    B200_TRY(b200_malloc(ctx, sizeof(int) * ((size_t)number_of_rows + 1), &buffer_ptr));
    B200_TRY(b200_check_sorted_rows(ctx, (const int *)d.rows, number_of_nonzeroes, number_of_rows));
    B200_TRY(b200_build_csr_ptr(ctx, (const int *)d.rows, number_of_nonzeroes, number_of_rows, (int *)buffer_ptr));
    B200_TRY(b200_row_length_stats(ctx, (const int *)buffer_ptr, number_of_rows, &st));
    double average_col_len = (double)st.sum_len_excl_last / (double)number_of_rows;
    printf("average column length %lf, shortest col %d, longest col %d\n", average_col_len,
           st.min_len_excl_last, st.max_len_excl_last);
    /* the reference sizes the rows with max_len_excl_last and overruns if the last row is longer;
     * the true maximum is identical on its well-defined domain */
    const int longest_col = st.max_len;
    const size_t slots = (size_t)number_of_rows * (size_t)longest_col;
    const int pitch = (number_of_rows + 31) / 32 * 32;

    /* fill pass (device): row-major = the reference's arrays; column-major = the kernel's layout */
    void *buffer_data, *buffer_indices, *buffer_data_cm = NULL, *buffer_indices_cm = NULL, *buffer_output;
    B200_TRY(b200_malloc(ctx, V * slots, &buffer_data));
    B200_TRY(b200_malloc(ctx, sizeof(int) * slots, &buffer_indices));
    B200_TRY(b200_malloc(ctx, V * (size_t)number_of_rows, &buffer_output));
    if (opt.use_f32)
        B200_TRY(b200_build_ell_f32(ctx, (const int *)buffer_ptr, (const int *)d.cols, (const double *)d.data64,
                                    number_of_rows, longest_col, (int *)buffer_indices, (float *)buffer_data));
    else
        B200_TRY(b200_build_ell_f64(ctx, (const int *)buffer_ptr, (const int *)d.cols, (const double *)d.data64,
                                    number_of_rows, longest_col, (int *)buffer_indices, (double *)buffer_data));
    if (!opt.rowmajor) {
        B200_TRY(b200_malloc(ctx, V * (size_t)pitch * (size_t)longest_col, &buffer_data_cm));
        B200_TRY(b200_malloc(ctx, sizeof(int) * (size_t)pitch * (size_t)longest_col, &buffer_indices_cm));
        if (opt.use_f32)
            B200_TRY(b200_build_ell_colmajor_f32(ctx, (const int *)buffer_ptr, (const int *)d.cols,
                                                 (const double *)d.data64, number_of_rows, longest_col, pitch,
                                                 (int *)buffer_indices_cm, (float *)buffer_data_cm));
        else
            B200_TRY(b200_build_ell_colmajor_f64(ctx, (const int *)buffer_ptr, (const int *)d.cols,
                                                 (const double *)d.data64, number_of_rows, longest_col, pitch,
                                                 (int *)buffer_indices_cm, (double *)buffer_data_cm));
    }
    B200_TRY(b200_sync(ctx));

#define LAUNCH()                                                                                         \
    (opt.rowmajor                                                                                        \
         ? (opt.use_f32 ? b200_spmv_ell_f32(ctx, (const float *)buffer_data, (const int *)buffer_indices, \
                                            (const float *)d.vect, (float *)buffer_output,                \
                                            number_of_rows, longest_col)                                  \
                        : b200_spmv_ell_f64(ctx, (const double *)buffer_data, (const int *)buffer_indices, \
                                            (const double *)d.vect, (double *)buffer_output,              \
                                            number_of_rows, longest_col))                                 \
         : (opt.use_f32 ? b200_spmv_ellcm_f32(ctx, (const float *)buffer_data_cm,                         \
                                              (const int *)buffer_indices_cm, (const float *)d.vect,      \
                                              (float *)buffer_output, number_of_rows, longest_col, pitch) \
                        : b200_spmv_ellcm_f64(ctx, (const double *)buffer_data_cm,                        \
                                              (const int *)buffer_indices_cm, (const double *)d.vect,     \
                                              (double *)buffer_output, number_of_rows, longest_col, pitch)))

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
    if (error != B200_SUCCESS) return report_b200_error("b200_spmv_ell", error);

    /* read output */
    double *output = (double *)malloc(sizeof(double) * (size_t)number_of_rows + 16);
    rc = driver_read_output(ctx, buffer_output, number_of_rows, opt.use_f32, output);
    if (rc != Success) return rc;
    if (check_result(opt.matrix, m.vect, output) == true) printf("result is ok\n");
    else printf("result is wrong\n");

    /* CPU: over the GPU-built row-major arrays, padding included, as the reference does */
    if (!opt.no_cpu) {
        int *ell_cols = (int *)malloc(sizeof(int) * slots + 16);
        double *ell_data = (double *)malloc(sizeof(double) * slots + 16);
        double *output_cpu = (double *)calloc((size_t)number_of_rows + 1, sizeof(double));
        B200_TRY(b200_memcpy_d2h(ctx, ell_cols, buffer_indices, sizeof(int) * slots));
        if (opt.use_f32) {
            float *tmp = (float *)malloc(sizeof(float) * slots + 16);
            B200_TRY(b200_memcpy_d2h(ctx, tmp, buffer_data, sizeof(float) * slots));
            for (size_t k = 0; k < slots; ++k) ell_data[k] = tmp[k];
            free(tmp);
        } else {
            B200_TRY(b200_memcpy_d2h(ctx, ell_data, buffer_data, sizeof(double) * slots));
        }
        compute_using_cpu(ell_data, m.vect, ell_cols, number_of_rows, longest_col, number_of_nonzeroes, &output_cpu);
        if (check_result(opt.matrix, m.vect, output_cpu) == true) printf("cpu result is ok\n");
        else printf("cpu result is wrong\n");
        free(ell_cols);
        free(ell_data);
        free(output_cpu);
    }

    /* release memory */
    b200_free(ctx, buffer_ptr);
    b200_free(ctx, buffer_data);
    b200_free(ctx, buffer_indices);
    b200_free(ctx, buffer_data_cm);
    b200_free(ctx, buffer_indices_cm);
    b200_free(ctx, buffer_output);
    driver_free_triples(ctx, &d);
    b200_ctx_destroy(ctx);
    driver_free_matrix(&m);
    free(output);
    return Success;
}

void compute_using_cpu(double *data, double *vect, int *cols, int number_of_rows, int longest_col,
                       int number_of_nonzeroes, double **result)
{
    double start = now_ms();
#pragma omp parallel for
    for (int row = 0; row < number_of_rows; ++row) {
        const size_t offset = (size_t)row * (size_t)longest_col;
        double sum = 0.0;
        for (int k = 0; k < longest_col; ++k) sum += data[offset + k] * vect[cols[offset + k]];
        (*result)[row] = sum;
    }
    double ms = now_ms() - start;
    printf("\nCPU calculations\n");
    calculate_and_print_performance(ms, number_of_nonzeroes);
}
