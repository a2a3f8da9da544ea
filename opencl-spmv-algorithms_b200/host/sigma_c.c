/*
 * sigma_c -- drop-in for the reference's ./bin/sigma_c (sigma_c.c): same flow, stdout, exit codes.
 *
 *   load databases/cant-sorted.mtx, parse                        sigma_c.c:57-69,93-128
 *   FORMAT BUILD on the GPU: SELL-C, C = 32                       sigma_c.c:71-202
 *       slice pointer (row_indices) + column-major-in-slice fill  -> b200_build_sell_ptr / _fill_*
 *       --sigma N adds the sigma-window sort + permutation (new; the reference has none)
 *   timed launch, one warp per slice                             sigma_c.c:306-319 -> b200_spmv_sell_*
 *   read back the PADDED output (groups*32 values)               sigma_c.c:212,330
 *   check_result                                                 sigma_c.c:339-346
 * The reference has no CPU section in this driver, and neither has this one.
 */
#include <stdio.h>
#include <stdlib.h>

#include "helper_functions.h"

int main(int argc, char *argv[])
{
    driver_options opt;
    host_matrix m;
    device_triples d;
    if (driver_parse_args(argc, argv, "databases/cant-sorted.mtx", &opt)) return OtherError;
    /* --iters K [--gpus N]: where the reference's device loop breaks after the first GPU
     * (sigma_c.c:375), this one goes on -- power iteration over N devices, fused SELL kernel */
    if (opt.iters > 0 && opt.synthetic) return driver_run_iterated(&opt, NULL, B200_FORMAT_SELL, "sigma_c");
    int rc = driver_load_matrix(&opt, &m);
    if (rc != Success) return rc;
    if (opt.iters > 0) {
        rc = driver_run_iterated(&opt, &m, B200_FORMAT_SELL, "sigma_c");
        driver_free_matrix(&m);
        return rc;
    }
    const int number_of_rows = m.n_rows, number_of_nonzeroes = m.nnz;
    const int max_rows_to_check = 32; /* C */
    const size_t V = opt.use_f32 ? sizeof(float) : sizeof(double);

    b200_ctx *ctx = NULL;
    B200_TRY(b200_ctx_create(opt.device, &ctx));
    rc = driver_upload_triples(ctx, &m, opt.use_f32, &d);
    if (rc != Success) return rc;

    const int number_of_slices = b200_sell_num_slices(number_of_rows, max_rows_to_check);
    const int number_of_groups = number_of_slices;
    void *buffer_ptr, *buffer_slice_ptr, *buffer_row_indices, *buffer_perm = NULL;
    long long elements_sum = 0;
    B200_TRY(b200_malloc(ctx, sizeof(int) * ((size_t)number_of_rows + 1), &buffer_ptr));
    B200_TRY(b200_malloc(ctx, sizeof(long long) * ((size_t)number_of_slices + 1), &buffer_slice_ptr));
    B200_TRY(b200_malloc(ctx, sizeof(int) * ((size_t)number_of_slices + 1), &buffer_row_indices));
    if (opt.sigma > 1) B200_TRY(b200_malloc(ctx, sizeof(int) * (size_t)number_of_rows, &buffer_perm));
    B200_TRY(b200_check_sorted_rows(ctx, (const int *)d.rows, number_of_nonzeroes, number_of_rows));
    B200_TRY(b200_build_csr_ptr(ctx, (const int *)d.rows, number_of_nonzeroes, number_of_rows, (int *)buffer_ptr));
    B200_TRY(b200_build_sell_ptr(ctx, (const int *)buffer_ptr, number_of_rows, max_rows_to_check, opt.sigma,
                                 (int *)buffer_perm, (long long *)buffer_slice_ptr, &elements_sum));
    /* the reference's row_indices is cl_int: refuse, like an allocation failure, if it overflows */
    B200_TRY(b200_sell_ptr_to_i32(ctx, (const long long *)buffer_slice_ptr, number_of_slices, (int *)buffer_row_indices));

    void *buffer_data, *buffer_indices, *buffer_output;
    const size_t padded_rows = (size_t)number_of_groups * max_rows_to_check;
    B200_TRY(b200_malloc(ctx, V * (size_t)elements_sum, &buffer_data));
    B200_TRY(b200_malloc(ctx, sizeof(int) * (size_t)elements_sum, &buffer_indices));
    B200_TRY(b200_malloc(ctx, V * padded_rows, &buffer_output));
    if (opt.use_f32)
        B200_TRY(b200_build_sell_fill_f32(ctx, (const int *)buffer_ptr, (const int *)d.cols, (const double *)d.data64,
                                          number_of_rows, max_rows_to_check, (const int *)buffer_perm,
                                          (const long long *)buffer_slice_ptr, (int *)buffer_indices, (float *)buffer_data));
    else
        B200_TRY(b200_build_sell_fill_f64(ctx, (const int *)buffer_ptr, (const int *)d.cols, (const double *)d.data64,
                                          number_of_rows, max_rows_to_check, (const int *)buffer_perm,
                                          (const long long *)buffer_slice_ptr, (int *)buffer_indices, (double *)buffer_data));
    B200_TRY(b200_memset_async(ctx, buffer_output, 0, V * padded_rows));
    B200_TRY(b200_sync(ctx));

    /* with a permutation the padding rows have no destination: write number_of_rows results;
     * without one write the reference's padded groups*32 */
    const int n_out = buffer_perm ? number_of_rows : (int)padded_rows;
#define LAUNCH()                                                                                            \
    (opt.use_f32 ? b200_spmv_sell_f32(ctx, (const float *)buffer_data, (const int *)buffer_indices,         \
                                      (const float *)d.vect, (float *)buffer_output,                        \
                                      (const int *)buffer_row_indices, max_rows_to_check, number_of_slices, \
                                      n_out, (const int *)buffer_perm, plan)                                \
                 : b200_spmv_sell_f64(ctx, (const double *)buffer_data, (const int *)buffer_indices,        \
                                      (const double *)d.vect, (double *)buffer_output,                      \
                                      (const int *)buffer_row_indices, max_rows_to_check, number_of_slices, \
                                      n_out, (const int *)buffer_perm, plan))
    /* wide-chunk work list (empty for FEM-like inputs; splits hub chunks of power-law inputs) */
    b200_sell_plan *plan = NULL;
    B200_TRY(b200_sell_plan_create(ctx, (const int *)buffer_row_indices, number_of_slices, &plan));

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
    if (error != B200_SUCCESS) return report_b200_error("b200_spmv_sell", error);

    /* read output (padded, as the reference) */
    double *output = (double *)malloc(sizeof(double) * padded_rows + 16);
    rc = driver_read_output(ctx, buffer_output, (int)padded_rows, opt.use_f32, output);
    if (rc != Success) return rc;
    if (check_result(opt.matrix, m.vect, output) == true) printf("result is ok\n");
    else printf("result is wrong\n");

    /* release memory */
    b200_sell_plan_destroy(plan);
    b200_free(ctx, buffer_ptr);
    b200_free(ctx, buffer_slice_ptr);
    b200_free(ctx, buffer_row_indices);
    b200_free(ctx, buffer_perm);
    b200_free(ctx, buffer_data);
    b200_free(ctx, buffer_indices);
    b200_free(ctx, buffer_output);
    driver_free_triples(ctx, &d);
    b200_ctx_destroy(ctx);
    driver_free_matrix(&m);
    free(output);
    return Success;
}
