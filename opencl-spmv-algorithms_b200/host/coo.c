/*
 * coo -- drop-in for the reference's ./bin/coo (coo.c): same flow, same stdout, same exit codes.
 *
 *   load databases/cant.mtx (column-major file order, NOT row sorted)   coo.c:43,54-66
 *   parse the triples = the COO arrays, in file order                   coo.c:75-84
 *   x = ramp                                                            coo.c:88-92
 *   context + buffers + uploads                                         coo.c:100-190 -> b200_* C ABI
 *   timed launch ("GPU calculations" header)                            coo.c:192-204 -> b200_spmv_coo_*
 *   read back, check_result                                             coo.c:214-230
 *   "CPU calculations" block (OpenMP + atomic, as the reference)        coo.c:238-250,280-300
 * The output vector is zero-filled by the launch (the reference never zeroes it, coo.c:120).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "helper_functions.h"

void compute_using_cpu(double *data, double *vect, int *rows, int *cols, int number_of_nonzeroes,
                       double **result);

int main(int argc, char *argv[])
{
    driver_options opt;
    if (driver_parse_args(argc, argv, "databases/cant.mtx", &opt)) return OtherError;
    if (opt.iters > 0) {
        fprintf(stderr, "the iterated mode (--iters / --gpus) is implemented by csr and sigma_c\n");
        return OtherError;
    }

    int number_of_devices = 0;
    if (b200_get_device_count(&number_of_devices) != B200_SUCCESS) {
        printf("No CUDA devices found\n");
        return OpenCLDeviceError;
    }
    if (number_of_devices > DEVICES_DEFAULT_SIZE) number_of_devices = DEVICES_DEFAULT_SIZE;
    if (opt.device >= number_of_devices) return OpenCLDeviceError;

    int number_of_rows, number_of_columns, number_of_nonzeroes, i;
    const char *filename = opt.matrix;

    /* prepare data for calculations */
    FILE *file = fopen(filename, "r");
    if (file == NULL) {
        perror(filename);
        return FileError;
    }
    if (read_size_of_matrices_from_file(file, &number_of_rows, &number_of_columns, &number_of_nonzeroes) == false) {
        fclose(file);
        return FileError;
    }
    const size_t nnz = (size_t)number_of_nonzeroes;
    int *rows = (int *)malloc(nnz * sizeof(int));
    int *cols = (int *)malloc(nnz * sizeof(int));
    double *data = (double *)malloc(nnz * sizeof(double));
    if (!read_entries(file, number_of_nonzeroes, rows, cols, data)) {
        fclose(file);
        return FileError;
    }
    fclose(file);

    double *vect = (double *)malloc(sizeof(double) * (size_t)number_of_columns);
    for (i = 0; i < number_of_columns; ++i) vect[i] = i;
    double *output = (double *)malloc(sizeof(double) * (size_t)number_of_rows);
    double *output_cpu = (double *)calloc((size_t)number_of_rows, sizeof(double));

    /* prepare the device */
    const size_t V = opt.use_f32 ? sizeof(float) : sizeof(double);
    b200_ctx *ctx = NULL;
    B200_TRY(b200_ctx_create(opt.device, &ctx));
    void *buffer_row, *buffer_col, *buffer_data64, *buffer_data, *buffer_vect, *buffer_output;
    B200_TRY(b200_malloc(ctx, sizeof(int) * nnz, &buffer_row));
    B200_TRY(b200_malloc(ctx, sizeof(int) * nnz, &buffer_col));
    B200_TRY(b200_malloc(ctx, sizeof(double) * nnz, &buffer_data64));
    B200_TRY(b200_malloc(ctx, V * (size_t)number_of_columns, &buffer_vect));
    B200_TRY(b200_malloc(ctx, V * (size_t)number_of_rows, &buffer_output));
    B200_TRY(b200_memcpy_h2d_async(ctx, buffer_row, rows, sizeof(int) * nnz));
    B200_TRY(b200_memcpy_h2d_async(ctx, buffer_col, cols, sizeof(int) * nnz));
    B200_TRY(b200_memcpy_h2d_async(ctx, buffer_data64, data, sizeof(double) * nnz));
    buffer_data = buffer_data64;
    if (opt.use_f32) {
        B200_TRY(b200_malloc(ctx, sizeof(float) * nnz, &buffer_data));
        B200_TRY(b200_convert_f64_to_f32(ctx, (const double *)buffer_data64, (float *)buffer_data, number_of_nonzeroes));
        B200_TRY(b200_fill_ramp_f32(ctx, (float *)buffer_vect, number_of_columns));
    } else {
        B200_TRY(b200_memcpy_h2d_async(ctx, buffer_vect, vect, sizeof(double) * (size_t)number_of_columns));
    }
    B200_TRY(b200_sync(ctx));

#define LAUNCH()                                                                                   \
    (opt.use_f32 ? b200_spmv_coo_f32(ctx, (const int *)buffer_row, (const int *)buffer_col,        \
                                     (const float *)buffer_data, (const float *)buffer_vect,       \
                                     (float *)buffer_output, number_of_nonzeroes, number_of_rows)  \
                 : b200_spmv_coo_f64(ctx, (const int *)buffer_row, (const int *)buffer_col,        \
                                     (const double *)buffer_data, (const double *)buffer_vect,     \
                                     (double *)buffer_output, number_of_nonzeroes, number_of_rows))

    /* run program */
    B200_TRY(LAUNCH());
    B200_TRY(b200_sync(ctx));
    double start = now_ms();
    int error = B200_SUCCESS;
    for (i = 0; i < opt.reps && error == B200_SUCCESS; ++i) error = LAUNCH();
    if (error == B200_SUCCESS) error = b200_sync(ctx);
    double ms = (now_ms() - start) / opt.reps;

    printf("GPU calculations\n");
    calculate_and_print_performance(ms, number_of_nonzeroes);
    calculate_and_print_speed(ms, number_of_nonzeroes);
    if (error != B200_SUCCESS) return report_b200_error("b200_spmv_coo", error);

    /* read output */
    if (opt.use_f32) {
        float *tmp = (float *)malloc(sizeof(float) * (size_t)number_of_rows);
        B200_TRY(b200_memcpy_d2h(ctx, tmp, buffer_output, sizeof(float) * (size_t)number_of_rows));
        for (i = 0; i < number_of_rows; ++i) output[i] = tmp[i];
        free(tmp);
    } else {
        B200_TRY(b200_memcpy_d2h(ctx, output, buffer_output, sizeof(double) * (size_t)number_of_rows));
    }
    if (check_result(filename, vect, output) == true) printf("result is ok\n");
    else printf("result is wrong\n");

    /* CPU */
    if (!opt.no_cpu) {
        compute_using_cpu(data, vect, rows, cols, number_of_nonzeroes, &output_cpu);
        if (check_result(filename, vect, output_cpu) == true) printf("cpu result is ok\n");
        else printf("cpu result is wrong\n");
    }

    /* release memory */
    if (buffer_data != buffer_data64) b200_free(ctx, buffer_data);
    b200_free(ctx, buffer_row);
    b200_free(ctx, buffer_col);
    b200_free(ctx, buffer_data64);
    b200_free(ctx, buffer_vect);
    b200_free(ctx, buffer_output);
    b200_ctx_destroy(ctx);
    free(rows);
    free(cols);
    free(data);
    free(vect);
    free(output);
    free(output_cpu);
    return Success;
}

void compute_using_cpu(double *data, double *vect, int *rows, int *cols, int number_of_nonzeroes,
                       double **result)
{
    double start = now_ms();
#pragma omp parallel for
    for (int i = 0; i < number_of_nonzeroes; ++i) {
        const double product = data[i] * vect[cols[i]];
#pragma omp atomic
        (*result)[rows[i]] += product;
    }
    double ms = now_ms() - start;
    printf("\nCPU calculations\n");
    calculate_and_print_performance(ms, number_of_nonzeroes);
}
