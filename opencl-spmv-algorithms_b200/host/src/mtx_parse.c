/* mtx_parse -- loads a MatrixMarket file with the drivers' own reader (read_size_of_matrices_from_file
 * + read_entries) and dumps the triples as raw binary (int32 rows, int32 cols, float64 values) so
 * that tests can compare the fast parallel parse with the reference-style fscanf parse bit for bit.
 *   mtx_parse FILE.mtx OUT.bin [--expand-symmetric]   -> prints "rows cols nnz milliseconds" */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "helper_functions.h"

int main(int argc, char **argv)
{
    if (argc != 3 && !(argc == 4 && !strcmp(argv[3], "--expand-symmetric"))) {
        fprintf(stderr, "usage: %s FILE.mtx OUT.bin [--expand-symmetric]\n", argv[0]);
        return OtherError;
    }
    int n_rows, n_cols, nnz;
    FILE *file = fopen(argv[1], "r");
    if (!file) {
        perror(argv[1]);
        return FileError;
    }
    if (!read_size_of_matrices_from_file(file, &n_rows, &n_cols, &nnz)) return FileError;
    int *rows = (int *)malloc(sizeof(int) * (size_t)nnz + 16);
    int *cols = (int *)malloc(sizeof(int) * (size_t)nnz + 16);
    double *data = (double *)malloc(sizeof(double) * (size_t)nnz + 16);
    double t0 = now_ms();
    if (!read_entries(file, nnz, rows, cols, data)) return FileError;
    if (argc == 4 && !expand_symmetric_entries(n_rows, n_cols, last_banner_symmetry(), &nnz, &rows, &cols, &data))
        return FileError;
    double ms = now_ms() - t0;
    fclose(file);
    FILE *out = fopen(argv[2], "wb");
    if (!out) return FileError;
    fwrite(rows, sizeof(int), (size_t)nnz, out);
    fwrite(cols, sizeof(int), (size_t)nnz, out);
    fwrite(data, sizeof(double), (size_t)nnz, out);
    fclose(out);
    printf("%d %d %d %.3f\n", n_rows, n_cols, nnz, ms);
    return Success;
}
