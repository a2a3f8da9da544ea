/* mtx_parse -- loads a MatrixMarket file with the drivers' own reader (read_size_of_matrices_from_file
 * + read_entries) and dumps the triples as raw binary (int32 rows, int32 cols, float64 values) so
 * that tests can compare the fast parallel parse with the reference-style fscanf parse bit for bit.
 *   mtx_parse FILE.mtx OUT.bin [--expand-symmetric] [--cache]   -> prints "rows cols nnz milliseconds" */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "helper_functions.h"

int main(int argc, char **argv)
{
    int expand = 0, cache = 0, bad = argc < 3;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--expand-symmetric")) expand = 1;
        else if (!strcmp(argv[i], "--cache")) cache = 1;
        else bad = 1;
    }
    if (bad) {
        fprintf(stderr, "usage: %s FILE.mtx OUT.bin [--expand-symmetric] [--cache]\n", argv[0]);
        return OtherError;
    }
    int n_rows, n_cols, nnz;
    int *rows, *cols;
    double *data;
    set_use_cache(cache);
    double t0 = now_ms();
    if (!load_triples(argv[1], &n_rows, &n_cols, &nnz, &rows, &cols, &data)) return FileError;
    if (expand && !expand_symmetric_entries(n_rows, n_cols, last_banner_symmetry(), &nnz, &rows, &cols, &data))
        return FileError;
    double ms = now_ms() - t0;
    FILE *out = fopen(argv[2], "wb");
    if (!out) return FileError;
    fwrite(rows, sizeof(int), (size_t)nnz, out);
    fwrite(cols, sizeof(int), (size_t)nnz, out);
    fwrite(data, sizeof(double), (size_t)nnz, out);
    fclose(out);
    printf("%d %d %d %.3f\n", n_rows, n_cols, nnz, ms);
    return Success;
}
