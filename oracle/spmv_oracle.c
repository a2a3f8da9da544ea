/*
 * oracle/spmv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see spmv_oracle.h).
 *
 * Plain-C restatement of the reference algorithm for the SpMV hot path.  The reference
 * streams "%d %d %lg" triples from the .mtx file inside each main(); here the same state
 * machines run over in-memory triples kept in file order, which is equivalent.  Where the
 * reference has undefined behaviour the oracle defines
 * the value (zero / ORC_UNSET) and returns a status so tests can assert the input is on the
 * reference's well-defined domain: entries sorted by row, no empty rows, first row = 1.
 *
 * Parity: pinned against the unmodified reference (oracle/_ref) -- tests/test_oracle_vs_ref.py.
 */
#include "spmv_oracle.h"

#include <ctype.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * MatrixMarket header.  Follows read_size_of_matrices_from_file
 * (inc/helper_functions.h:134-165): banner via mm_read_banner (mmio/mmio.c:96-179), reject
 * only complex sparse matrices (helper_functions.h:151), then mm_read_mtx_crd_size
 * (mmio/mmio.c:189-217): skip '%' lines, read "M N nz".  Symmetry is read and IGNORED, as in
 * the reference.  Returns 0 ok, 1 cannot open, 2 bad banner, 3 unsupported, 4 bad size line.
 * ---------------------------------------------------------------------------------------- */
static void lower_inplace(char *s)
{
    for (; *s; ++s) *s = (char)tolower((unsigned char)*s);
}

static int header_from_stream(FILE *f, int *n_rows, int *n_cols, int *nnz)
{
    char line[1025];
    char tok[5][64];

    *n_rows = *n_cols = *nnz = 0;
    if (!fgets(line, sizeof line, f)) return 2;
    if (sscanf(line, "%63s %63s %63s %63s %63s", tok[0], tok[1], tok[2], tok[3], tok[4]) != 5)
        return 2;
    for (int k = 1; k < 5; ++k) lower_inplace(tok[k]);
    if (strncmp(tok[0], "%%MatrixMarket", 14) != 0) return 2;
    if (strcmp(tok[1], "matrix") != 0) return 2;
    int sparse = strcmp(tok[2], "coordinate") == 0;
    if (!sparse && strcmp(tok[2], "array") != 0) return 2;
    int is_complex = strcmp(tok[3], "complex") == 0;
    if (!is_complex && strcmp(tok[3], "real") != 0 && strcmp(tok[3], "pattern") != 0 &&
        strcmp(tok[3], "integer") != 0)
        return 2;
    if (strcmp(tok[4], "general") != 0 && strcmp(tok[4], "symmetric") != 0 &&
        strcmp(tok[4], "hermitian") != 0 && strcmp(tok[4], "skew-symmetric") != 0)
        return 2;
    if (is_complex && sparse) return 3; /* helper_functions.h:151-156 */

    do {
        if (!fgets(line, sizeof line, f)) return 4;
    } while (line[0] == '%');
    if (sscanf(line, "%d %d %d", n_rows, n_cols, nnz) == 3) return 0;
    for (;;) { /* blank line(s) before the size line: mmio.c:207-213 */
        int got = fscanf(f, "%d %d %d", n_rows, n_cols, nnz);
        if (got == EOF) return 4;
        if (got == 3) return 0;
    }
}

int orc_mtx_read_size(const char *path, int *n_rows, int *n_cols, int *nnz)
{
    FILE *f = fopen(path, "r");
    if (!f) return 1;
    int rc = header_from_stream(f, n_rows, n_cols, nnz);
    fclose(f);
    return rc;
}

/* Entry parse: the per-driver loop `fscanf(file, "%d %d %lg\n", ...)` with the 1-based ->
 * 0-based adjustment (coo.c:79-84, csr.c:77-83, cmrs.c:84-90). */
int orc_mtx_read_coo(const char *path, int nnz, int *rows, int *cols, double *vals)
{
    FILE *f = fopen(path, "r");
    int r, c, n;
    if (!f) return 1;
    int rc = header_from_stream(f, &r, &c, &n);
    if (rc) {
        fclose(f);
        return rc;
    }
    if (n != nnz) {
        fclose(f);
        return 4;
    }
    for (int i = 0; i < nnz; ++i) {
        if (fscanf(f, "%d %d %lg\n", &rows[i], &cols[i], &vals[i]) != 3) {
            fclose(f);
            return 5;
        }
        rows[i] -= 1;
        cols[i] -= 1;
    }
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CSR build -- csr.c:72-91.  ptr[0]=0, ptr[R]=nnz, and every time the row of entry i differs
 * from the previous entry's row (initially row 0) the next ptr slot receives i.  Equal to the
 * exclusive scan of row counts iff rows are sorted, none is empty and the first is row 0.
 * Returns the number of row changes seen (== n_rows-1 on the well-defined domain).
 * ---------------------------------------------------------------------------------------- */
int orc_build_csr(int n_rows, int nnz, const int *rows, int *ptr)
{
    for (int k = 0; k <= n_rows; ++k) ptr[k] = ORC_UNSET;
    ptr[0] = 0;
    ptr[n_rows] = nnz;
    int slot = 1, last = 0, changes = 0;
    for (int i = 0; i < nnz; ++i) {
        if (rows[i] != last) {
            if (slot < n_rows) ptr[slot] = i; /* reference writes unguarded (UB past R) */
            ++slot;
            ++changes;
            last = rows[i];
        }
    }
    return changes;
}

/* ------------------------------------------------------------------------------------------
 * ELL statistics pass -- ell.c:68-101.  A row's length is only folded into longest/shortest/
 * sum when the NEXT row starts, so the last row is never counted (quirk q3).  sum is `int`.
 * ---------------------------------------------------------------------------------------- */
int orc_ell_stats(int n_rows, int nnz, const int *rows, int *longest, int *shortest, int *sum_len)
{
    int hi = 0, lo = INT_MAX, total = 0, run = 0, last = 0;
    (void)n_rows;
    for (int i = 0; i < nnz; ++i) {
        if (rows[i] == last) {
            ++run;
        } else {
            last = rows[i];
            if (run > hi) hi = run;
            if (run < lo) lo = run;
            total += run;
            run = 1;
        }
    }
    *longest = hi;
    *shortest = lo;
    *sum_len = total;
    return run; /* length of the (uncounted) last row */
}

/* ------------------------------------------------------------------------------------------
 * ELL fill pass -- ell.c:118-164.  Row-major, `row_size` slots per row.  On a row change the
 * previous row is padded with column 0 up to row_size*diff slots (diff = row gap), padding
 * DATA is never written by the reference (malloc, ell.c:119): the oracle defines it as 0.0.
 * Returns 0, or 1 if the reference would have run past its allocation (off-domain input).
 * ---------------------------------------------------------------------------------------- */
int orc_build_ell(int n_rows, int nnz, int row_size, const int *rows, const int *cols,
                  const double *vals, int *ell_cols, double *ell_data)
{
    long cap = (long)n_rows * (long)row_size;
    long at = 0;
    int in_row = 0, last = 0;
    for (long k = 0; k < cap; ++k) {
        ell_cols[k] = ORC_UNSET;
        ell_data[k] = 0.0;
    }
    for (int i = 0; i < nnz; ++i) {
        if (rows[i] == last) {
            if (at >= cap) return 1;
            ell_data[at] = vals[i];
            ell_cols[at] = cols[i];
            ++at;
            ++in_row;
        } else {
            int gap = rows[i] - last;
            last = rows[i];
            for (long k = in_row; k < (long)row_size * (long)gap; ++k) {
                if (at >= cap) return 1;
                ell_cols[at++] = 0;
            }
            in_row = 1;
            if (at >= cap) return 1;
            ell_cols[at] = cols[i];
            ell_data[at] = vals[i];
            ++at;
        }
    }
    for (int k = in_row; k < row_size; ++k) {
        if (at >= cap) return 1;
        ell_cols[at++] = 0;
    }
    return 0;
}

void orc_ell_to_colmajor(int n_rows, int row_size, int pitch, const int *ell_cols,
                         const double *ell_data, int *cm_cols, double *cm_data)
{
    for (long k = 0; k < (long)pitch * row_size; ++k) {
        cm_cols[k] = 0;
        cm_data[k] = 0.0;
    }
    for (int r = 0; r < n_rows; ++r)
        for (int k = 0; k < row_size; ++k) {
            cm_cols[(long)k * pitch + r] = ell_cols[(long)r * row_size + k];
            cm_data[(long)k * pitch + r] = ell_data[(long)r * row_size + k];
        }
}

/* ------------------------------------------------------------------------------------------
 * SELL-C build (C = chunk = 32 in the reference, no sigma, no permutation).
 * Slice count: sigma_c.c:74-81.  Pointer pass: sigma_c.c:87-139 -- a slice is closed when the
 * 32nd row CHANGE is seen; its width is the longest completed row since the last close, and
 * after a close the running maximum restarts at 1 (sigma_c.c:120), not 0.  The final slice is
 * always closed after the loop and always spans `chunk` rows (sigma_c.c:130-139).
 * Returns elements_sum (the allocation size the reference callocs, sigma_c.c:153-154).
 * ---------------------------------------------------------------------------------------- */
int orc_sell_num_slices(int n_rows, int chunk)
{
    return n_rows % chunk == 0 ? n_rows / chunk : n_rows / chunk + 1;
}

long orc_build_sell_ptr(int n_rows, int nnz, int chunk, const int *rows, int *row_indices)
{
    int n_slices = orc_sell_num_slices(n_rows, chunk);
    int widest = 0, last = 0, run = 0, seen = 0, slot = 0;
    long total = 0;
    for (int k = 0; k <= n_slices; ++k) row_indices[k] = ORC_UNSET;
    row_indices[0] = 0;
    for (int i = 0; i < nnz; ++i) {
        if (rows[i] == last) {
            ++run;
            continue;
        }
        ++seen;
        if (run > widest) widest = run;
        if (seen == chunk) {
            total += widest * chunk; /* int*int product, as in the reference */
            if (slot + 1 <= n_slices) row_indices[slot + 1] = (int)total;
            widest = 1;
            ++slot;
            seen = 0;
        }
        run = 1;
        last = rows[i];
    }
    if (seen != chunk) { /* always true: `seen` is reset on every close */
        if (run > widest) widest = run;
        total += widest * chunk;
        row_indices[n_slices] = (int)total;
    }
    return total;
}

/* Fill pass: sigma_c.c:156-202.  Row r of slice s starts at row_indices[s] + r and steps by
 * `chunk`; arrays are calloc'ed so padding is (col 0, 0.0). */
int orc_build_sell_fill(int n_rows, int nnz, int chunk, const int *rows, const int *cols,
                        const double *vals, const int *row_indices, int *sell_cols,
                        double *sell_data)
{
    int n_slices = orc_sell_num_slices(n_rows, chunk);
    long cap = row_indices[n_slices];
    int last = 0, seen = 0, slice = 0;
    long row_start = row_indices[0], at = row_start;
    for (long k = 0; k < cap; ++k) {
        sell_cols[k] = 0;
        sell_data[k] = 0.0;
    }
    for (int i = 0; i < nnz; ++i) {
        if (rows[i] != last) {
            ++seen;
            if (seen == chunk) {
                seen = 0;
                ++slice;
                if (slice > n_slices) return 1;
                row_start = row_indices[slice];
            } else {
                ++row_start;
            }
            at = row_start;
            last = rows[i];
        }
        if (at < 0 || at >= cap) return 1;
        sell_data[at] = vals[i];
        sell_cols[at] = cols[i];
        at += chunk;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CMRS build (height = 8 in the reference) -- cmrs.c:72-117.  strip_ptr has ceil(R/h)+1
 * entries; a new strip starts at the row change that follows in-strip row h-1.
 * ---------------------------------------------------------------------------------------- */
int orc_cmrs_num_strips(int n_rows, int height)
{
    return (int)ceil((double)n_rows / (double)height); /* cmrs.c:72 */
}

int orc_build_cmrs(int n_rows, int nnz, int height, const int *rows, int *strip_ptr,
                   int *row_in_strip)
{
    int n_strips = orc_cmrs_num_strips(n_rows, height);
    int last = 0, local = 0, slot = 1;
    for (int k = 0; k <= n_strips; ++k) strip_ptr[k] = ORC_UNSET;
    strip_ptr[0] = 0;
    for (int i = 0; i < nnz; ++i) {
        if (rows[i] != last) {
            last = rows[i];
            if (local == height - 1) {
                if (slot <= n_strips) strip_ptr[slot] = i;
                ++slot;
                local = 0;
            } else {
                ++local;
            }
        }
        row_in_strip[i] = local;
    }
    strip_ptr[n_strips] = nnz;
    return slot - 1; /* strips opened after the first (== n_strips-1 on-domain) */
}

/* ------------------------------------------------------------------------------------------
 * SELL-C-sigma: NEW capability (the reference's sigma_c.c has no sigma and no permutation,
 * SURVEY.md section 0.4).  Specification: rows are grouped in windows of `sigma` consecutive
 * rows; inside a window rows are stably sorted by DESCENDING length (ties keep ascending
 * original index); perm[new] = old.  Slices are then cut over the permuted order exactly as
 * in the reference: width = max(1, longest row in slice), every slice spans `chunk` rows.
 * sigma <= 1 gives the identity permutation and therefore the reference layout.
 * Requires rows sorted ascending (empty rows allowed here: length 0).
 * ---------------------------------------------------------------------------------------- */
static int *row_lengths(int n_rows, int nnz, const int *rows)
{
    int *len = (int *)calloc((size_t)n_rows > 0 ? (size_t)n_rows : 1, sizeof(int));
    for (int i = 0; i < nnz; ++i) len[rows[i]]++;
    return len;
}

int orc_sell_sigma_perm(int n_rows, int nnz, int sigma, const int *rows, int *perm)
{
    int *len = row_lengths(n_rows, nnz, rows);
    for (int r = 0; r < n_rows; ++r) perm[r] = r;
    if (sigma > 1) {
        for (int w0 = 0; w0 < n_rows; w0 += sigma) {
            int w1 = w0 + sigma < n_rows ? w0 + sigma : n_rows;
            /* stable insertion-free approach: counting sort by length, descending */
            int maxlen = 0;
            for (int r = w0; r < w1; ++r)
                if (len[r] > maxlen) maxlen = len[r];
            int *start = (int *)calloc((size_t)maxlen + 2, sizeof(int));
            for (int r = w0; r < w1; ++r) start[maxlen - len[r] + 1]++;
            for (int k = 1; k <= maxlen + 1; ++k) start[k] += start[k - 1];
            for (int r = w0; r < w1; ++r) perm[w0 + start[maxlen - len[r]]++] = r;
            free(start);
        }
    }
    free(len);
    return 0;
}

long orc_build_sell_sigma(int n_rows, int nnz, int chunk, int sigma, const int *rows,
                          const int *cols, const double *vals, int *perm, long long *slice_ptr,
                          int *sell_cols, double *sell_data, long capacity)
{
    int n_slices = orc_sell_num_slices(n_rows, chunk);
    int *len = row_lengths(n_rows, nnz, rows);
    int *first = (int *)malloc(((size_t)n_rows + 1) * sizeof(int));
    first[0] = 0;
    for (int r = 0; r < n_rows; ++r) first[r + 1] = first[r] + len[r];
    orc_sell_sigma_perm(n_rows, nnz, sigma, rows, perm);

    slice_ptr[0] = 0;
    for (int s = 0; s < n_slices; ++s) {
        int width = 1;
        for (int j = 0; j < chunk; ++j) {
            int nr = s * chunk + j;
            if (nr < n_rows && len[perm[nr]] > width) width = len[perm[nr]];
        }
        slice_ptr[s + 1] = slice_ptr[s] + (long long)width * chunk;
    }
    long total = (long)slice_ptr[n_slices];
    if (sell_cols && sell_data) {
        if (total > capacity) {
            free(len);
            free(first);
            return -total;
        }
        for (long k = 0; k < total; ++k) {
            sell_cols[k] = 0;
            sell_data[k] = 0.0;
        }
        for (int nr = 0; nr < n_rows; ++nr) {
            int old = perm[nr];
            long at = (long)slice_ptr[nr / chunk] + nr % chunk;
            for (int k = 0; k < len[old]; ++k, at += chunk) {
                sell_cols[at] = cols[first[old] + k];
                sell_data[at] = vals[first[old] + k];
            }
        }
    }
    free(len);
    free(first);
    return total;
}

/* ------------------------------------------------------------------------------------------
 * SpMV.  The reference's CPU paths accumulate into a malloc'ed, never-zeroed buffer (quirk
 * q2); the oracle zeroes y first.  Loop shapes follow the reference:
 *   y_ref : check_result, inc/helper_functions.h:207-219 (serial fp64 COO, file order)
 *   COO   : coo.c:288-293 (omp parallel for + omp atomic)
 *   CSR   : csr.c:293-302        ELL : ell.c:365-376 (all row_size slots incl. padding)
 *   CMRS  : cmrs.c:327-338       SELL: kernels/Sigma_C.cl:3-17 (no CPU path in sigma_c.c)
 * ---------------------------------------------------------------------------------------- */
void orc_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_yref_coo_serial(int n_rows, int nnz, const int *rows, const int *cols, const double *vals,
                         const double *x, double *y)
{
    for (int r = 0; r < n_rows; ++r) y[r] = 0.0;
    for (int i = 0; i < nnz; ++i) y[rows[i]] += vals[i] * x[cols[i]];
}

#define ORC_DEFINE_SPMV(T, SUF)                                                                   \
    void orc_spmv_coo_##SUF(int n_rows, int nnz, const int *rows, const int *cols,                \
                            const T *vals, const T *x, T *y)                                      \
    {                                                                                             \
        for (int r = 0; r < n_rows; ++r) y[r] = 0;                                                \
        _Pragma("omp parallel for") for (int i = 0; i < nnz; ++i)                                 \
        {                                                                                         \
            T prod = vals[i] * x[cols[i]];                                                        \
            _Pragma("omp atomic") y[rows[i]] += prod;                                             \
        }                                                                                         \
    }                                                                                             \
    void orc_spmv_csr_##SUF(int n_rows, const int *ptr, const int *cols, const T *vals,           \
                            const T *x, T *y)                                                     \
    {                                                                                             \
        _Pragma("omp parallel for") for (int r = 0; r < n_rows; ++r)                              \
        {                                                                                         \
            T acc = 0;                                                                            \
            for (int j = ptr[r]; j < ptr[r + 1]; ++j) acc += vals[j] * x[cols[j]];                \
            y[r] = acc;                                                                           \
        }                                                                                         \
    }                                                                                             \
    void orc_spmv_ell_##SUF(int n_rows, int row_size, const int *cols, const T *vals, const T *x, \
                            T *y)                                                                 \
    {                                                                                             \
        _Pragma("omp parallel for") for (int r = 0; r < n_rows; ++r)                              \
        {                                                                                         \
            long base = (long)r * row_size;                                                       \
            T acc = 0;                                                                            \
            for (int k = 0; k < row_size; ++k) acc += vals[base + k] * x[cols[base + k]];         \
            y[r] = acc;                                                                           \
        }                                                                                         \
    }                                                                                             \
    void orc_spmv_sell_##SUF(int n_slices, int chunk, const int *row_indices, const int *cols,    \
                             const T *vals, const T *x, T *y_padded)                              \
    {                                                                                             \
        _Pragma("omp parallel for") for (int s = 0; s < n_slices; ++s)                            \
        {                                                                                         \
            for (int lane = 0; lane < chunk; ++lane) {                                            \
                T acc = 0;                                                                        \
                for (long j = (long)row_indices[s] + lane; j < row_indices[s + 1]; j += chunk)    \
                    acc += vals[j] * x[cols[j]];                                                  \
                y_padded[(long)s * chunk + lane] = acc;                                           \
            }                                                                                     \
        }                                                                                         \
    }                                                                                             \
    void orc_spmv_sell64_##SUF(int n_rows, int n_slices, int chunk, const long long *slice_ptr,   \
                               const int *perm, const int *cols, const T *vals, const T *x, T *y) \
    {                                                                                             \
        _Pragma("omp parallel for") for (int s = 0; s < n_slices; ++s)                            \
        {                                                                                         \
            for (int lane = 0; lane < chunk; ++lane) {                                            \
                long nr = (long)s * chunk + lane;                                                 \
                if (nr >= n_rows) continue;                                                       \
                T acc = 0;                                                                        \
                for (long long j = slice_ptr[s] + lane; j < slice_ptr[s + 1]; j += chunk)         \
                    acc += vals[j] * x[cols[j]];                                                  \
                y[perm ? perm[nr] : nr] = acc;                                                    \
            }                                                                                     \
        }                                                                                         \
    }                                                                                             \
    void orc_spmv_cmrs_##SUF(int n_rows, int n_strips, int height, const int *strip_ptr,          \
                             const int *row_in_strip, const int *cols, const T *vals, const T *x, \
                             T *y)                                                                \
    {                                                                                             \
        for (int r = 0; r < n_rows; ++r) y[r] = 0;                                                \
        _Pragma("omp parallel for") for (int t = 0; t < n_strips; ++t)                            \
        {                                                                                         \
            long row0 = (long)t * height;                                                         \
            for (int j = strip_ptr[t]; j < strip_ptr[t + 1]; ++j)                                 \
                y[row0 + row_in_strip[j]] += vals[j] * x[cols[j]];                                \
        }                                                                                         \
    }

ORC_DEFINE_SPMV(double, f64)
ORC_DEFINE_SPMV(float, f32)

/* ------------------------------------------------------------------------------------------
 * Comparators.  BASELINE.json: relative max-norm ||y - y_ref||_inf / ||y_ref||_inf <= 1e-5
 * (fp32) / 1e-12 (fp64).  orc_check_abs is the reference's own criterion (absolute, EPSILON
 * 1e-6, inc/helper_functions.h:11,221-230); returns the first failing index or -1.
 * ---------------------------------------------------------------------------------------- */
double orc_rel_maxnorm_f64(int n, const double *y, const double *y_ref)
{
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; ++i) {
        double d = fabs(y[i] - y_ref[i]);
        if (d != d) return NAN; /* a NaN anywhere fails every tolerance */
        if (d > num) num = d;
        if (fabs(y_ref[i]) > den) den = fabs(y_ref[i]);
    }
    return den > 0.0 ? num / den : num;
}

double orc_rel_maxnorm_f32(int n, const float *y, const double *y_ref)
{
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; ++i) {
        double d = fabs((double)y[i] - y_ref[i]);
        if (d != d) return NAN;
        if (d > num) num = d;
        if (fabs(y_ref[i]) > den) den = fabs(y_ref[i]);
    }
    return den > 0.0 ? num / den : num;
}

int orc_check_abs(int n, const double *y, const double *y_ref, double eps)
{
    for (int i = 0; i < n; ++i)
        if (fabs(y_ref[i] - y[i]) > eps) return i;
    return -1;
}
