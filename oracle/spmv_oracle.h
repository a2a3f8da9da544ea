/*
 * oracle/spmv_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's (sgartkink/opencl-spmv-algorithms) SpMV hot
 * path: MatrixMarket load, the five format builds and the per-format CPU SpMV.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (libb200spmv.so and the
 * bin/<fmt> drivers) never links or calls it.
 *
 * Parity status: PINNED.  Every builder here is checked bit-for-bit against the
 * arrays the UNMODIFIED reference drivers upload (oracle/_ref, built from
 * /root/reference by oracle/Makefile against the recording fake OpenCL runtime in
 * oracle/fake_cl/), and every SpMV against the reference's own compute_using_cpu /
 * check_result symbols (tests/test_oracle_vs_ref.py, tests/golden/).
 *
 * All "file:line" citations are relative to the reference repository root.
 */
#ifndef SPMV_ORACLE_H
#define SPMV_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_UNSET (-1) /* marks slots the reference would leave uninitialised */

/* ---- MatrixMarket load (inc/helper_functions.h:134-165, mmio/mmio.c:96-217) ---- */
int orc_mtx_read_size(const char *path, int *n_rows, int *n_cols, int *nnz);
int orc_mtx_read_coo(const char *path, int nnz, int *rows, int *cols, double *vals);

/* ---- format builds; inputs are 0-based triples in FILE ORDER ---- */
int orc_build_csr(int n_rows, int nnz, const int *rows, int *ptr);
int orc_ell_stats(int n_rows, int nnz, const int *rows, int *longest, int *shortest, int *sum_len);
int orc_build_ell(int n_rows, int nnz, int row_size, const int *rows, const int *cols,
                  const double *vals, int *ell_cols, double *ell_data);
int orc_sell_num_slices(int n_rows, int chunk);
long orc_build_sell_ptr(int n_rows, int nnz, int chunk, const int *rows, int *row_indices);
int orc_build_sell_fill(int n_rows, int nnz, int chunk, const int *rows, const int *cols,
                        const double *vals, const int *row_indices, int *sell_cols,
                        double *sell_data);
int orc_cmrs_num_strips(int n_rows, int height);
int orc_build_cmrs(int n_rows, int nnz, int height, const int *rows, int *strip_ptr,
                   int *row_in_strip);

/* SELL-C-sigma (new capability, no reference code; reduces to the reference layout at sigma=1) */
int orc_sell_sigma_perm(int n_rows, int nnz, int sigma, const int *rows, int *perm);
long orc_build_sell_sigma(int n_rows, int nnz, int chunk, int sigma, const int *rows,
                          const int *cols, const double *vals, int *perm, long long *slice_ptr,
                          int *sell_cols, double *sell_data, long capacity);

/* ELL column-major device layout = transpose of the row-major build, pitch-padded */
void orc_ell_to_colmajor(int n_rows, int row_size, int pitch, const int *ell_cols,
                         const double *ell_data, int *cm_cols, double *cm_data);

/* ---- SpMV ---- */
void orc_set_threads(int n);
int orc_get_max_threads(void);

void orc_yref_coo_serial(int n_rows, int nnz, const int *rows, const int *cols, const double *vals,
                         const double *x, double *y);

void orc_spmv_coo_f64(int n_rows, int nnz, const int *rows, const int *cols, const double *vals,
                      const double *x, double *y);
void orc_spmv_coo_f32(int n_rows, int nnz, const int *rows, const int *cols, const float *vals,
                      const float *x, float *y);
void orc_spmv_csr_f64(int n_rows, const int *ptr, const int *cols, const double *vals,
                      const double *x, double *y);
void orc_spmv_csr_f32(int n_rows, const int *ptr, const int *cols, const float *vals,
                      const float *x, float *y);
void orc_spmv_ell_f64(int n_rows, int row_size, const int *cols, const double *vals,
                      const double *x, double *y);
void orc_spmv_ell_f32(int n_rows, int row_size, const int *cols, const float *vals, const float *x,
                      float *y);
void orc_spmv_sell_f64(int n_slices, int chunk, const int *row_indices, const int *cols,
                       const double *vals, const double *x, double *y_padded);
void orc_spmv_sell_f32(int n_slices, int chunk, const int *row_indices, const int *cols,
                       const float *vals, const float *x, float *y_padded);
void orc_spmv_sell64_f64(int n_rows, int n_slices, int chunk, const long long *slice_ptr,
                         const int *perm, const int *cols, const double *vals, const double *x,
                         double *y);
void orc_spmv_sell64_f32(int n_rows, int n_slices, int chunk, const long long *slice_ptr,
                         const int *perm, const int *cols, const float *vals, const float *x,
                         float *y);
void orc_spmv_cmrs_f64(int n_rows, int n_strips, int height, const int *strip_ptr,
                       const int *row_in_strip, const int *cols, const double *vals,
                       const double *x, double *y);
void orc_spmv_cmrs_f32(int n_rows, int n_strips, int height, const int *strip_ptr,
                       const int *row_in_strip, const int *cols, const float *vals, const float *x,
                       float *y);

/* ---- comparators ---- */
double orc_rel_maxnorm_f64(int n, const double *y, const double *y_ref);
double orc_rel_maxnorm_f32(int n, const float *y, const double *y_ref);
int orc_check_abs(int n, const double *y, const double *y_ref, double eps);

#ifdef __cplusplus
}
#endif
#endif
