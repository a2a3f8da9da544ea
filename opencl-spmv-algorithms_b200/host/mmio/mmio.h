/*
 * mmio.h -- a from-scratch, read-only subset of the NIST MatrixMarket I/O interface.
 *
 * The reference vendors NIST's mmio (mmio/mmio.c, mmio/mmio.h) but only ever calls
 * mm_read_banner, mm_read_mtx_crd_size, mm_typecode_to_str and the mm_is_* predicates
 * (inc/helper_functions.h:143-159).  This file provides exactly that surface, with the same
 * names, typecode encoding and error codes, so code written against mmio keeps compiling.  The
 * writers and the mm_read_mtx_crd* / mm_read_unsymmetric_sparse helpers are dead code in the
 * reference and are not provided.
 */
#ifndef B200_MMIO_H
#define B200_MMIO_H

#include <stdio.h>

#define MM_MAX_LINE_LENGTH 1025
#define MM_MAX_TOKEN_LENGTH 64
#define MatrixMarketBanner "%%MatrixMarket"

/* typecode[0] object 'M'; [1] 'C'oordinate | 'A'rray; [2] 'R'eal | 'C'omplex | 'P'attern |
 * 'I'nteger; [3] 'G'eneral | 'S'ymmetric | 'H'ermitian | s'K'ew-symmetric */
typedef char MM_typecode[4];

#define MM_PREMATURE_EOF 12
#define MM_NO_HEADER 14
#define MM_UNSUPPORTED_TYPE 15

#define mm_is_matrix(t) ((t)[0] == 'M')
#define mm_is_sparse(t) ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t) ((t)[1] == 'A')
#define mm_is_array(t) ((t)[1] == 'A')
#define mm_is_complex(t) ((t)[2] == 'C')
#define mm_is_real(t) ((t)[2] == 'R')
#define mm_is_pattern(t) ((t)[2] == 'P')
#define mm_is_integer(t) ((t)[2] == 'I')
#define mm_is_symmetric(t) ((t)[3] == 'S')
#define mm_is_general(t) ((t)[3] == 'G')
#define mm_is_skew(t) ((t)[3] == 'K')
#define mm_is_hermitian(t) ((t)[3] == 'H')

int mm_read_banner(FILE *f, MM_typecode *matcode);
int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz);
char *mm_typecode_to_str(MM_typecode matcode);

#endif
