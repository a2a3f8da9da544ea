/* mmio.c -- see mmio.h.  Behaviour follows the MatrixMarket exchange-format specification and
 * matches what the reference's callers rely on (mmio/mmio.c:96-217 in the reference): the banner's
 * five tokens (the last four case-insensitive), '%' comment lines skipped, blank lines tolerated
 * before the "M N nz" size line. */
#include "mmio.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

struct keyword {
    const char *word;
    char code;
};

static int lookup(const struct keyword *table, char *token, char *out)
{
    for (char *p = token; *p; ++p) *p = (char)tolower((unsigned char)*p);
    for (; table->word; ++table)
        if (strcmp(table->word, token) == 0) {
            *out = table->code;
            return 1;
        }
    return 0;
}

int mm_read_banner(FILE *f, MM_typecode *matcode)
{
    static const struct keyword objects[] = {{"matrix", 'M'}, {NULL, 0}};
    static const struct keyword formats[] = {{"coordinate", 'C'}, {"array", 'A'}, {NULL, 0}};
    static const struct keyword fields[] = {{"real", 'R'}, {"complex", 'C'}, {"pattern", 'P'},
                                            {"integer", 'I'}, {NULL, 0}};
    static const struct keyword symmetries[] = {{"general", 'G'}, {"symmetric", 'S'},
                                                {"hermitian", 'H'}, {"skew-symmetric", 'K'}, {NULL, 0}};
    char line[MM_MAX_LINE_LENGTH];
    char tok[5][MM_MAX_TOKEN_LENGTH];

    (*matcode)[0] = (*matcode)[1] = (*matcode)[2] = ' ';
    (*matcode)[3] = 'G';
    if (!fgets(line, sizeof line, f)) return MM_PREMATURE_EOF;
    if (sscanf(line, "%63s %63s %63s %63s %63s", tok[0], tok[1], tok[2], tok[3], tok[4]) != 5)
        return MM_PREMATURE_EOF;
    if (strncmp(tok[0], MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0) return MM_NO_HEADER;
    if (!lookup(objects, tok[1], &(*matcode)[0])) return MM_UNSUPPORTED_TYPE;
    if (!lookup(formats, tok[2], &(*matcode)[1])) return MM_UNSUPPORTED_TYPE;
    if (!lookup(fields, tok[3], &(*matcode)[2])) return MM_UNSUPPORTED_TYPE;
    if (!lookup(symmetries, tok[4], &(*matcode)[3])) return MM_UNSUPPORTED_TYPE;
    return 0;
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = *nz = 0;
    for (;;) {
        if (!fgets(line, sizeof line, f)) return MM_PREMATURE_EOF;
        if (line[0] == '%') continue;
        if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
        /* blank (or short) line: keep looking */
        const char *p = line;
        while (*p && isspace((unsigned char)*p)) ++p;
        if (*p) return MM_PREMATURE_EOF; /* non-blank garbage where the size line should be */
    }
}

char *mm_typecode_to_str(MM_typecode matcode)
{
    const char *fmt = mm_is_sparse(matcode) ? "coordinate" : mm_is_dense(matcode) ? "array" : NULL;
    const char *field = mm_is_real(matcode) ? "real" : mm_is_complex(matcode) ? "complex"
                        : mm_is_pattern(matcode) ? "pattern" : mm_is_integer(matcode) ? "integer" : NULL;
    const char *sym = mm_is_general(matcode) ? "general" : mm_is_symmetric(matcode) ? "symmetric"
                      : mm_is_hermitian(matcode) ? "hermitian" : mm_is_skew(matcode) ? "skew-symmetric" : NULL;
    if (!mm_is_matrix(matcode) || !fmt || !field || !sym) return NULL;
    char *out = (char *)malloc(MM_MAX_LINE_LENGTH);
    if (out) snprintf(out, MM_MAX_LINE_LENGTH, "%s %s %s %s", "matrix", fmt, field, sym);
    return out;
}
