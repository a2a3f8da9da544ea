/* driver_common.c -- definitions of the helpers declared in inc/helper_functions.h. */
#define _POSIX_C_SOURCE 200809L
#include "helper_functions.h"

#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_value_bytes = 8;
static double g_rel_tolerance = 1e-12;
static int g_expand_symmetric = 0;   /* --expand-symmetric */
static int g_use_cache = 0;          /* --cache */
static int g_banner_symmetry = 0;    /* of the last header read: 0 general, 1 symmetric/hermitian, -1 skew */

int driver_parse_args(int argc, char **argv, const char *default_matrix, driver_options *opt)
{
    opt->matrix = default_matrix;
    opt->use_f32 = 0;
    opt->sigma = 1;
    opt->reps = 1;
    opt->device = 0;
    opt->no_cpu = 0;
    opt->rowmajor = 1;
    opt->expand_symmetric = 0;
    opt->cache = 0;
    opt->gpus = 1;
    opt->iters = 0;
    opt->json = 0;
    opt->synthetic = NULL;
    opt->sync_mcast = 0;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        const char *v = i + 1 < argc ? argv[i + 1] : NULL;
        if (!strcmp(a, "--matrix") && v) opt->matrix = argv[++i];
        else if (!strcmp(a, "--dtype") && v) {
            ++i;
            if (!strcmp(v, "f32")) opt->use_f32 = 1;
            else if (!strcmp(v, "f64")) opt->use_f32 = 0;
            else goto bad;
        } else if (!strcmp(a, "--sigma") && v) opt->sigma = atoi(argv[++i]);
        else if (!strcmp(a, "--reps") && v) opt->reps = atoi(argv[++i]);
        else if (!strcmp(a, "--device") && v) opt->device = atoi(argv[++i]);
        else if (!strcmp(a, "--no-cpu")) opt->no_cpu = 1;
        else if (!strcmp(a, "--rowmajor")) opt->rowmajor = 1;
        else if (!strcmp(a, "--colmajor")) opt->rowmajor = 0;
        else if (!strcmp(a, "--expand-symmetric")) opt->expand_symmetric = 1;
        else if (!strcmp(a, "--cache")) opt->cache = 1;
        else if (!strcmp(a, "--gpus") && v) opt->gpus = atoi(argv[++i]);
        else if (!strcmp(a, "--iters") && v) opt->iters = atoi(argv[++i]);
        else if (!strcmp(a, "--json")) opt->json = 1;
        else if (!strcmp(a, "--synthetic") && v) opt->synthetic = argv[++i];
        else if (!strcmp(a, "--sync") && v) {
            ++i;
            if (!strcmp(v, "mcast")) opt->sync_mcast = 1;
            else if (!strcmp(v, "nccl")) opt->sync_mcast = 0;
            else goto bad;
        }
        else goto bad;
    }
    if (opt->reps < 1 || opt->sigma < 1 || opt->device < 0 || opt->gpus < 1 || opt->gpus > DEVICES_DEFAULT_SIZE ||
        opt->iters < 0)
        goto bad;
    if ((opt->gpus > 1 || opt->synthetic) && opt->iters == 0) {
        fprintf(stderr, "--gpus / --synthetic belong to the iterated mode: give --iters K as well\n");
        goto bad;
    }
    set_value_bytes(opt->use_f32 ? 4 : 8);
    set_check_tolerance(opt->use_f32 ? 1e-5 : 1e-12);
    set_expand_symmetric(opt->expand_symmetric);
    set_use_cache(opt->cache);
    return 0;
bad:
    fprintf(stderr, "usage: %s [--matrix FILE.mtx] [--dtype f32|f64] [--sigma N] [--reps N] "
                    "[--device D] [--no-cpu] [--rowmajor|--colmajor] [--expand-symmetric] [--cache]\n"
                    "       iterated mode (csr, sigma_c): --iters K [--gpus N] [--synthetic laplace7:NXxNYxNZ] [--sync nccl|mcast] [--json]\n", argv[0]);
    return 1;
}

bool read_size_of_matrices_from_file(FILE *file, int *number_of_rows, int *number_of_columns,
                                     int *number_of_nonzeroes)
{
    MM_typecode matcode;

    if (file == NULL) return false;
    if (mm_read_banner(file, &matcode) != 0) {
        printf("Could not process Matrix Market banner.\n");
        return false;
    }
    /* only complex matrices are refused; symmetry is read and ignored, as in the reference */
    if (mm_is_complex(matcode) && mm_is_matrix(matcode) && mm_is_sparse(matcode)) {
        char *name = mm_typecode_to_str(matcode);
        printf("Sorry, this application does not support ");
        printf("Market Market type: [%s]\n", name ? name : "?");
        free(name);
        return false;
    }
    g_banner_symmetry = mm_is_skew(matcode) ? -1 : (mm_is_symmetric(matcode) || mm_is_hermitian(matcode)) ? 1 : 0;
    return mm_read_mtx_crd_size(file, number_of_rows, number_of_columns, number_of_nonzeroes) == 0;
}

/* ---- number scanning --------------------------------------------------------
 * strtol / strtod are what fscanf("%d %d %lg") runs underneath, and they are ~all of a load: 240 ns per
 * entry on one core.  The two scanners below take the common spellings themselves and hand everything
 * else to strtol / strtod, so the values stay bit-identical to the reference's reading:
 *   fast_int     [ws][+-]digits, at most 18 digits;
 *   fast_double  [ws][+-]digits[.digits][(e|E)[+-]digits] with at most 19 significant digits in total.  When
 *                the decimal mantissa fits 53 bits and the decimal exponent lies in [-22, 22] the value is
 *                ONE IEEE operation on two exactly representable numbers (mantissa * or / 10^k) and hence
 *                correctly rounded -- the same double glibc's correctly rounded strtod returns (Clinger's
 *                fast path).  Longer mantissas, larger exponents, hex floats, inf / nan: strtod. */
static inline char *skip_space(char *p)
{
    while (*p == ' ' || (*p >= '\t' && *p <= '\r')) ++p;
    return p;
}

static inline char *fast_int(char *p, long *out)
{
    char *t = skip_space(p);
    const bool neg = *t == '-';
    if (*t == '-' || *t == '+') ++t;
    unsigned long v = 0;
    int digits = 0;
    while ((unsigned)(*t - '0') <= 9u && digits < 19) {
        v = v * 10u + (unsigned long)(*t - '0');
        ++t;
        ++digits;
    }
    if (digits == 0 || digits > 18) {  /* nothing there, or too long for the plain loop: let strtol decide */
        char *q;
        *out = strtol(p, &q, 10);
        return q == p ? NULL : q;
    }
    *out = neg ? -(long)v : (long)v;
    return t;
}

static inline char *fast_double(char *p, double *out)
{
    static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                     1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    char *t = skip_space(p);
    const bool neg = *t == '-';
    if (*t == '-' || *t == '+') ++t;
    char *const first = t;
    unsigned long long mant = 0;
    int sig = 0;      /* significant digits taken into mant (leading zeros do not count) */
    int exp10 = 0;
    bool any = false, fast = true;
    while ((unsigned)(*t - '0') <= 9u) {
        any = true;
        if (sig < 19) {
            mant = mant * 10u + (unsigned)(*t - '0');
            sig += (mant != 0);
        } else {
            fast = false;
        }
        ++t;
    }
    if (*t == '.') {
        ++t;
        while ((unsigned)(*t - '0') <= 9u) {
            any = true;
            if (sig < 19) {
                mant = mant * 10u + (unsigned)(*t - '0');
                sig += (mant != 0);
                --exp10;
            } else {
                fast = false;
            }
            ++t;
        }
    }
    /* "0x..." is a hex float for strtod; no digits at all may still be inf / nan (or garbage) */
    if (!any || (first[0] == '0' && (first[1] == 'x' || first[1] == 'X'))) fast = false;
    if (fast && (*t == 'e' || *t == 'E')) {
        char *e = t + 1;
        const bool eneg = *e == '-';
        if (*e == '-' || *e == '+') ++e;
        if ((unsigned)(*e - '0') <= 9u) {  /* an exponent needs a digit, else the 'e' is not part of the number */
            int ev = 0;
            while ((unsigned)(*e - '0') <= 9u) {
                if (ev < 10000) ev = ev * 10 + (*e - '0');
                ++e;
            }
            exp10 += eneg ? -ev : ev;
            t = e;
        }
    }
    if (fast && mant <= (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
        double v = (double)mant;
        if (exp10 < 0) v /= kPow10[-exp10];
        else v *= kPow10[exp10];
        *out = neg ? -v : v;
        return t;
    }
    char *q;
    *out = strtod(p, &q);
    return q == p ? NULL : q;
}

/* Parses `count` "%d %d %lg" triples starting at p into slots [first, first+count); returns the
 * position after the last one, NULL on malformed input. */
static char *parse_entries_serial(char *p, int first, int count, int *rows, int *cols, double *data)
{
    for (int i = first; i < first + count; ++i) {
        long r, c;
        double v;
        p = fast_int(p, &r);
        if (!p) return NULL;
        p = fast_int(p, &c);
        if (!p) return NULL;
        p = fast_double(p, &v);
        if (!p) return NULL;
        rows[i] = (int)r - 1; /* adjust from 1-based to 0-based */
        cols[i] = (int)c - 1;
        data[i] = v;
    }
    return p;
}

/* The same parse on all host cores (the text parse is > 99 % of the reference's
 * wall-clock).  The buffer is cut into chunks at line ends, non-blank lines are counted per chunk,
 * an exclusive scan gives every chunk its first entry slot, then the chunks are parsed
 * independently.  Valid only when every entry sits on its own line (what MatrixMarket files look
 * like); returns false -- and the caller falls back to the serial parse, which like fscanf is
 * indifferent to line breaks -- if the line count does not match. */
static bool parse_entries_parallel(char *buf, size_t len, int nnz, int *rows, int *cols, double *data)
{
    enum { MAX_CHUNKS = 256 };
    int n_chunks = 1;
#ifdef _OPENMP
    n_chunks = omp_get_max_threads() * 4;
#endif
    if (n_chunks > MAX_CHUNKS) n_chunks = MAX_CHUNKS;
    if (len < (size_t)n_chunks * 4096 || nnz < n_chunks * 64) return false; /* small file: serial is fine */
    size_t cut[MAX_CHUNKS + 1];
    long lines[MAX_CHUNKS + 1];
    cut[0] = 0;
    for (int k = 1; k < n_chunks; ++k) {
        size_t at = len / (size_t)n_chunks * (size_t)k;
        if (at < cut[k - 1]) at = cut[k - 1];
        char *nl = (char *)memchr(buf + at, '\n', len - at);
        cut[k] = nl ? (size_t)(nl - buf) + 1 : len;
    }
    cut[n_chunks] = len;
#pragma omp parallel for schedule(static, 1)
    for (int k = 0; k < n_chunks; ++k) {
        /* non-blank lines of the chunk: memchr finds the line ends (vectorised in libc), and whether a line
         * is blank is decided by its first few characters; a last line without a newline counts too */
        long count = 0;
        size_t i = cut[k];
        const size_t end = cut[k + 1];
        while (i < end) {
            const char *nl = (const char *)memchr(buf + i, '\n', end - i);
            const size_t e = nl ? (size_t)(nl - buf) : end;
            size_t j = i;
            while (j < e && (buf[j] == ' ' || buf[j] == '\t' || buf[j] == '\r')) ++j;
            count += j < e;
            i = e + 1;
        }
        lines[k] = count;
    }
    long total = 0;
    for (int k = 0; k < n_chunks; ++k) {
        const long c = lines[k];
        lines[k] = total;
        total += c;
    }
    lines[n_chunks] = total;
    if (total < nnz) return false; /* entries span lines: not the simple layout */
    int failed = 0;
#pragma omp parallel for schedule(static, 1)
    for (int k = 0; k < n_chunks; ++k) {
        long first = lines[k], count = lines[k + 1] - lines[k];
        if (first >= nnz) continue;
        const bool clipped = first + count > nnz;
        if (clipped) count = nnz - first; /* trailing lines beyond nnz are ignored, as fscanf would */
        char *stop = parse_entries_serial(buf + cut[k], (int)first, (int)count, rows, cols, data);
        bool bad = stop == NULL;
        if (!bad && !clipped) {
            /* one entry per line means the chunk is consumed exactly: anything else (an entry
             * wrapped over two lines, two entries on one line) sends the caller to the serial parse */
            if (stop > buf + cut[k + 1]) bad = true;
            for (char *t = stop; !bad && t < buf + cut[k + 1]; ++t)
                if (*t != ' ' && *t != '\t' && *t != '\r' && *t != '\n') bad = true;
        }
        if (bad) {
#pragma omp atomic write
            failed = 1;
        }
    }
    return !failed;
}

/* Entries are parsed from one in-memory copy of the rest of the file with strtol/strtod, which is
 * what fscanf("%d %d %lg") does underneath, minus the per-call stdio overhead (the text parse is
 * ~all of the reference's wall-clock). */
bool read_entries(FILE *file, int number_of_nonzeroes, int *rows, int *cols, double *data)
{
    long here = ftell(file);
    if (here < 0 || fseek(file, 0, SEEK_END) != 0) return false;
    long end = ftell(file);
    if (end < here || fseek(file, here, SEEK_SET) != 0) return false;
    size_t len = (size_t)(end - here);
    const double t_read = now_ms();
    /* The text is parsed in place from a read-only mapping of the file (no 90 MB copy out of the page cache:
     * 63 ms of a 125 ms load of the cant-sized file).  The scanners stop at the first byte that cannot belong
     * to a number, so the text needs a terminator: the bytes of the last page beyond the end of the file are
     * zero -- unless the size is a multiple of the page size, or this is not a mappable file: then the old
     * way, a copy with a NUL behind it. */
    char *buf = NULL, *map = NULL;
    const long page = sysconf(_SC_PAGESIZE);
    if (page > 0 && end > 0 && end % page != 0) {
        void *m = mmap(NULL, (size_t)end, PROT_READ, MAP_PRIVATE, fileno(file), 0);
        if (m != MAP_FAILED) {
            map = (char *)m;
            (void)posix_madvise(map, (size_t)end, POSIX_MADV_WILLNEED);
            buf = map + here;
        }
    }
    if (!map) {
        buf = (char *)malloc(len + 1);
        if (!buf) return false;
        if (fread(buf, 1, len, file) != len) {
            free(buf);
            return false;
        }
        buf[len] = '\0';
    }
    const double t_parse = now_ms();
    bool ok = parse_entries_parallel(buf, len, number_of_nonzeroes, rows, cols, data);
    if (!ok) ok = parse_entries_serial(buf, 0, number_of_nonzeroes, rows, cols, data) != NULL;
    if (getenv("B200_PARSE_TIMING")) /* where a load spends its time (stderr; never in the drivers' stdout) */
        fprintf(stderr, "read_entries: %zu bytes %s in %.1f ms, %d entries parsed in %.1f ms\n", len, map ? "mapped" : "read",
                t_parse - t_read, number_of_nonzeroes, now_ms() - t_parse);
    if (map) munmap(map, (size_t)end);
    else free(buf);
    return ok;
}

/* ---- optional symmetric expansion (new) ---------------------------------------------------
 * The reference reads the banner's symmetry and ignores it (inc/helper_functions.h:143-156), so on
 * the shipped cant.mtx -- stored `symmetric`, lower triangle only -- it multiplies by the lower
 * triangle.  That stays the default.  With --expand-symmetric every off-diagonal entry (r, c, v) of
 * a symmetric / hermitian / skew-symmetric file also yields (c, r, +-v), and the result is sorted
 * by (row, column) with two stable counting sorts, which is the order the CSR / ELL / SELL / CMRS
 * builders need.  A `general` file is left untouched. */
void set_expand_symmetric(int enable) { g_expand_symmetric = enable; }
int last_banner_symmetry(void) { return g_banner_symmetry; }

static bool counting_sort_by(const int *key, int n_keys, int n, const int *in_perm, int *out_perm)
{
    size_t *start = (size_t *)calloc((size_t)n_keys + 1, sizeof(size_t));
    if (!start) return false;
    for (int i = 0; i < n; ++i) start[key[in_perm[i]] + 1]++;
    for (int k = 0; k < n_keys; ++k) start[k + 1] += start[k];
    for (int i = 0; i < n; ++i) out_perm[start[key[in_perm[i]]]++] = in_perm[i];
    free(start);
    return true;
}

bool expand_symmetric_entries(int number_of_rows, int number_of_columns, int symmetry, int *number_of_nonzeroes,
                              int **rows, int **cols, double **data)
{
    if (symmetry == 0) return true;
    const int nnz = *number_of_nonzeroes;
    long long extra = 0;
    for (int i = 0; i < nnz; ++i) {
        const int r = (*rows)[i], c = (*cols)[i];
        if (r < 0 || r >= number_of_rows || c < 0 || c >= number_of_columns) return false; /* not a valid entry */
        extra += r != c;
    }
    const long long total = (long long)nnz + extra;
    if (total > 0x7fffffffll || number_of_rows != number_of_columns) return false;
    const int n = (int)total;
    int *r = (int *)malloc(sizeof(int) * (size_t)n + 16), *c = (int *)malloc(sizeof(int) * (size_t)n + 16);
    double *v = (double *)malloc(sizeof(double) * (size_t)n + 16);
    int *p0 = (int *)malloc(sizeof(int) * (size_t)n + 16), *p1 = (int *)malloc(sizeof(int) * (size_t)n + 16);
    bool ok = r && c && v && p0 && p1;
    if (ok) {
        int at = nnz;
        for (int i = 0; i < nnz; ++i) {
            r[i] = (*rows)[i];
            c[i] = (*cols)[i];
            v[i] = (*data)[i];
            if (r[i] != c[i]) {
                r[at] = c[i];
                c[at] = r[i];
                v[at] = symmetry < 0 ? -v[i] : v[i];
                ++at;
            }
        }
        for (int i = 0; i < n; ++i) p0[i] = i;
        /* LSD: stable by column, then stable by row -> sorted by (row, column) */
        ok = counting_sort_by(c, number_of_columns, n, p0, p1) && counting_sort_by(r, number_of_rows, n, p1, p0);
    }
    if (ok) {
        int *r2 = (int *)malloc(sizeof(int) * (size_t)n + 16), *c2 = (int *)malloc(sizeof(int) * (size_t)n + 16);
        double *v2 = (double *)malloc(sizeof(double) * (size_t)n + 16);
        ok = r2 && c2 && v2;
        if (ok) {
            for (int i = 0; i < n; ++i) {
                r2[i] = r[p0[i]];
                c2[i] = c[p0[i]];
                v2[i] = v[p0[i]];
            }
            free(*rows);
            free(*cols);
            free(*data);
            *rows = r2;
            *cols = c2;
            *data = v2;
            *number_of_nonzeroes = n;
        } else {
            free(r2);
            free(c2);
            free(v2);
        }
    }
    free(r);
    free(c);
    free(v);
    free(p0);
    free(p1);
    return ok;
}

/* ---- optional binary cache of the parsed triples (new) ---------------------------------------
 * The reference parses its 60-90 MB text file three to four times per run (the builder's one or two
 * passes plus one per check_result, inc/helper_functions.h:209-219).  With --cache the first load
 * writes "<file>.b200cache" next to the matrix -- a 64-byte header (magic, version, source size and
 * mtime, dimensions, banner symmetry) followed by the raw rows / cols / data arrays as parsed, i.e.
 * BEFORE any symmetric expansion -- and every later load of an unchanged file is one fread.  A cache
 * that does not match the source (size, mtime, version) is ignored and rewritten. */
void set_use_cache(int enable) { g_use_cache = enable; }

typedef struct {
    char magic[8]; /* "B200MTX\0" */
    uint32_t version, symmetry_plus_one;
    int64_t source_size, source_mtime_ns;
    int32_t n_rows, n_cols, nnz, pad;
    char reserved[16];
} cache_header;

static bool cache_identity(const char *filename, int64_t *size, int64_t *mtime_ns)
{
    struct stat st;
    if (stat(filename, &st) != 0) return false;
    *size = (int64_t)st.st_size;
    *mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + (int64_t)st.st_mtim.tv_nsec;
    return true;
}

static char *cache_path(const char *filename)
{
    const size_t n = strlen(filename);
    char *p = (char *)malloc(n + 16);
    if (p) snprintf(p, n + 16, "%s.b200cache", filename);
    return p;
}

/* true = arrays (malloc'ed here) and sizes come from a valid cache */
static bool cache_load(const char *filename, int *n_rows, int *n_cols, int *nnz, int **rows, int **cols, double **data)
{
    int64_t size, mtime;
    char *path = cache_path(filename);
    if (!path || !cache_identity(filename, &size, &mtime)) {
        free(path);
        return false;
    }
    FILE *f = fopen(path, "rb");
    free(path);
    if (!f) return false;
    cache_header h;
    bool ok = fread(&h, sizeof h, 1, f) == 1 && !memcmp(h.magic, "B200MTX", 8) && h.version == 1 &&
              h.source_size == size && h.source_mtime_ns == mtime && h.nnz >= 0 && h.n_rows >= 0 && h.n_cols >= 0;
    int *r = NULL, *c = NULL;
    double *v = NULL;
    if (ok) {
        const size_t n = (size_t)h.nnz;
        r = (int *)malloc(sizeof(int) * n + 16);
        c = (int *)malloc(sizeof(int) * n + 16);
        v = (double *)malloc(sizeof(double) * n + 16);
        ok = r && c && v && fread(r, sizeof(int), n, f) == n && fread(c, sizeof(int), n, f) == n &&
             fread(v, sizeof(double), n, f) == n;
    }
    fclose(f);
    if (!ok) {
        free(r);
        free(c);
        free(v);
        return false;
    }
    *n_rows = h.n_rows;
    *n_cols = h.n_cols;
    *nnz = h.nnz;
    *rows = r;
    *cols = c;
    *data = v;
    g_banner_symmetry = (int)h.symmetry_plus_one - 1;
    return true;
}

static void cache_store(const char *filename, int n_rows, int n_cols, int nnz, const int *rows, const int *cols,
                        const double *data)
{
    int64_t size, mtime;
    char *path = cache_path(filename);
    if (!path || !cache_identity(filename, &size, &mtime)) {
        free(path);
        return;
    }
    const size_t len = strlen(path);
    char *tmp = (char *)malloc(len + 32);
    if (tmp) {
        snprintf(tmp, len + 32, "%s.tmp%ld", path, (long)getpid());
        FILE *f = fopen(tmp, "wb");
        if (f) {
            cache_header h;
            memset(&h, 0, sizeof h);
            memcpy(h.magic, "B200MTX", 8);
            h.version = 1;
            h.symmetry_plus_one = (uint32_t)(g_banner_symmetry + 1);
            h.source_size = size;
            h.source_mtime_ns = mtime;
            h.n_rows = n_rows;
            h.n_cols = n_cols;
            h.nnz = nnz;
            const size_t n = (size_t)nnz;
            const bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(rows, sizeof(int), n, f) == n &&
                            fwrite(cols, sizeof(int), n, f) == n && fwrite(data, sizeof(double), n, f) == n;
            if (fclose(f) != 0 || !ok || rename(tmp, path) != 0) remove(tmp); /* best effort */
        }
        free(tmp);
    }
    free(path);
}

/* header + entries of `filename` as the drivers parse them (1-based -> 0-based, file order), through
 * the cache when enabled; arrays are malloc'ed here (+16 spare bytes) */
bool load_triples(const char *filename, int *n_rows, int *n_cols, int *nnz, int **rows, int **cols, double **data)
{
    if (g_use_cache && cache_load(filename, n_rows, n_cols, nnz, rows, cols, data)) return true;
    FILE *file = fopen(filename, "r");
    if (file == NULL) {
        perror(filename);
        return false;
    }
    if (!read_size_of_matrices_from_file(file, n_rows, n_cols, nnz)) {
        fclose(file);
        return false;
    }
    *rows = (int *)malloc(sizeof(int) * (size_t)*nnz + 16);
    *cols = (int *)malloc(sizeof(int) * (size_t)*nnz + 16);
    *data = (double *)malloc(sizeof(double) * (size_t)*nnz + 16);
    const bool ok = *rows && *cols && *data && read_entries(file, *nnz, *rows, *cols, *data);
    fclose(file);
    if (!ok) {
        free(*rows);
        free(*cols);
        free(*data);
        *rows = *cols = NULL;
        *data = NULL;
        return false;
    }
    if (g_use_cache) cache_store(filename, *n_rows, *n_cols, *nnz, *rows, *cols, *data);
    return true;
}

void set_value_bytes(int bytes) { g_value_bytes = bytes; }
void set_check_tolerance(double relative_max_norm) { g_rel_tolerance = relative_max_norm; }

void calculate_and_print_performance(double ms, int number_of_nonzeroes)
{
    printf("Your calculations took %.2lf ms to run.\n", ms);
    printf("Number of operations %d, PERFORMANCE %lf GFlops\n", 2 * number_of_nonzeroes,
           (2 * number_of_nonzeroes) / ms * 1e-6);
}

void calculate_and_print_speed(double ms, int number_of_nonzeroes)
{
    const double lo = (double)number_of_nonzeroes * g_value_bytes;
    const double hi = 2.0 * number_of_nonzeroes * g_value_bytes;
    printf("GBytes transferred to processor %lf - %lf, speed %lf - %lf GB/s\n", lo * 1e-9, hi * 1e-9,
           lo / ms * 1e-6, hi / ms * 1e-6);
}

bool check_result(const char *filename, double *vect, double *result)
{
    int number_of_rows, number_of_columns, number_of_nonzeroes;
    int *rows = NULL, *cols = NULL;
    double *vals = NULL, *expect = NULL;
    bool ok;
    if (g_use_cache) { /* --cache: the triples come from the binary cache written by the first load */
        ok = load_triples(filename, &number_of_rows, &number_of_columns, &number_of_nonzeroes, &rows, &cols, &vals);
        if (!ok) return false;
        expect = (double *)calloc((size_t)number_of_rows, sizeof(double));
        ok = expect != NULL;
    } else {
        FILE *file = fopen(filename, "r");
        if (file == NULL) {
            perror(filename);
            return false;
        }
        if (!read_size_of_matrices_from_file(file, &number_of_rows, &number_of_columns, &number_of_nonzeroes)) {
            fclose(file);
            return false;
        }
        rows = (int *)malloc(sizeof(int) * (size_t)number_of_nonzeroes);
        cols = (int *)malloc(sizeof(int) * (size_t)number_of_nonzeroes);
        vals = (double *)malloc(sizeof(double) * (size_t)number_of_nonzeroes);
        expect = (double *)calloc((size_t)number_of_rows, sizeof(double));
        ok = rows && cols && vals && expect && read_entries(file, number_of_nonzeroes, rows, cols, vals);
        fclose(file);
    }
    if (ok && g_expand_symmetric)
        ok = expand_symmetric_entries(number_of_rows, number_of_columns, g_banner_symmetry, &number_of_nonzeroes,
                                      &rows, &cols, &vals);
    if (ok) {
        /* Per-row criterion (the reference's is per row too: |diff| <= 1e-6 absolute,
         * inc/helper_functions.h:221-228 -- below one fp64 ulp of cant's row sums, so it only ever
         * passes when the summation order matches).  Row i passes when
         * |diff| <= EPSILON, or |diff| <= tol * sum_j |a_ij x_j| (the scale rounding errors of that
         * row grow with; tol = 1e-12 fp64 / 1e-5 fp32).  The max-norm of BASELINE.json
         * (max|diff| / max|expect| <= tol) is implied, and a wrong or missing small-magnitude row
         * can no longer hide behind a large one. */
        double *abs_sum = (double *)calloc((size_t)number_of_rows, sizeof(double));
        if (!abs_sum) ok = false;
        for (int i = 0; ok && i < number_of_nonzeroes; ++i) {
            const double t = vals[i] * vect[cols[i]];
            expect[rows[i]] += t;
            abs_sum[rows[i]] += fabs(t);
        }
        int first_bad = -1;
        for (int i = 0; ok && i < number_of_rows; ++i) {
            const double d = fabs(expect[i] - result[i]);
            const double scale = fabs(expect[i]) > abs_sum[i] ? fabs(expect[i]) : abs_sum[i];
            if (!(d <= EPSILON) && !(d <= g_rel_tolerance * scale)) { /* NaN fails both */
                first_bad = i;
                break;
            }
        }
        if (first_bad >= 0) {
            printf("wrong value at index %d: expected %f - calculated %f\n", first_bad,
                   expect[first_bad], result[first_bad]);
            ok = false;
        }
        free(abs_sum);
    }
    free(rows);
    free(cols);
    free(vals);
    free(expect);
    return ok;
}

double now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec * 1000.0 + (double)t.tv_nsec / 1000000.0;
}

int report_b200_error(const char *what, int status)
{
    printf("%s error %d\n", what, status);
    fprintf(stderr, "%s: %s\n", b200_status_string(status), b200_last_error());
    return status == B200_ERR_NO_DEVICE ? OpenCLDeviceError : OpenCLProgramError;
}

/* ---- shared prologue / epilogue of the drivers -------------------------------------------- */
int driver_load_matrix(const driver_options *opt, host_matrix *m)
{
    int number_of_devices = 0;
    memset(m, 0, sizeof *m);
    if (b200_get_device_count(&number_of_devices) != B200_SUCCESS) {
        printf("No CUDA devices found\n");
        return OpenCLDeviceError;
    }
    if (number_of_devices > DEVICES_DEFAULT_SIZE) number_of_devices = DEVICES_DEFAULT_SIZE;
    if (opt->device >= number_of_devices) return OpenCLDeviceError;

    if (g_use_cache) { /* --cache */
        if (!load_triples(opt->matrix, &m->n_rows, &m->n_cols, &m->nnz, &m->rows, &m->cols, &m->data)) return FileError;
        m->vect = (double *)malloc((size_t)m->n_cols * sizeof(double) + 16);
        if (!m->vect) return FileError;
    } else {
        FILE *file = fopen(opt->matrix, "r");
        if (file == NULL) {
            perror(opt->matrix);
            return FileError;
        }
        if (read_size_of_matrices_from_file(file, &m->n_rows, &m->n_cols, &m->nnz) == false) {
            fclose(file);
            return FileError;
        }
        m->rows = (int *)malloc((size_t)m->nnz * sizeof(int) + 16);
        m->cols = (int *)malloc((size_t)m->nnz * sizeof(int) + 16);
        m->data = (double *)malloc((size_t)m->nnz * sizeof(double) + 16);
        m->vect = (double *)malloc((size_t)m->n_cols * sizeof(double) + 16);
        if (!m->rows || !m->cols || !m->data || !m->vect ||
            !read_entries(file, m->nnz, m->rows, m->cols, m->data)) {
            fclose(file);
            return FileError;
        }
        fclose(file);
    }
    if (g_expand_symmetric &&
        !expand_symmetric_entries(m->n_rows, m->n_cols, g_banner_symmetry, &m->nnz, &m->rows, &m->cols, &m->data)) {
        printf("Could not expand the symmetric matrix.\n");
        return FileError;
    }
    for (int i = 0; i < m->n_cols; ++i) m->vect[i] = i; /* x = ramp, csr.c:95-99 */
    return Success;
}

void driver_free_matrix(host_matrix *m)
{
    free(m->rows);
    free(m->cols);
    free(m->data);
    free(m->vect);
    memset(m, 0, sizeof *m);
}

int driver_upload_triples(b200_ctx *ctx, const host_matrix *m, int use_f32, device_triples *d)
{
    const size_t nnz = (size_t)m->nnz, V = use_f32 ? sizeof(float) : sizeof(double);
    memset(d, 0, sizeof *d);
    B200_TRY(b200_malloc(ctx, sizeof(int) * nnz, &d->rows));
    B200_TRY(b200_malloc(ctx, sizeof(int) * nnz, &d->cols));
    B200_TRY(b200_malloc(ctx, sizeof(double) * nnz, &d->data64));
    B200_TRY(b200_malloc(ctx, V * (size_t)m->n_cols, &d->vect));
    B200_TRY(b200_memcpy_h2d_async(ctx, d->rows, m->rows, sizeof(int) * nnz));
    B200_TRY(b200_memcpy_h2d_async(ctx, d->cols, m->cols, sizeof(int) * nnz));
    B200_TRY(b200_memcpy_h2d_async(ctx, d->data64, m->data, sizeof(double) * nnz));
    if (use_f32) B200_TRY(b200_fill_ramp_f32(ctx, (float *)d->vect, m->n_cols));
    else B200_TRY(b200_memcpy_h2d_async(ctx, d->vect, m->vect, sizeof(double) * (size_t)m->n_cols));
    return Success;
}

void driver_free_triples(b200_ctx *ctx, device_triples *d)
{
    b200_free(ctx, d->rows);
    b200_free(ctx, d->cols);
    b200_free(ctx, d->data64);
    b200_free(ctx, d->vect);
    memset(d, 0, sizeof *d);
}

int driver_read_output(b200_ctx *ctx, const void *buffer_output, int n, int use_f32, double *output)
{
    if (use_f32) {
        float *tmp = (float *)malloc(sizeof(float) * (size_t)n + 16);
        if (!tmp) return OtherError;
        int status = b200_memcpy_d2h(ctx, tmp, buffer_output, sizeof(float) * (size_t)n);
        if (status != B200_SUCCESS) {
            free(tmp);
            return report_b200_error("b200_memcpy_d2h", status);
        }
        for (int i = 0; i < n; ++i) output[i] = tmp[i];
        free(tmp);
        return Success;
    }
    B200_TRY(b200_memcpy_d2h(ctx, output, buffer_output, sizeof(double) * (size_t)n));
    return Success;
}
