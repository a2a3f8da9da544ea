/*
 * gen_mtx -- writes the cant-SHAPED stand-in matrices as MatrixMarket text.
 *
 * The reference ships databases/cant.mtx and databases/cant-sorted.mtx only as Git-LFS pointer
 * stubs (databases/cant.mtx:1-3), so the real SuiteSparse `cant` is unobtainable offline.  This
 * tool generates a matrix of the same shape: a 27-point stencil on an NX x NY x NZ node grid
 * (z fastest) with DOF unknowns per node and dense DOF x DOF coupling blocks.  The default
 * 9 x 9 x 257 x 3 gives 62 451 rows (= cant) and 4 325 625 nnz, row length 24 / 69.26 / 81
 * (cant: 4 007 383 nnz, max 78); its lower triangle has 2 194 038 nnz (cant file: 2 034 917).
 *
 * Values are a symmetric integer hash of (min(r,c), max(r,c), seed) mapped to the 6-decimal
 * grid in [-1, 1] \ {0}, so every value round-trips exactly through "%lg".
 *
 *   gen_mtx --out FILE [--grid NX NY NZ] [--dof D] [--order row|col] [--tri full|lower]
 *           [--banner general|symmetric] [--seed S]
 *
 * --order row : sorted by (row, col)  -> what csr/ell/sigma_c/cmrs read (cant-sorted.mtx)
 * --order col : sorted by (col, row)  -> what coo reads (cant.mtx is column-major on disk)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double entry_value(long r, long c, uint64_t seed)
{
    long lo = r < c ? r : c, hi = r < c ? c : r;
    uint64_t h = mix64(mix64((uint64_t)lo * 0x100000001B3ull + seed) ^ (uint64_t)hi);
    long q = (long)(h % 2000001ull) - 1000000; /* [-1e6, 1e6] */
    if (q == 0) q = 1;
    return (double)q / 1e6;
}

typedef struct {
    int nx, ny, nz, dof;
} grid_t;

/* neighbours of unknown u, ascending; returns count */
static int neighbours(const grid_t *g, long u, long *out)
{
    long node = u / g->dof;
    int z = (int)(node % g->nz), y = (int)((node / g->nz) % g->ny), x = (int)(node / ((long)g->nz * g->ny));
    int n = 0;
    for (int dx = -1; dx <= 1; ++dx) {
        int xx = x + dx;
        if (xx < 0 || xx >= g->nx) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = y + dy;
            if (yy < 0 || yy >= g->ny) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                int zz = z + dz;
                if (zz < 0 || zz >= g->nz) continue;
                long nb = ((long)xx * g->ny + yy) * g->nz + zz;
                for (int d = 0; d < g->dof; ++d) out[n++] = nb * g->dof + d;
            }
        }
    }
    return n;
}

int main(int argc, char **argv)
{
    grid_t g = {9, 9, 257, 3};
    const char *out = NULL, *order = "row", *tri = "full", *banner = "general";
    uint64_t seed = 42;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--grid") && i + 3 < argc) {
            g.nx = atoi(argv[++i]);
            g.ny = atoi(argv[++i]);
            g.nz = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--dof") && i + 1 < argc) g.dof = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--order") && i + 1 < argc) order = argv[++i];
        else if (!strcmp(argv[i], "--tri") && i + 1 < argc) tri = argv[++i];
        else if (!strcmp(argv[i], "--banner") && i + 1 < argc) banner = argv[++i];
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 10);
        else {
            fprintf(stderr, "gen_mtx: bad argument '%s'\n", argv[i]);
            return 4;
        }
    }
    if (!out || g.nx < 1 || g.ny < 1 || g.nz < 1 || g.dof < 1) {
        fprintf(stderr, "usage: gen_mtx --out FILE [--grid NX NY NZ] [--dof D] [--order row|col] "
                        "[--tri full|lower] [--banner general|symmetric] [--seed S]\n");
        return 4;
    }
    int by_col = !strcmp(order, "col"), lower = !strcmp(tri, "lower");
    long n = (long)g.nx * g.ny * g.nz * g.dof;
    long *nb = (long *)malloc(sizeof(long) * 27 * (size_t)g.dof);

    /* count first: the size line precedes the entries */
    long nnz = 0;
    for (long u = 0; u < n; ++u) {
        int k = neighbours(&g, u, nb);
        for (int j = 0; j < k; ++j) {
            long r = by_col ? nb[j] : u, c = by_col ? u : nb[j];
            if (!lower || c <= r) ++nnz;
        }
    }
    FILE *f = fopen(out, "w");
    if (!f) {
        perror(out);
        return 3;
    }
    static char iobuf[1 << 22];
    setvbuf(f, iobuf, _IOFBF, sizeof iobuf);
    fprintf(f, "%%%%MatrixMarket matrix coordinate real %s\n", banner);
    fprintf(f, "%% cant-shaped stand-in: 27-point stencil, grid %dx%dx%d, %d dof/node, seed %llu, %s, %s-major\n",
            g.nx, g.ny, g.nz, g.dof, (unsigned long long)seed, tri, by_col ? "column" : "row");
    fprintf(f, "%ld %ld %ld\n", n, n, nnz);
    for (long u = 0; u < n; ++u) {
        int k = neighbours(&g, u, nb);
        for (int j = 0; j < k; ++j) {
            long r = by_col ? nb[j] : u, c = by_col ? u : nb[j];
            if (lower && c > r) continue;
            fprintf(f, "%ld %ld %.6f\n", r + 1, c + 1, entry_value(r, c, seed));
        }
    }
    fclose(f);
    free(nb);
    return 0;
}
