/* driver_iterate.c -- the iterated, multi-GPU mode of the drivers:  ./bin/sigma_c --gpus N --iters K [--json]
 *
 * The reference enumerates up to DEVICES_DEFAULT_SIZE GPUs (csr.c:12,22-30) and then leaves its device
 * loop after the first one (csr.c:279).  This is where that loop would go on: one host thread per
 * device, every thread owning one context, one row block of the matrix (equal blocks, multiples of 32
 * rows) and one rank of an NCCL communicator the library creates (b200_comm_*).  The K steps of
 *     y = A x / ||x||_2 ;  x <- y
 * are issued by the library (b200_iterator_*): sigma_c runs the fused SELL kernel with halo-limited
 * peer stores (peer access between the devices of this process), csr the SpMV + ncclAllGather
 * formulation.  Host code stays C; nothing here computes y.
 *
 * Matrix: the loaded MatrixMarket file (square, row-sorted), or --synthetic laplace7:NXxNYxNZ generated
 * on the devices (the 64 M-row Laplacian of BASELINE.json configs[4] cannot go through text).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "helper_functions.h"

/* untimed steps before the K timed ones */
#define DRIVER_ITER_WARMUP(iters) ((iters) >= 10 ? 1 + 2 * ((iters) / 2 * 2) : 1)

typedef struct {
    /* shared, read-only after start */
    const driver_options *opt;
    const host_matrix *m; /* NULL with --synthetic */
    int format;           /* B200_FORMAT_SELL (fused) or B200_FORMAT_CSR (all-gather) */
    int world, nx, ny, nz;
    long long n_rows, rows_per_rank;
    unsigned char comm_id[B200_COMM_ID_BYTES];
    pthread_barrier_t *barrier;
    /* shared, written by the ranks between barriers */
    double **x[2];        /* x[b][r] */
    int *col_min, *col_max;
    long long *nnz;
    int *status;          /* per rank: first failing ReturnCode */
    double *ms_per_step, *norm;
    unsigned long long *launches;
    b200_mcast *mcast_owner; /* --sync mcast: rank 0's multicast object, shared by the other threads */
    int *mcast_status;
} shared_state;

typedef struct {
    shared_state *s;
    int rank;
} rank_arg;

#define RANK_TRY(call)                                                      \
    do {                                                                    \
        int status__ = (call);                                              \
        if (status__ != B200_SUCCESS && rc == Success) rc = report_b200_error(#call, status__); \
    } while (0)

static void wait_all(shared_state *s) { pthread_barrier_wait(s->barrier); }

static void *rank_main(void *argp)
{
    rank_arg *arg = (rank_arg *)argp;
    shared_state *s = arg->s;
    const int r = arg->rank, world = s->world;
    const driver_options *opt = s->opt;
    int rc = Success;
    b200_ctx *ctx = NULL;
    b200_comm *comm = NULL;
    b200_iterator *it = NULL;
    b200_csr_plan *plan = NULL;
    void *d_rows = NULL, *d_cols = NULL, *d_vals = NULL, *d_ptr = NULL, *d_sptr = NULL, *d_ri = NULL, *d_sc = NULL, *d_sd = NULL;
    void *xb[2] = {NULL, NULL};

    long long lo = (long long)r * s->rows_per_rank, hi = lo + s->rows_per_rank;
    if (lo > s->n_rows) lo = s->n_rows;
    if (hi > s->n_rows) hi = s->n_rows;
    const int n_local = (int)(hi - lo);
    const long long padded = s->rows_per_rank * world;
    long long nnz = 0;

    RANK_TRY(b200_ctx_create(opt->device + r, &ctx));
    if (rc == Success)
        for (int p = 0; p < world; ++p) RANK_TRY(b200_ctx_enable_peer_access(ctx, opt->device + p));
    /* this rank's row block as device triples (rows rebased to the block, columns global) */
    if (rc == Success) {
        const int *h_rows = NULL;
        long long first = 0;
        if (s->m) {
            const host_matrix *m = s->m;
            long long a = 0, b = m->nnz; /* entries with row in [lo, hi): the file is row-sorted */
            while (a < b) {
                const long long mid = (a + b) / 2;
                if (m->rows[mid] < lo) a = mid + 1;
                else b = mid;
            }
            first = a;
            b = m->nnz;
            while (a < b) {
                const long long mid = (a + b) / 2;
                if (m->rows[mid] < hi) a = mid + 1;
                else b = mid;
            }
            nnz = a - first;
            h_rows = m->rows + first;
        } else {
            nnz = b200_gen_laplace7_nnz(s->nx, s->ny, s->nz, (int)lo, n_local);
        }
        RANK_TRY(b200_malloc(ctx, sizeof(int) * (size_t)nnz, &d_rows));
        RANK_TRY(b200_malloc(ctx, sizeof(int) * (size_t)nnz, &d_cols));
        RANK_TRY(b200_malloc(ctx, sizeof(double) * (size_t)nnz, &d_vals));
        if (rc == Success && s->m) {
            RANK_TRY(b200_memcpy_h2d_async(ctx, d_rows, h_rows, sizeof(int) * (size_t)nnz));
            RANK_TRY(b200_memcpy_h2d_async(ctx, d_cols, s->m->cols + first, sizeof(int) * (size_t)nnz));
            RANK_TRY(b200_memcpy_h2d_async(ctx, d_vals, s->m->data + first, sizeof(double) * (size_t)nnz));
        } else if (rc == Success) {
            RANK_TRY(b200_gen_laplace7_coo(ctx, s->nx, s->ny, s->nz, (int)lo, n_local, (int *)d_rows, (int *)d_cols, (double *)d_vals));
        }
        RANK_TRY(b200_offset_i32(ctx, (int *)d_rows, nnz, (int)-lo));
        RANK_TRY(b200_check_sorted_rows(ctx, (const int *)d_rows, (int)nnz, n_local));
    }
    s->nnz[r] = nnz;
    /* format build on the device: CSR row pointer, then SELL-32 for the fused mode */
    b200_block_f64 blk;
    memset(&blk, 0, sizeof blk);
    blk.format = s->format;
    blk.n_rows = n_local;
    if (rc == Success) {
        RANK_TRY(b200_malloc(ctx, sizeof(int) * ((size_t)n_local + 1), &d_ptr));
        RANK_TRY(b200_build_csr_ptr(ctx, (const int *)d_rows, (int)nnz, n_local, (int *)d_ptr));
    }
    if (rc == Success && s->format == B200_FORMAT_SELL) {
        const int n_slices = b200_sell_num_slices(n_local, 32);
        long long total = 0;
        RANK_TRY(b200_malloc(ctx, sizeof(long long) * ((size_t)n_slices + 1), &d_sptr));
        RANK_TRY(b200_malloc(ctx, sizeof(int) * ((size_t)n_slices + 1), &d_ri));
        RANK_TRY(b200_build_sell_ptr(ctx, (const int *)d_ptr, n_local, 32, 1, NULL, (long long *)d_sptr, &total));
        RANK_TRY(b200_sell_ptr_to_i32(ctx, (const long long *)d_sptr, n_slices, (int *)d_ri));
        RANK_TRY(b200_malloc(ctx, sizeof(int) * (size_t)total, &d_sc));
        RANK_TRY(b200_malloc(ctx, sizeof(double) * (size_t)total, &d_sd));
        RANK_TRY(b200_build_sell_fill_f64(ctx, (const int *)d_ptr, (const int *)d_cols, (const double *)d_vals, n_local, 32,
                                          NULL, (const long long *)d_sptr, (int *)d_sc, (double *)d_sd));
        blk.n_slices = n_slices;
        blk.ptr = (const int *)d_ri;
        blk.indices = (const int *)d_sc;
        blk.data = (const double *)d_sd;
    } else if (rc == Success) {
        RANK_TRY(b200_csr_plan_create(ctx, (const int *)d_ptr, n_local, &plan));
        blk.ptr = (const int *)d_ptr;
        blk.indices = (const int *)d_cols;
        blk.data = (const double *)d_vals;
        blk.csr_plan = plan;
    }
    /* the two alternating x buffers of this rank; x0 = the ramp of the reference (csr.c:95-99) for a
     * file, a seeded uniform vector for the synthetic matrix (bench.py's start vector) */
    for (int b = 0; b < 2 && rc == Success; ++b) {
        RANK_TRY(b200_malloc(ctx, sizeof(double) * (size_t)padded, &xb[b]));
        RANK_TRY(b200_memset_async(ctx, xb[b], 0, sizeof(double) * (size_t)padded));
    }
    if (rc == Success) {
        if (s->m) RANK_TRY(b200_fill_ramp_f64(ctx, (double *)xb[0], (int)s->n_rows));
        else RANK_TRY(b200_gen_uniform_f64(ctx, (double *)xb[0], s->n_rows, 11, 0.0, 1.0));
    }
    int cmin = 0, cmax = -1;
    if (rc == Success) RANK_TRY(b200_minmax_i32(ctx, (const int *)d_cols, nnz, &cmin, &cmax));
    if (rc == Success) RANK_TRY(b200_sync(ctx));
    /* --sync mcast: the per-step all-reduce + barrier through NVSwitch multicast instead of NCCL.  Rank 0
     * creates the object, the other threads share it; add (all) -> barrier -> bind (all) -> barrier. */
    b200_mcast *mc = NULL;
    const int want_mcast = opt->sync_mcast && world > 1 && s->format == B200_FORMAT_SELL;
    if (want_mcast) {
        int st = B200_SUCCESS;
        if (r == 0) {
            st = rc == Success ? b200_mcast_create(ctx, world, &mc, NULL) : B200_ERR_INVALID_VALUE;
            s->mcast_owner = mc;
            s->mcast_status[0] = st;
        }
        wait_all(s);
        if (r != 0) st = (rc == Success && s->mcast_status[0] == B200_SUCCESS) ? b200_mcast_share(ctx, s->mcast_owner, &mc) : B200_ERR_INVALID_VALUE;
        if (st == B200_SUCCESS && s->mcast_status[0] == B200_SUCCESS) st = b200_mcast_add_device(mc);
        s->mcast_status[r] = st;
        wait_all(s);
        int all_ok = 1;
        for (int p = 0; p < world; ++p) all_ok &= s->mcast_status[p] == B200_SUCCESS;
        st = all_ok ? b200_mcast_bind(mc, r) : B200_ERR_UNSUPPORTED;
        wait_all(s);
        s->mcast_status[r] = st;
        wait_all(s);
        for (int p = 0; p < world; ++p) all_ok &= s->mcast_status[p] == B200_SUCCESS;
        if (!all_ok && rc == Success) {
            if (r == 0) fprintf(stderr, "NVSwitch multicast is not available here (%s): %s\n", b200_status_string(B200_ERR_UNSUPPORTED), b200_last_error());
            rc = OpenCLDeviceError;
        }
    }
    s->x[0][r] = (double *)xb[0];
    s->x[1][r] = (double *)xb[1];
    s->col_min[r] = cmin;
    s->col_max[r] = cmax;
    s->status[r] = rc;
    wait_all(s); /* every rank's buffers, ranges and status are published */
    for (int p = 0; p < world; ++p)
        if (s->status[p] != Success && rc == Success) rc = s->status[p]; /* all ranks give up together */

    int *halo_lo = (int *)calloc((size_t)world, sizeof(int)), *halo_hi = (int *)calloc((size_t)world, sizeof(int));
    if (rc == Success && world > 1 && !want_mcast) RANK_TRY(b200_comm_create(ctx, s->comm_id, r, world, &comm));
    if (rc == Success) {
        RANK_TRY(b200_halo_rows(s->col_min, s->col_max, world, r, s->rows_per_rank, s->n_rows, halo_lo, halo_hi));
        b200_iter_desc d;
        memset(&d, 0, sizeof d);
        d.mode = s->format == B200_FORMAT_SELL ? (want_mcast ? B200_ITER_FUSED_MCAST : B200_ITER_FUSED) : B200_ITER_ALLGATHER;
        d.mcast = want_mcast ? mc : NULL;
        d.world = world;
        d.rank = r;
        d.rows_per_rank = s->rows_per_rank;
        d.x[0] = s->x[0];
        d.x[1] = s->x[1];
        if (d.mode != B200_ITER_ALLGATHER) { /* both fused modes send only the rows a peer reads */
            d.halo_lo = halo_lo;
            d.halo_hi = halo_hi;
        }
        /* the K timed steps are ONE replay of a K-step launch graph (a replay of a graph that holds NCCL
         * nodes ends in a host callback, so short graphs would pay it every few steps) */
        d.graph_steps = opt->iters >= 10 ? opt->iters / 2 * 2 : 0;
        RANK_TRY(b200_iterator_create(ctx, comm, &blk, &d, &it));
    }
    /* warm-up, untimed: step 0 (always issued directly: NCCL sets itself up there) and, with a graph, two
     * replays that record and warm it.  The norm printed is the one after warm-up + K steps. */
    if (rc == Success) RANK_TRY(b200_iterator_run(it, DRIVER_ITER_WARMUP(opt->iters)));
    /* timed region: K steps, bracketed by a barrier over the ranks and a sync of every queue */
    if (rc == Success) RANK_TRY(b200_sync(ctx));
    wait_all(s);
    const double t0 = now_ms();
    if (rc == Success) RANK_TRY(b200_iterator_run(it, opt->iters));
    double norm = NAN;
    if (rc == Success) RANK_TRY(b200_iterator_norm(it, &norm)); /* waits for the queue */
    wait_all(s);
    s->ms_per_step[r] = (now_ms() - t0) / opt->iters;
    s->norm[r] = norm;
    if (it) b200_iterator_state(it, NULL, NULL, &s->launches[r]);
    s->status[r] = rc;

    b200_iterator_destroy(it);
    wait_all(s); /* nobody unbinds the multicast block while a peer's last step is still signalling */
    if (r != 0) b200_mcast_destroy(mc);
    wait_all(s);
    if (r == 0) b200_mcast_destroy(mc);
    b200_comm_destroy(comm);
    b200_csr_plan_destroy(plan);
    wait_all(s); /* nobody frees a buffer a peer might still be storing into */
    void *bufs[] = {d_rows, d_cols, d_vals, d_ptr, d_sptr, d_ri, d_sc, d_sd, xb[0], xb[1]};
    for (size_t i = 0; i < sizeof bufs / sizeof bufs[0]; ++i) b200_free(ctx, bufs[i]);
    b200_ctx_destroy(ctx);
    free(halo_lo);
    free(halo_hi);
    return NULL;
}

/* ||A x / ||x|| || after `iters` steps on the host: serial fp64, entries in file order, like check_result */
static double cpu_power_iteration(const host_matrix *m, int iters)
{
    const int n = m->n_rows;
    double *x = (double *)malloc(sizeof(double) * (size_t)n), *y = (double *)malloc(sizeof(double) * (size_t)n);
    double norm = NAN;
    if (!x || !y) {
        free(x);
        free(y);
        return NAN;
    }
    for (int i = 0; i < n; ++i) x[i] = m->vect[i];
    for (int k = 0; k < iters; ++k) {
        memset(y, 0, sizeof(double) * (size_t)n);
        for (int i = 0; i < m->nnz; ++i) y[m->rows[i]] += m->data[i] * x[m->cols[i]];
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += y[i] * y[i];
        norm = sqrt(s);
        /* the device path leaves y unscaled and divides by its norm inside the next product */
        for (int i = 0; i < n; ++i) x[i] = y[i] / norm;
    }
    free(x);
    free(y);
    return norm;
}

int driver_run_iterated(const driver_options *opt, const host_matrix *m, int format, const char *driver_name)
{
    shared_state s;
    memset(&s, 0, sizeof s);
    int n_devices = 0;
    if (b200_get_device_count(&n_devices) != B200_SUCCESS) {
        printf("No CUDA devices found\n");
        return OpenCLDeviceError;
    }
    if (n_devices > DEVICES_DEFAULT_SIZE) n_devices = DEVICES_DEFAULT_SIZE;
    const int world = opt->gpus;
    if (opt->device + world > n_devices) {
        printf("Asked for %d device(s) from device %d, found %d\n", world, opt->device, n_devices);
        return OpenCLDeviceError;
    }
    s.opt = opt;
    s.format = format;
    s.world = world;
    if (opt->synthetic) {
        if (sscanf(opt->synthetic, "laplace7:%dx%dx%d", &s.nx, &s.ny, &s.nz) != 3 || s.nx < 1 || s.ny < 1 || s.nz < 1 ||
            (long long)s.nx * s.ny * s.nz > 0x7fffffffll) {
            fprintf(stderr, "--synthetic expects laplace7:NXxNYxNZ with at most 2^31-1 rows\n");
            return OtherError;
        }
        s.n_rows = (long long)s.nx * s.ny * s.nz;
    } else {
        if (m->n_rows != m->n_cols) {
            printf("The iterated mode needs a square matrix (%d x %d).\n", m->n_rows, m->n_cols);
            return OtherError;
        }
        s.m = m;
        s.n_rows = m->n_rows;
    }
    const long long per = (s.n_rows + world - 1) / world;
    s.rows_per_rank = (per + 31) / 32 * 32;
    if (world > 1 && !(opt->sync_mcast && format == B200_FORMAT_SELL)) {
        int st = b200_comm_get_unique_id(s.comm_id);
        if (st != B200_SUCCESS) return report_b200_error("b200_comm_get_unique_id", st);
    }
    pthread_barrier_t barrier;
    pthread_barrier_init(&barrier, NULL, (unsigned)world);
    s.barrier = &barrier;
    s.x[0] = (double **)calloc((size_t)world, sizeof(double *));
    s.x[1] = (double **)calloc((size_t)world, sizeof(double *));
    s.col_min = (int *)calloc((size_t)world, sizeof(int));
    s.col_max = (int *)calloc((size_t)world, sizeof(int));
    s.nnz = (long long *)calloc((size_t)world, sizeof(long long));
    s.status = (int *)calloc((size_t)world, sizeof(int));
    s.ms_per_step = (double *)calloc((size_t)world, sizeof(double));
    s.norm = (double *)calloc((size_t)world, sizeof(double));
    s.launches = (unsigned long long *)calloc((size_t)world, sizeof(unsigned long long));
    s.mcast_status = (int *)calloc((size_t)world, sizeof(int));
    pthread_t *threads = (pthread_t *)calloc((size_t)world, sizeof(pthread_t));
    rank_arg *args = (rank_arg *)calloc((size_t)world, sizeof(rank_arg));
    for (int r = 0; r < world; ++r) {
        args[r].s = &s;
        args[r].rank = r;
        pthread_create(&threads[r], NULL, rank_main, &args[r]);
    }
    for (int r = 0; r < world; ++r) pthread_join(threads[r], NULL);
    pthread_barrier_destroy(&barrier);

    int rc = Success;
    double ms = 0.0;
    long long nnz = 0;
    for (int r = 0; r < world; ++r) {
        if (s.status[r] != Success && rc == Success) rc = s.status[r];
        if (s.ms_per_step[r] > ms) ms = s.ms_per_step[r]; /* max over ranks */
        nnz += s.nnz[r];
    }
    if (rc == Success) {
        const double norm = s.norm[0];
        int ok = 1, checked = 0;
        for (int r = 1; r < world; ++r)
            if (s.norm[r] != norm) ok = 0; /* every rank reduced the same sums */
        double cpu_norm = NAN;
        if (s.m && !opt->no_cpu && (double)m->nnz * (DRIVER_ITER_WARMUP(opt->iters) + opt->iters) <= 2e10) {
            cpu_norm = cpu_power_iteration(m, DRIVER_ITER_WARMUP(opt->iters) + opt->iters);
            checked = 1;
            if (!(fabs(cpu_norm - norm) <= 1e-10 * fabs(cpu_norm))) ok = 0;
        }
        const double gflops = 2.0 * (double)nnz / ms * 1e-6;
        if (opt->json) {
            printf("{\"driver\": \"%s\", \"mode\": \"%s\", \"gpus\": %d, \"iters\": %d, \"rows\": %lld, \"nnz\": %lld, "
                   "\"warmup\": %d, \"ms_per_step\": %.6f, \"gflops\": %.3f, \"norm\": %.17g, \"cpu_norm\": %s%.17g%s, \"checked\": %s, "
                   "\"ok\": %s, \"launches_rank0\": %llu, \"timing\": \"wall clock around b200_iterator_run + "
                   "b200_iterator_norm, barrier on both sides, max over ranks\"}\n",
                   driver_name,
                   format == B200_FORMAT_SELL ? (opt->sync_mcast && world > 1 ? "fused halo exchange + NVSwitch multicast all-reduce"
                                                                            : "fused halo exchange + all-reduce")
                                              : "SpMV + ncclAllGather",
                   world, opt->iters, s.n_rows, nnz, DRIVER_ITER_WARMUP(opt->iters), ms, gflops, norm, checked ? "" : "\"", checked ? cpu_norm : 0.0,
                   checked ? "" : " (not run)\"", checked ? "true" : "false", ok ? "true" : "false", s.launches[0]);
        } else {
            printf("Power iteration: %d steps (after %d warm-up steps), %d device(s), %lld rows, %lld nonzeroes\n", opt->iters,
                   DRIVER_ITER_WARMUP(opt->iters), world, s.n_rows, nnz);
            printf("Your calculations took %.4lf ms per step to run.\n", ms);
            printf("Number of operations %lld per step, PERFORMANCE %lf GFlops\n", 2 * nnz, gflops);
            printf("eigenvalue estimate %.15g\n", norm);
            if (checked) printf(ok ? "result is ok\n" : "result is wrong\n");
            else printf(ok ? "result not checked against the CPU (matrix too large or --no-cpu)\n" : "result is wrong\n");
        }
        if (!ok) rc = OtherError;
    }
    free(s.x[0]);
    free(s.x[1]);
    free(s.col_min);
    free(s.col_max);
    free(s.nnz);
    free(s.status);
    free(s.ms_per_step);
    free(s.norm);
    free(s.launches);
    free(s.mcast_status);
    free(threads);
    free(args);
    return rc;
}
