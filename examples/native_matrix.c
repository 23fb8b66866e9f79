/* A plain-C caller of the C ABI (include/pharmsol_cuda.h) — what a cgo / Rust-FFI / JNI binding does, without any
 * Python in the way.  It rebuilds the reference's own criterion shapes (benches/native_matrix.rs:23-24 with the data
 * of benches/common/mod.rs:117-271: 32 subjects x 64 support points; `1cpt-12h-po` 1 bolus + 9 observations,
 * `2cpt-120h-q12h` 10 boluses + 14 observations; additive ErrorPoly(0.1, 0.1, 0, 0)), evaluates the
 * log-likelihood matrix through the host-buffer call, and prints per case one JSON line with the median wall time
 * of 1000 calls plus a checksum (sum of the matrix) that tests/test_gpu_c_caller.py compares with the Python path.
 *
 *   gcc -O2 -std=c11 -Iinclude examples/native_matrix.c -Lpharmsol_b200 -lpharmsol_cuda -Wl,-rpath,'$ORIGIN/../pharmsol_b200' -lm
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pharmsol_cuda.h"

#define NSUB 32
#define NSPP 64
#define CHECK(call)                                                                                       \
    do {                                                                                                  \
        int32_t rc_ = (call);                                                                             \
        if (rc_ != PCU_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, pharmsol_cuda_last_error_message()); return 1; } \
    } while (0)

static const char* kShortAnalytical =
    "name = bench_short_a\nkind = analytical\nparams = ka, ke, v\nstates = gut, central\noutputs = plasma\nbolus(po) -> gut\n"
    "structure = one_compartment_with_absorption\nout(plasma) = central / v ~ continuous()\n";
static const char* kShortOde =
    "name = bench_short_o\nkind = ode\nparams = ka, ke, v\nstates = gut, central\noutputs = plasma\nbolus(po) -> gut\n"
    "dx(gut) = -ka * gut\ndx(central) = ka * gut - ke * central\nout(plasma) = central / v ~ continuous()\n";
static const char* kRepeatAnalytical =
    "name = bench_repeat_a\nkind = analytical\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = plasma\nbolus(iv) -> central\n"
    "structure = two_compartments\nout(plasma) = central / v ~ continuous()\n";
static const char* kRepeatOde =
    "name = bench_repeat_o\nkind = ode\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = plasma\nbolus(iv) -> central\n"
    "dx(central) = -(ke + kcp) * central + kpc * peripheral\ndx(peripheral) = kcp * central - kpc * peripheral\n"
    "out(plasma) = central / v ~ continuous()\n";

static const double kShortT[9] = {0.25, 0.5, 1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0};
static const double kShortY[9] = {0.50, 0.90, 1.60, 2.40, 2.10, 1.50, 1.05, 0.72, 0.48};
static const double kRepeatT[14] = {0.5, 2.0, 6.0, 10.0, 14.0, 24.0, 36.0, 48.0, 60.0, 72.0, 84.0, 96.0, 108.0, 120.0};
static const double kRepeatY[14] = {1.80, 1.45, 1.10, 0.90, 1.30, 1.60, 1.55, 1.50, 1.48, 1.45, 1.43, 1.42, 1.41, 0.95};

static int cmp_double(const void* a, const void* b) { const double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

static pcu_data* make_data(int repeat) {
    pcu_data* d = pharmsol_data_new();
    for (int i = 0; i < NSUB; ++i) {
        char id[32];
        snprintf(id, sizeof id, "%s-%03d", repeat ? "repeat" : "short", i);
        pcu_subject_builder* b = pharmsol_subject_builder_new(id);
        const double off = i * 0.01;
        if (repeat) {
            for (int k = 0; k < 10; ++k) pharmsol_subject_builder_bolus(b, 12.0 * k, 100.0, "iv");
            for (int k = 0; k < 14; ++k) pharmsol_subject_builder_observation(b, kRepeatT[k], kRepeatY[k] + off, "plasma");
        } else {
            pharmsol_subject_builder_bolus(b, 0.0, 100.0, "po");
            for (int k = 0; k < 9; ++k) pharmsol_subject_builder_observation(b, kShortT[k], kShortY[k] + off, "plasma");
        }
        pcu_subject* s = pharmsol_subject_builder_build(b);
        pharmsol_data_add_subject(d, s);
        pharmsol_subject_free(s);
    }
    return d;
}

static int run_case(pcu_ctx* ctx, const char* label, const char* dsl, int repeat, int ode) {
    pcu_model* m = NULL;
    CHECK(pharmsol_cuda_model_from_dsl(ctx, dsl, strlen(dsl), &m));
    if (ode) CHECK(pharmsol_cuda_model_set_solver(m, PCU_SOLVER_DOPRI5, 1e-4, 1e-4));
    const int np = pharmsol_cuda_model_nparams(m);
    pcu_data* d = make_data(repeat);
    pcu_error_model em;
    memset(&em, 0, sizeof em);
    em.kind = PCU_ERRMODEL_ADDITIVE; em.factor = 0.0; em.c0 = 0.1; em.c1 = 0.1;
    pcu_population* pop = NULL;
    CHECK(pharmsol_cuda_population_create(ctx, m, d, &em, 1, &pop));
    const double base_short[3] = {1.0, 0.2, 50.0}, base_repeat[4] = {0.10, 0.05, 0.04, 50.0};
    const double* base = repeat ? base_repeat : base_short;
    double spp[NSPP * 4], psi[NSUB * NSPP];
    for (int r = 0; r < NSPP; ++r)
        for (int k = 0; k < np; ++k) {
            const double p = base[k], a = p < 0 ? -p : p;
            spp[r * np + k] = p + r * 0.001 * (a > 1e-3 ? a : 1e-3);
        }
    int32_t code = 0; int64_t pair = -1;
    for (int w = 0; w < 50; ++w) CHECK(pharmsol_cuda_log_likelihood_matrix(ctx, m, pop, spp, NSPP, np, psi, &code, &pair));
    enum { REPS = 1000 };
    static double t[REPS], t1[REPS];
    for (int r = 0; r < REPS; ++r) {
        const double t0 = now_s();
        CHECK(pharmsol_cuda_log_likelihood_matrix(ctx, m, pop, spp, NSPP, np, psi, &code, &pair));
        t[r] = now_s() - t0;
    }
    double sum = 0.0;
    for (int k = 0; k < NSUB * NSPP; ++k) sum += psi[k];
    double col[NSUB];
    for (int r = 0; r < REPS; ++r) {      /* one support point x all subjects: an optimiser's cost function */
        const double t0 = now_s();
        CHECK(pharmsol_cuda_log_likelihood_matrix(ctx, m, pop, spp, 1, np, col, &code, &pair));
        t1[r] = now_s() - t0;
    }
    qsort(t, REPS, sizeof(double), cmp_double);
    qsort(t1, REPS, sizeof(double), cmp_double);
    printf("{\"bench\": \"native/likelihood-matrix/%s\", \"caller\": \"C\", \"nsub\": %d, \"nspp\": %d, \"us_per_matrix\": %.3f, \"us_per_single_column\": %.3f, "
           "\"pairs_per_s\": %.6e, \"kernel_ms\": %.5f, \"first_error_code\": %d, \"psi_sum\": %.17g, \"psi_00\": %.17g}\n",
           label, NSUB, NSPP, t[REPS / 2] * 1e6, t1[REPS / 2] * 1e6, NSUB * NSPP / t[REPS / 2], pharmsol_cuda_last_kernel_ms(ctx), (int)code, sum, psi[0]);
    pharmsol_cuda_population_destroy(pop);
    pharmsol_data_free(d);
    pharmsol_cuda_model_destroy(m);
    return 0;
}

/* --devices a,b,...: the same call through a MULTI-DEVICE context (pharmsol_cuda_ctx_create_multi): the library splits
 * the support-point columns over the listed devices inside this one process and every device copies its slab straight
 * into the caller's matrix.  A larger case than the criterion shapes (256 subjects x `nspp` support points of the
 * 2-compartment ODE) is evaluated on device list[0] alone and on the whole list; the two matrices must be identical
 * bit for bit.  A device may be listed twice (two column shards on one GPU). */
static int run_multi(const int32_t* devs, int ndev, int64_t nspp) {
    enum { NS = 256 };
    pcu_ctx *one = NULL, *many = NULL;
    CHECK(pharmsol_cuda_ctx_create(devs[0], &one));
    CHECK(pharmsol_cuda_ctx_create_multi(devs, ndev, &many));
    pcu_model* m = NULL;
    CHECK(pharmsol_cuda_model_from_dsl(one, kRepeatOde, strlen(kRepeatOde), &m));
    CHECK(pharmsol_cuda_model_set_solver(m, PCU_SOLVER_DOPRI5, 1e-6, 1e-6));
    pcu_data* d = pharmsol_data_new();
    for (int i = 0; i < NS; ++i) {
        char id[32];
        snprintf(id, sizeof id, "multi-%03d", i);
        pcu_subject_builder* b = pharmsol_subject_builder_new(id);
        for (int k = 0; k < 10; ++k) pharmsol_subject_builder_bolus(b, 12.0 * k, 100.0 + 0.1 * i, "iv");
        for (int k = 0; k < 14; ++k) pharmsol_subject_builder_observation(b, kRepeatT[k], kRepeatY[k] + i * 0.001, "plasma");
        pcu_subject* s = pharmsol_subject_builder_build(b);
        pharmsol_data_add_subject(d, s);
        pharmsol_subject_free(s);
    }
    pcu_error_model em;
    memset(&em, 0, sizeof em);
    em.kind = PCU_ERRMODEL_ADDITIVE; em.c0 = 0.1; em.c1 = 0.1;
    pcu_population *pop1 = NULL, *popn = NULL;
    CHECK(pharmsol_cuda_population_create(one, m, d, &em, 1, &pop1));
    CHECK(pharmsol_cuda_population_create(many, m, d, &em, 1, &popn));
    double* spp = malloc((size_t)nspp * 4 * sizeof(double));      /* pageable, like a caller's own arrays */
    double* a = malloc((size_t)NS * nspp * sizeof(double));
    double* b2 = malloc((size_t)NS * nspp * sizeof(double));
    if (!spp || !a || !b2) return 1;
    const double base[4] = {0.10, 0.05, 0.04, 50.0};
    for (int64_t r = 0; r < nspp; ++r)
        for (int k = 0; k < 4; ++k) spp[r * 4 + k] = base[k] * (1.0 + 0.9 * (double)((r * 2654435761u + k * 40503u) % 1000) / 1000.0);
    int32_t code = 0; int64_t pair = -1;
    double t1 = 1e30, tn = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        double t0 = now_s();
        CHECK(pharmsol_cuda_log_likelihood_matrix(one, m, pop1, spp, nspp, 4, a, &code, &pair));
        if (rep && now_s() - t0 < t1) t1 = now_s() - t0;
        t0 = now_s();
        CHECK(pharmsol_cuda_log_likelihood_matrix(many, m, popn, spp, nspp, 4, b2, &code, &pair));
        if (rep && now_s() - t0 < tn) tn = now_s() - t0;
    }
    const int same = memcmp(a, b2, (size_t)NS * nspp * sizeof(double)) == 0;
    /* the whole psi resident on every device, gathered over NVLink by the copy engines */
    double* dev_out[64];
    double t0 = now_s();
    CHECK(pharmsol_cuda_log_likelihood_matrix_replicated(many, m, popn, spp, nspp, 4, PCU_GATHER_COPY_ENGINE, dev_out, &code, &pair));
    const double trep = now_s() - t0;
    int all_ptrs = 1;
    for (int k = 0; k < ndev; ++k) all_ptrs = all_ptrs && dev_out[k] != NULL;
    printf("{\"bench\": \"native/likelihood-matrix/multi-device\", \"caller\": \"C\", \"devices\": %d, \"nsub\": %d, \"nspp\": %lld, "
           "\"ms_single_device\": %.4f, \"ms_multi_device\": %.4f, \"ms_replicated\": %.4f, \"matches_single_device\": %s, \"replicated_ptrs\": %s, "
           "\"first_error_code\": %d, \"launches\": %lld}\n",
           pharmsol_cuda_ctx_num_devices(many), NS, (long long)nspp, t1 * 1e3, tn * 1e3, trep * 1e3, same ? "true" : "false", all_ptrs ? "true" : "false", (int)code,
           (long long)pharmsol_cuda_launch_count(many));
    free(spp); free(a); free(b2);
    pharmsol_cuda_population_destroy(pop1);
    pharmsol_cuda_population_destroy(popn);
    pharmsol_data_free(d);
    pharmsol_cuda_model_destroy(m);
    pharmsol_cuda_ctx_destroy(many);
    pharmsol_cuda_ctx_destroy(one);
    return same && all_ptrs ? 0 : 1;
}

int main(int argc, char** argv) {
    int32_t ndev = 0;
    if (pharmsol_cuda_device_count(&ndev) != PCU_OK || ndev < 1) { fprintf(stderr, "no CUDA device: %s\n", pharmsol_cuda_last_error_message()); return 2; }
    if (argc >= 3 && strcmp(argv[1], "--devices") == 0) {
        int32_t devs[64];
        int n = 0;
        for (char* tok = strtok(argv[2], ","); tok && n < 64; tok = strtok(NULL, ",")) devs[n++] = (int32_t)atoi(tok);
        const int64_t nspp = argc >= 4 ? atoll(argv[3]) : 8192;
        return n > 0 ? run_multi(devs, n, nspp) : 2;
    }
    pcu_ctx* ctx = NULL;
    CHECK(pharmsol_cuda_ctx_create(0, &ctx));
    int rc = 0;
    rc |= run_case(ctx, "1cpt-12h-po/analytical", kShortAnalytical, 0, 0);
    rc |= run_case(ctx, "1cpt-12h-po/ode", kShortOde, 0, 1);
    rc |= run_case(ctx, "2cpt-120h-q12h/analytical", kRepeatAnalytical, 1, 0);
    rc |= run_case(ctx, "2cpt-120h-q12h/ode", kRepeatOde, 1, 1);
    pharmsol_cuda_ctx_destroy(ctx);
    return rc;
}
