/*
 * cpu_driver.c -- TEST INFRASTRUCTURE ONLY (CPU baseline leg of bench.py).
 *
 * Drives a push+deposit kernel with the reference's own threading model so that the CPU
 * baseline is timed the way OSIRIS runs it:
 *   - T threads (pthreads here; OpenMP in the reference), each with a private J grid, disjoint particle chunks whose size is
 *     rounded up to p_vecwidth = 16 (spec/os-spec-vpush.f03:380-397);
 *   - each thread zeroes its J grid first (particles/os-particles.f03:1423-1431);
 *   - then J grids 2..T are added into grid 1 (vdf/os-vdf-math.f90:389-474, reduce_vdf).
 * The kernel is passed as a function pointer with the reference C ABI
 * (simd/os-spec-push-avx2.c:1141-1145), so it can be the compiled reference (oracle/_ref)
 * or any ABI-compatible stand-in.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef void (*push_fn)(int *ix, void *x, void *u, void *q, int *i0, int *i1, void *rqm,
                        void *e, void *b, int *esz, int *eoff, void *j, int *jsz, int *joff,
                        double *dx, double *dt, double *ene);

int cpu_driver_max_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

typedef struct {
  push_fn fn;
  int tid, nthreads, real_bytes;
  int *ix; void *x, *u, *q; long np, chunk;
  void *rq, *e, *b; int *esz, *eoff; char *jthreads; long jcount; int *jsz, *joff;
  double *dx, dt, energy;
  int phase; /* 0 = push, 1 = reduce */
} job_t;

static void *worker(void *arg) {
  job_t *w = (job_t *)arg;
  size_t rb = (size_t)w->real_bytes;
  if (w->phase == 0) {
    char *jt = w->jthreads + (size_t)w->tid * (size_t)w->jcount * rb;
    memset(jt, 0, (size_t)w->jcount * rb);
    long a = (w->nthreads > 1) ? w->tid * w->chunk + 1 : 1;
    long z = (w->nthreads > 1) ? (w->tid + 1) * w->chunk : w->np;
    if (z > w->np) z = w->np;
    w->energy = 0.0;
    if (z >= a) {
      int i0 = (int)a, i1 = (int)z;
      double dtl = w->dt;
      w->fn(w->ix, w->x, w->u, w->q, &i0, &i1, w->rq, w->e, w->b, w->esz, w->eoff, jt, w->jsz,
            w->joff, w->dx, &dtl, &w->energy);
    }
  } else {
    /* reduce_vdf: pf(1) += pf(t), t = 2..T; each thread owns a slice of the grid */
    long lo = w->jcount * w->tid / w->nthreads, hi = w->jcount * (w->tid + 1) / w->nthreads;
    for (int t = 1; t < w->nthreads; t++) {
      if (w->real_bytes == 4) {
        float *d = (float *)w->jthreads, *s = d + (size_t)t * w->jcount;
        for (long i = lo; i < hi; i++) d[i] += s[i];
      } else {
        double *d = (double *)w->jthreads, *s = d + (size_t)t * w->jcount;
        for (long i = lo; i < hi; i++) d[i] += s[i];
      }
    }
  }
  return NULL;
}

/* jthreads: nthreads consecutive J grids of jcount reals each; result in the first. */
double cpu_driver_advance_deposit(push_fn fn, int nthreads, int real_bytes, int dim, int *ix,
                                  void *x, void *u, void *q, long np, double rqm, void *e,
                                  void *b, int *esz, int *eoff, void *jthreads, long jcount,
                                  int *jsz, int *joff, double *dx, double dt) {
  const int vecw = 16;
  long chunk = (np + nthreads - 1) / nthreads;
  chunk = ((chunk + vecw - 1) / vecw) * vecw;
  float rqm_f = (float)rqm;
  double rqm_d = rqm;
  void *rq = (real_bytes == 4) ? (void *)&rqm_f : (void *)&rqm_d;
  (void)dim;
  job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  double energy = 0.0;
  for (int phase = 0; phase < 2; phase++) {
    if (phase == 1 && nthreads == 1) break;
    for (int t = 0; t < nthreads; t++) {
      job_t w = {fn, t, nthreads, real_bytes, ix, x, u, q, np, chunk, rq, e, b, esz, eoff,
                 (char *)jthreads, jcount, jsz, joff, dx, dt, 0.0, phase};
      if (phase == 1) w.energy = jobs[t].energy;
      jobs[t] = w;
      pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  }
  for (int t = 0; t < nthreads; t++) energy += jobs[t].energy;
  free(jobs);
  free(th);
  return energy;
}
