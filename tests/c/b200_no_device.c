/* TEST-ONLY: the thin CUDA layer's entry points that tests/c/ref_driver.c and tests/c/petsc_mini use for
 * device-resident Vecs, as refusals.  Linked into oracle/_ref/ref_driver_cpu (the reference's host code on the CPU
 * restatement of libCEED, oracle/ceed_cpu.c), which runs with -memtype host only and therefore never calls them. */
#include <stddef.h>

#include "b200_kernels.h"

#define NO_DEVICE { return 1; }
const char *b200_last_error(void) { return "ref_driver_cpu has no device (host memtype only)"; }
int b200_malloc(void **p, size_t bytes) { (void)bytes; *p = NULL; return 1; }
int b200_free(void *p) { (void)p; return 1; }
int b200_memset(void *p, int value, size_t bytes) { (void)p; (void)value; (void)bytes; return 1; }
int b200_memcpy_h2d(void *dst, const void *src, size_t bytes) { (void)dst; (void)src; (void)bytes; return 1; }
int b200_memcpy_d2h(void *dst, const void *src, size_t bytes) { (void)dst; (void)src; (void)bytes; return 1; }
int b200_sync(void) NO_DEVICE
int b200_vec_reciprocal(double *d, size_t n) { (void)d; (void)n; return 1; }
int b200_vec_axpy(double *y, double alpha, const double *x, size_t n) { (void)y; (void)alpha; (void)x; (void)n; return 1; }
int b200_vec_pointwise_mult(double *w, const double *x, const double *y, size_t n) { (void)w; (void)x; (void)y; (void)n; return 1; }
int b200_gather(double *dst, const double *src, const int *idx, size_t n) { (void)dst; (void)src; (void)idx; (void)n; return 1; }
int b200_scatter_set(double *dst, const int *idx, const double *src, size_t n) { (void)dst; (void)idx; (void)src; (void)n; return 1; }
