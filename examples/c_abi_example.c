/* Plain-C use of the drop-in boundary (include/garpack_b200.h): eigs(A, 3; ncv=12) on a small 1-D Laplacian.
 * Build:  gcc -std=c99 -Iinclude examples/c_abi_example.c -Lgenericarpack.jl_b200 -lgarpack_b200 \
 *             -Wl,-rpath,$PWD/genericarpack.jl_b200 -o c_abi_example
 * Without a CUDA device ga_create returns GA_ERR_CUDA and the program says so: there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "garpack_b200.h"

int main(void) {
  enum { N = 200, NEV = 3, NCV = 12 };
  int64_t rowptr[N + 1];
  int32_t* col = (int32_t*)malloc(sizeof(int32_t) * 3 * N);
  double* val = (double*)malloc(sizeof(double) * 3 * N);
  int64_t k = 0;
  int i;
  for (i = 0; i < N; ++i) { /* full CSR (both triangles) of tridiag(-1, 2, -1) */
    rowptr[i] = k;
    if (i > 0) { col[k] = i - 1; val[k++] = -1.0; }
    col[k] = i; val[k++] = 2.0;
    if (i < N - 1) { col[k] = i + 1; val[k++] = -1.0; }
  }
  rowptr[N] = k;

  ga_ctx* ctx = NULL;
  int rc = ga_create(&ctx, GA_F64, N, N, 0, NCV, /*device*/ 0, /*rank*/ 0, /*world*/ 1);
  if (rc != 0) {
    printf("ga_create returned %d (%s): no device, no result\n", rc, rc == GA_ERR_CUDA ? "GA_ERR_CUDA" : "error");
    free(col); free(val);
    return rc == GA_ERR_CUDA ? 0 : 1;
  }
  rc = ga_set_csr(ctx, rowptr, col, val);
  double values[NEV], bounds[NEV];
  double* Z = (double*)malloc(sizeof(double) * N * NEV);
  int64_t iparam[11];
  if (rc == 0) rc = ga_symeigs(ctx, GA_LM, NEV, /*tol*/ NULL, /*maxiter*/ 300, /*v0*/ NULL, /*ritzvec*/ 1, values, bounds, Z, N, iparam);
  if (rc != 0) printf("error %d: %s\n", rc, ga_last_error(ctx));
  else for (i = 0; i < (int)iparam[4]; ++i) printf("lambda[%d] = %.15g  (bound %.2e)\n", i, values[i], bounds[i]);
  ga_destroy(ctx);
  free(col); free(val); free(Z);
  return rc;
}
