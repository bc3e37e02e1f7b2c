/* test_sharded.c — plain C driver of the multi-GPU entry points (pdenlp_sharded_*), through include/pdenlp_cuda.h only.
 *
 *   gcc -O2 -I include tests/c/test_sharded.c -L pdenlpmodels.jl_b200/csrc -lpdenlp_cuda -lm -o test_sharded
 *   ./test_sharded tests/golden/desc_membrane8.bin 2
 *
 * Reads a flattened model (written by pdenlp_b200.capi.dump_desc: a C program has no Gridap host layer), creates the
 * single-device model (PDENLP_MEM_HOST) and the sharded model over `ndev` GPUs and requires every callback of the
 * sharded handle to reproduce the single-device result: structures and FE segments of the COO streams bit for bit,
 * sums over cells (dof-indexed vectors, kappa blocks, obj) to 1e-12 of the array's scale.
 * The reference has no multi-GPU path; this replaces nothing in it (new scaling axis). */
#include "pdenlp_cuda.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(call)                                                                              \
  do {                                                                                        \
    int rc__ = (call);                                                                        \
    if (rc__ != PDENLP_OK) {                                                                  \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc__, pdenlp_last_error());              \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

static void* rd(FILE* fh, size_t n, size_t sz) {
  void* p = malloc(n * sz + 16);
  if (n && fread(p, sz, n, fh) != n) { fprintf(stderr, "short read\n"); exit(3); }
  return p;
}
static unsigned long long rng_state = 88172645463325252ull;
static double rnd(void) {  /* xorshift64, uniform in (-1, 1) */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return 2.0 * ((double)(rng_state >> 11) / 9007199254740992.0) - 1.0;
}
static int cmp(const char* what, const double* a, const double* b, long long n, double tol) {
  double md = 0, mb = 0; long long same = 0;
  for (long long i = 0; i < n; ++i) {
    double d = fabs(a[i] - b[i]);
    if (!(d <= md)) md = d;   /* NaN sticks */
    if (fabs(b[i]) > mb) mb = fabs(b[i]);
    same += a[i] == b[i];
  }
  const int ok = md <= tol * (mb > 0 ? mb : 1.0);
  printf("  %-14s n=%-9lld max|diff|=%.3e scale=%.3e bitwise-equal=%lld %s\n", what, n, md, mb, same, ok ? "ok" : "MISMATCH");
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s desc.bin ndev\n", argv[0]); return 64; }
  const int ndev = atoi(argv[2]);
  FILE* fh = fopen(argv[1], "rb");
  if (!fh) { perror(argv[1]); return 3; }
  char magic[8];
  if (fread(magic, 1, 8, fh) != 8 || memcmp(magic, "PDENLPD3", 8)) { fprintf(stderr, "bad magic\n"); return 3; }
  int32_t* i32 = rd(fh, 10, 4);
  int64_t* i64 = rd(fh, 5, 8);
  double* par = rd(fh, 8, 8);
  pdenlp_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = PDENLP_ABI_VERSION;
  d.integrand = i32[0]; d.dim = i32[1]; d.ny = i32[2]; d.nu = i32[3]; d.nq = i32[4]; d.nparam = i32[5];
  d.nfields_y = i32[6]; d.nfields_u = i32[7]; d.ncoef = i32[8]; d.geom_nq = i32[9];
  d.ncells = i64[0]; d.nfree_y = i64[1]; d.nfree_u = i64[2]; d.ndir_y = i64[3]; d.ndir_u = i64[4];
  memcpy(d.params, par, sizeof d.params);
  d.memspace = PDENLP_MEM_HOST; d.device = 0;
  const int nfy = d.nfields_y > 0 ? d.nfields_y : 1, nfu = d.nu ? (d.nfields_u > 0 ? d.nfields_u : 1) : 0;
  d.cell_dofs_y = rd(fh, (size_t)d.ncells * nfy * d.ny, 4);
  d.cell_dofs_u = rd(fh, (size_t)d.ncells * nfu * d.nu, 4);
  d.dir_y = rd(fh, (size_t)d.ndir_y, 8); d.dir_u = rd(fh, (size_t)d.ndir_u, 8);
  d.phi_y = rd(fh, (size_t)d.nq * d.ny, 8); d.grad_y = rd(fh, (size_t)d.nq * d.ny * d.dim, 8);
  d.phi_u = rd(fh, (size_t)d.nq * d.nu, 8); d.wdet = rd(fh, (size_t)d.nq, 8);
  d.coef = rd(fh, (size_t)d.ncoef * d.ncells * d.nq, 8);
  if (!d.ncoef) d.coef = NULL;
  d.field_nfree = rd(fh, (size_t)(nfy + nfu), 8); d.field_ndir = rd(fh, (size_t)(nfy + nfu), 8);
  if (d.geom_nq) {
    d.cell_jinvT = rd(fh, (size_t)d.ncells * d.geom_nq * d.dim * d.dim, 8);
    d.cell_detj = rd(fh, (size_t)d.ncells * d.geom_nq, 8);
  }
  fclose(fh);

  pdenlp_model* m = NULL;
  pdenlp_sharded* s = NULL;
  CK(pdenlp_create(&d, &m));
  CK(pdenlp_sharded_create(&d, ndev, NULL, &s));
  pdenlp_info I, G;
  CK(pdenlp_get_info(m, &I));
  CK(pdenlp_sharded_get_info(s, &G));
  printf("integrand %d, %lld cells on %d device(s): nvar %lld ncon %lld nnzj %lld nnzh %lld\n", d.integrand, (long long)d.ncells, ndev,
         (long long)I.nvar, (long long)I.ncon, (long long)I.nnzj, (long long)I.nnzh);
  for (int k = 0; k < ndev; ++k) {
    int64_t c0, c1;
    CK(pdenlp_sharded_slab(s, k, &c0, &c1));
    printf("  slab %d: cells [%lld, %lld)\n", k, (long long)c0, (long long)c1);
  }
  if (I.nvar != G.nvar || I.ncon != G.ncon || I.nnzj != G.nnzj || I.nnzh != G.nnzh || I.nnzh_obj != G.nnzh_obj) {
    fprintf(stderr, "global sizes differ\n");
    return 1;
  }
  const long long n = I.nvar, nc = I.ncon;
  double *x, *lam, *v;  /* page-locked: the devices' copies overlap */
  CK(pdenlp_host_alloc((void**)&x, n * 8)); CK(pdenlp_host_alloc((void**)&lam, (nc + 1) * 8)); CK(pdenlp_host_alloc((void**)&v, n * 8));
  for (long long i = 0; i < n; ++i) { x[i] = rnd(); v[i] = rnd(); }
  for (int k = 0; k < d.nparam; ++k) x[k] = 1.0 + 0.1 * rnd();
  for (long long i = 0; i < nc; ++i) lam[i] = rnd();
  const long long big = (I.nnzh > I.nnzj ? I.nnzh : I.nnzj) + n + 2;
  double* a = malloc(big * 8); double* b = malloc(big * 8);
  int64_t *r1 = malloc(big * 8), *c1_ = malloc(big * 8), *r2 = malloc(big * 8), *c2 = malloc(big * 8);
  int bad = 0;
  const double tol = 1e-12;

  double f1 = 0, f2 = 0;
  CK(pdenlp_obj(m, x, &f1)); CK(pdenlp_sharded_obj(s, x, &f2));
  bad += cmp("obj", &f2, &f1, 1, tol);
  CK(pdenlp_grad(m, x, a)); CK(pdenlp_sharded_grad(s, x, b)); bad += cmp("grad!", b, a, n, tol);
  if (nc > 0) {
    CK(pdenlp_cons(m, x, a)); CK(pdenlp_sharded_cons(s, x, b)); bad += cmp("cons!", b, a, nc, tol);
    CK(pdenlp_jprod(m, x, v, a)); CK(pdenlp_sharded_jprod(s, x, v, b)); bad += cmp("jprod!", b, a, nc, tol);
    CK(pdenlp_jtprod(m, x, lam, a)); CK(pdenlp_sharded_jtprod(s, x, lam, b)); bad += cmp("jtprod!", b, a, n, tol);
    CK(pdenlp_jac_coord(m, x, a)); CK(pdenlp_sharded_jac_coord(s, x, b)); bad += cmp("jac_coord!", b, a, I.nnzj, tol);
    /* the FE segments are computed cell by cell, identically on one or many devices: bit for bit */
    bad += memcmp(a + I.nnzj_k, b + I.nnzj_k, (size_t)(I.nnzj - I.nnzj_k) * 8) != 0;
    CK(pdenlp_jac_structure(m, r1, c1_)); CK(pdenlp_sharded_jac_structure(s, r2, c2));
    bad += memcmp(r1, r2, (size_t)I.nnzj * 8) != 0 || memcmp(c1_, c2, (size_t)I.nnzj * 8) != 0;
    printf("  jac structure %s\n", bad ? "?" : "identical");
  }
  CK(pdenlp_hprod(m, x, nc > 0 ? lam : NULL, 0.37, v, a)); CK(pdenlp_sharded_hprod(s, x, nc > 0 ? lam : NULL, 0.37, v, b)); bad += cmp("hprod!", b, a, n, tol);
  CK(pdenlp_hess_coord(m, x, nc > 0 ? lam : NULL, 0.37, a)); CK(pdenlp_sharded_hess_coord(s, x, nc > 0 ? lam : NULL, 0.37, b)); bad += cmp("hess_coord!", b, a, I.nnzh, tol);
  CK(pdenlp_hess_coord(m, x, NULL, 1.0, a)); CK(pdenlp_sharded_hess_coord(s, x, NULL, 1.0, b)); bad += cmp("hess_coord!(x)", b, a, I.nnzh, tol);
  CK(pdenlp_hess_structure(m, r1, c1_)); CK(pdenlp_sharded_hess_structure(s, r2, c2));
  bad += memcmp(r1, r2, (size_t)I.nnzh * 8) != 0 || memcmp(c1_, c2, (size_t)I.nnzh * 8) != 0;
  printf("  hess structure %s\n", bad ? "?" : "identical");

  CK(pdenlp_sharded_destroy(s));
  CK(pdenlp_destroy(m));
  pdenlp_host_free(x); pdenlp_host_free(lam); pdenlp_host_free(v);
  if (bad) { printf("SHARDED_C_FAILED (%d mismatches)\n", bad); return 1; }
  printf("SHARDED_C_OK ndev=%d\n", ndev);
  return 0;
}
