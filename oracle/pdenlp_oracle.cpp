// pdenlp_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A CPU restatement of the algorithm PDENLPModels.jl v0.3.5 runs for the
// NLPModels callbacks of a GridapPDENLPModel.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the
// product path (libpdenlp_cuda.so) never does.
//
// PARITY PINNING: the reference is pure Julia on top of Gridap 0.15.5 +
// ForwardDiff, none of which exists in this image, so the oracle cannot be
// checked against outputs of the reference itself.  It is pinned by (i) the
// reference's own unit-test vectors for the dof numbering
// (test/unit-test.jl:828,846,855-874), (ii) its known-answer tests
// (test/problems/poissonparam.jl:54-63 etc.), (iii) its finite-difference /
// coord-vs-product self-consistency checks re-expressed in tests/.  There are no
// golden vectors for jac_coord!/hess_coord! values in the reference, so for
// those values parity is "unpinned"  — see DESIGN.md.
//
// What is restated, and where it comes from (paths relative to /root/reference):
//   * per-cell autodiff exactly as Gridap/ForwardDiff do it: the cell-local dof
//     vector z_e = [y_e ; u_e] (Dirichlet entries included) is seeded with
//     full-width duals (Chunk{length(x)}, src/test_autodiff.jl:8,26) and pushed
//     through  sum_q w_q detJ integrand(x_q);  Hessians use nested duals
//     (ForwardDiff.hessian = jacobian of gradient).
//   * obj:   src/GridapPDENLPModel.jl:245-255, src/additional_obj_terms.jl:135-143,234-242
//   * grad!: src/additional_obj_terms.jl:145-163,244-284 (assemble_vector!: zero-fill, add in cell order)
//   * cons:  src/jacobian_functions.jl:23-58
//   * jac:   src/jacobian_functions.jl:122-206 (segments [k | y-block | u-block];
//            loop order of Gridap's fill_matrix_coo_numeric! = the loop of
//            src/fill_matrix_hessian_functions.jl:50-67 without the <= filter)
//   * hess:  src/GridapPDENLPModel.jl:268-301,341-411, src/fill_matrix_hessian_functions.jl:50-94
//            (block order (1,1),(2,1),(1,2),(2,2); keep iff gidrow>0, gidcol>0, gidcol<=gidrow)
//   * structure: src/hessian_struct_nnzh_functions.jl:10-130, src/jacobian_functions.jl:1-7,60-120,
//            src/additional_constructors.jl:206-211,259-260
//   * products: NLPModels coo_prod!/coo_sym_prod! as used at src/GridapPDENLPModel.jl:321,428,485,517
//   * kappa blocks: the reference differentiates the whole assembly w.r.t. kappa
//     (src/jacobian_functions.jl:136-142, src/GridapPDENLPModel.jl:361-380,
//     src/additional_obj_terms.jl:286-314); here kappa is appended to the cell-local
//     dual vector and the per-cell partials are summed in cell order — the same sums.
//   * Gridap discretisation conventions (third party, not in /root/reference;
//     Gridap = "0.15.5", Project.toml:17): SURVEY Appendix A.3.
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp -shared -fPIC).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// ------------------------------------------------------------------ duals
template <class T, int N>
struct Dual {
  T v;
  T d[N];
};

template <class T> struct is_dual { static const bool value = false; };
template <class T, int N> struct is_dual<Dual<T, N>> { static const bool value = true; };

static inline double cst(double a, const double&) { return a; }
template <class T, int N>
static inline Dual<T, N> cst(double a, const Dual<T, N>& like) {
  Dual<T, N> r;
  r.v = cst(a, like.v);
  for (int i = 0; i < N; ++i) r.d[i] = cst(0.0, like.v);
  return r;
}

template <class T, int N> static inline Dual<T, N> operator+(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> operator-(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> operator*(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> operator*(double a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> operator*(const Dual<T, N>& b, double a) { return a * b; }
template <class T, int N> static inline Dual<T, N> operator+(const Dual<T, N>& a, double b) {
  Dual<T, N> r = a; r.v = a.v + b; return r; }
template <class T, int N> static inline Dual<T, N> operator+(double b, const Dual<T, N>& a) { return a + b; }
template <class T, int N> static inline Dual<T, N> operator-(const Dual<T, N>& a, double b) {
  Dual<T, N> r = a; r.v = a.v - b; return r; }
template <class T, int N> static inline Dual<T, N> operator-(double b, const Dual<T, N>& a) { return (-a) + b; }

static inline double xcos(double a) { return std::cos(a); }
static inline double xsin(double a) { return std::sin(a); }
template <class T, int N> static inline Dual<T, N> xsin(const Dual<T, N>& a);
template <class T, int N> static inline Dual<T, N> xcos(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = xcos(a.v); T s = xsin(a.v); for (int i = 0; i < N; ++i) r.d[i] = (-s) * a.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> xsin(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = xsin(a.v); T c = xcos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
static inline double xsinh(double a) { return std::sinh(a); }
static inline double xcosh(double a) { return std::cosh(a); }
template <class T, int N> static inline Dual<T, N> xcosh(const Dual<T, N>& a);
template <class T, int N> static inline Dual<T, N> xsinh(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = xsinh(a.v); T c = xcosh(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <class T, int N> static inline Dual<T, N> xcosh(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = xcosh(a.v); T s = xsinh(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }

// ------------------------------------------------------------------ long scalar sums
// The reference's scalar reductions over ALL cells are `sum(::DomainContribution)` (src/additional_obj_terms.jl:135-143,
// 234-242: obj, and through ForwardDiff the kappa entries of grad! and the kappa-kappa entries of the objective
// Hessian, :275-314).  That is Julia's generic `sum` over an indexable array: PAIRWISE summation — Base
// `mapreduce_impl(f, op, A, ifirst, ilast, blksize = 1024)` (third party: Julia Base reduce.jl, not in /root/reference):
// ranges shorter than 1024 are added left to right, longer ones are halved recursively.  Restated here on the per-cell
// values in cell order; its error grows like log(n)*eps, whereas a plain left-to-right loop over 10^6 same-sign cell
// values drifts by ~1e-12 relative — more than the parity tolerance.  (Inside a block Julia's loop is @simd, so even
// the reference's bits depend on the SIMD width; any summation of pairwise accuracy agrees with it to ~1e-15.)
static double pairwise_sum(const double* a, int64_t ifirst, int64_t ilast /* inclusive */) {
  if (ilast < ifirst) return 0.0;
  if (ifirst == ilast) return a[ifirst];
  if (ilast - ifirst < 1024) {
    double v = a[ifirst] + a[ifirst + 1];
    for (int64_t i = ifirst + 2; i <= ilast; ++i) v += a[i];
    return v;
  }
  const int64_t imid = ifirst + ((ilast - ifirst) >> 1);
  return pairwise_sum(a, ifirst, imid) + pairwise_sum(a, imid + 1, ilast);
}
// per-cell values of `ncomp` scalar quantities, summed pairwise over the cells
struct CellSums {
  int ncomp; int64_t ncells; std::vector<double> v;  // [comp][cell]
  CellSums(int nc, int64_t n) : ncomp(nc), ncells(n), v((size_t)(nc > 0 ? nc : 1) * (size_t)n, 0.0) {}
  double& at(int comp, int64_t cell) { return v[(size_t)comp * ncells + cell]; }
  double sum(int comp) const { return pairwise_sum(v.data() + (size_t)comp * ncells, 0, ncells - 1); }
};

// ------------------------------------------------------------------ problem
extern "C" {
typedef struct oracle_problem {
  int32_t integrand;  // same ids as include/pdenlp_cuda.h
  int32_t dim;
  int32_t order_y, order_u;  // order_u = 0 -> no control space (VoidFESpace)
  int32_t degree;            // Measure(trian, degree)
  int32_t nparam;
  int32_t n[3];
  int32_t nthreads;          // 0 = all OpenMP threads
  double lo[3], hi[3];
  int64_t ncells;            // cells of this slab
  int64_t cell0;             // global index of the first cell (x-fastest numbering)
  int64_t nfree_y, nfree_u;
  const int32_t* cell_dofs_y;  // [ncells][ny]
  const int32_t* cell_dofs_u;  // [ncells][nu]
  const double* dir_y;
  const double* dir_u;
  double params[8];
  // multi-field layout (nfy = 0 or 1: single-field legacy): nfy state fields x ny local dofs, nfu control
  // fields x nu; cell_dofs_y is [ncells][nfy*ny] field-major with per-field ids; field_nfree / field_ndir list
  // the state fields first, then the control fields; dir_y / dir_u are the concatenated Dirichlet values
  int32_t nfy, nfu;
  int64_t field_nfree[8];
  int64_t field_ndir[8];
  // 1: reproduce the reference's hess_coord!(x, lambda) literally: its FE constraint block is evaluated at
  // FEFunction(Y, x) with the UN-shifted x (src/GridapPDENLPModel.jl:384), i.e. the y dofs are read from
  // x[1:nfree_y] and the u dofs from x[nfree_y+1:...] although kappa occupies x[1:nparam]
  int32_t compat_shifted_x;
  int32_t reserved;
  // Mapped mesh — Gridap's CartesianDiscreteModel(domain, partition; map = f): the coordinates of ALL mesh vertices
  // (x-fastest vertex numbering, [nverts][dim]); NULL = the plain Cartesian mesh.  Every cell is then the Q1
  // (multi-linear) image of the reference cell spanned by its mapped vertices: x(xi) = sum_v N_v(xi) x_v, with
  // Jacobian J(xi) = sum_v x_v (x) grad N_v(xi); physical gradients are J^-T grad_xi and the measure w_q |det J(xi_q)|.
  const double* vertex_coords;
} oracle_problem;
}

enum { MEMBRANE = 1, PB = 2, BURGER_LIN = 3, BURGER_NL = 4, KAPPA_POISSON = 5, PENALIZED_POISSON = 6, BASIC_UNCONSTRAINED = 7,
       CONTROL_SIR = 8, CELL_INCREASE = 9, DYNAMIC_SIR = 10, TORE_BRACHISTOCHRONE = 11, POISSON_PARAM = 12, SIS = 13, KAPPA_POISSON2 = 14 };

// NoFETerm objectives (src/additional_obj_terms.jl:59-106): f(kappa) without any integral; the objective has no
// FE Hessian block (hessian_struct_nnzh_functions.jl:41-43) and its kappa-block is the p x p lower triangle
static bool is_nofe(const oracle_problem& P) { return P.integrand == POISSON_PARAM || P.integrand == SIS; }
// fk(k) = 0.5 * dot(k .- 1, k .- 1) (poissonparam.jl:43-44); SIS: NoFETerm() = 0
template <class S>
static S nofe_obj(const oracle_problem& P, const S* k, int np) {
  S a = cst(0.0, k[0]);
  if (P.integrand == POISSON_PARAM) for (int i = 0; i < np; ++i) { S d = k[i] - 1.0; a = a + 0.5 * d * d; }
  return a;
}

static const double PI = 3.14159265358979323846;

// coefficient functions, as written in the reference problem files
static double coef_target(const oracle_problem& P, const double* x) {
  switch (P.integrand) {
    case MEMBRANE: return -x[0] * x[0];           // yd, README.md:82
    case PB: {                                    // docs/src/poisson-boltzman.md:54
      double m = std::fmin(std::fmin(x[0] - 0.25, 0.75 - x[0]), std::fmin(x[1] - 0.25, 0.75 - x[1]));
      return m >= 0. ? 10. : 5.;
    }
    case BURGER_LIN:
    case BURGER_NL: return -x[0] * x[0];          // test/problems/burger1d.jl:33
    case KAPPA_POISSON2:
    case KAPPA_POISSON: {                         // sol, test/problems/poissonmixed.jl:20
      double s = std::sin(2 * PI * x[0]) * x[1];
      return P.dim == 3 ? s * x[2] : s;
    }
    case BASIC_UNCONSTRAINED: return x[0] * x[0] + x[1] * x[1];  // ubis, basicunconstrained.jl:20
    case DYNAMIC_SIR: return 1.0 + x[0];                         // w0, dynamicsir.jl:50
  }
  return 0.0;
}
static double coef_source(const oracle_problem& P, const double* x) {
  switch (P.integrand) {
    case MEMBRANE:
    case PB: {                                    // h, README.md:90-91
      double w = PI - 1.0 / 8.0;
      return -std::sin(w * x[0]) * std::sin(w * x[1]);
    }
    case BURGER_LIN:
    case BURGER_NL: return 2 * (P.params[1] + x[0] * x[0] * x[0]);  // burger1d.jl:24
    case PENALIZED_POISSON: return 1.0;           // w, penalizedpoisson.jl:32
    case KAPPA_POISSON2:
    case POISSON_PARAM:                           // f, poissonparam.jl:33
    case KAPPA_POISSON: {                         // f, poissonmixed.jl:31
      double s = (2 * PI * PI) * std::sin(2 * PI * x[0]) * x[1];
      return P.dim == 3 ? s * x[2] : s;
    }
  }
  return 0.0;
}

// ------------------------------------------------------------------ reference FE
// 1-D nodal Lagrange basis on [0,1]; node label 0: xi=0, 1: xi=1, 2: xi=1/2
static void lag1d(int order, double xi, double* v, double* g) {
  if (order == 1) {
    v[0] = 1 - xi; v[1] = xi; g[0] = -1; g[1] = 1;
  } else {
    v[0] = (1 - xi) * (1 - 2 * xi); v[1] = xi * (2 * xi - 1); v[2] = 4 * xi * (1 - xi);
    g[0] = 4 * xi - 3; g[1] = 4 * xi - 1; g[2] = 4 - 8 * xi;
  }
}
// local dof -> per-direction node label, Gridap order: vertices, edges, interior
static void node_label(int order, int dim, int a, int* lab) {
  if (order == 1) { for (int d = 0; d < dim; ++d) lab[d] = (a >> d) & 1; return; }
  if (dim == 1) { lab[0] = a; return; }
  static const int L2[9][2] = {{0,0},{1,0},{0,1},{1,1},{2,0},{2,1},{0,2},{1,2},{2,2}};
  lab[0] = L2[a][0]; lab[1] = L2[a][1];
}
static int nloc_of(int order, int dim) { int n = 1; for (int d = 0; d < dim; ++d) n *= (order + 1); return order == 0 ? 0 : n; }

struct RefTables {
  int nq = 0, ny = 0, nu = 0, dim = 0;
  std::vector<double> xi, w;          // [nq][dim], [nq]
  std::vector<double> phiy, dphiy;    // [nq][ny], [nq][ny][dim] (reference gradients)
  std::vector<double> phiu;           // [nq][nu]
};

static void gauss01(int npd, double* x, double* w) {
  // Gauss-Legendre on [0,1]
  if (npd == 1) { x[0] = 0.5; w[0] = 1.0; }
  else if (npd == 2) { double a = 0.5 / std::sqrt(3.0); x[0] = 0.5 - a; x[1] = 0.5 + a; w[0] = w[1] = 0.5; }
  else { double a = 0.5 * std::sqrt(0.6); x[0] = 0.5 - a; x[1] = 0.5; x[2] = 0.5 + a; w[0] = w[2] = 5.0 / 18.0; w[1] = 8.0 / 18.0; }
}

static RefTables make_tables(const oracle_problem& P) {
  RefTables T;
  T.dim = P.dim;
  int npd = (P.degree + 2) / 2;  // ceil((degree+1)/2)
  if (npd > 3) npd = 3;
  double x1[3], w1[3];
  gauss01(npd, x1, w1);
  T.nq = 1; for (int d = 0; d < P.dim; ++d) T.nq *= npd;
  T.ny = nloc_of(P.order_y, P.dim);
  T.nu = nloc_of(P.order_u, P.dim);
  T.xi.resize(T.nq * P.dim); T.w.resize(T.nq);
  T.phiy.resize(T.nq * T.ny); T.dphiy.resize(T.nq * T.ny * P.dim); T.phiu.resize(T.nq * (T.nu ? T.nu : 1));
  for (int q = 0; q < T.nq; ++q) {
    int r = q; double w = 1;
    for (int d = 0; d < P.dim; ++d) { int k = r % npd; r /= npd; T.xi[q * P.dim + d] = x1[k]; w *= w1[k]; }
    T.w[q] = w;
    double v[3][3], g[3][3]; int lab[3];
    for (int d = 0; d < P.dim; ++d) lag1d(P.order_y, T.xi[q * P.dim + d], v[d], g[d]);
    for (int a = 0; a < T.ny; ++a) {
      node_label(P.order_y, P.dim, a, lab);
      double val = 1; for (int d = 0; d < P.dim; ++d) val *= v[d][lab[d]];
      T.phiy[q * T.ny + a] = val;
      for (int e = 0; e < P.dim; ++e) {
        double gg = 1; for (int d = 0; d < P.dim; ++d) gg *= (d == e) ? g[d][lab[d]] : v[d][lab[d]];
        T.dphiy[(q * T.ny + a) * P.dim + e] = gg;
      }
    }
    if (T.nu) {
      for (int d = 0; d < P.dim; ++d) lag1d(P.order_u, T.xi[q * P.dim + d], v[d], g[d]);
      for (int b = 0; b < T.nu; ++b) {
        node_label(P.order_u, P.dim, b, lab);
        double val = 1; for (int d = 0; d < P.dim; ++d) val *= v[d][lab[d]];
        T.phiu[q * T.nu + b] = val;
      }
    }
  }
  return T;
}

// ------------------------------------------------------------------ integrands
// Pointwise data handed to the integrands: y, grad y, u at the quadrature point
// (as scalars of type S), physical coordinates, the test function value/gradient.
template <class S>
struct Pt {
  S y, gy[3], u;
  S k[2];
  double yd, h;   // coef_target, coef_source at x_q
};

// objective density f, exactly as written in the reference problem files
template <class S>
static S f_density(const oracle_problem& P, const Pt<S>& p) {
  switch (P.integrand) {
    case MEMBRANE: case PB: case BURGER_LIN: {
      // 0.5 * (yd - y) * (yd - y) + 0.5 * alpha * u * u
      S d = p.yd - p.y;
      return 0.5 * d * d + (0.5 * P.params[0]) * p.u * p.u;
    }
    case BURGER_NL: {
      S d = p.yd - p.y; S t = p.k[0] - P.params[2];
      return 0.5 * d * d + (0.5 * P.params[0]) * p.u * p.u + 0.5 * t * t;
    }
    case PENALIZED_POISSON: {
      // 0.5 * grad(y) . grad(y) - w * y   (penalizedpoisson.jl:33-35)
      S gg = cst(0.0, p.y);
      for (int d = 0; d < P.dim; ++d) gg = gg + p.gy[d] * p.gy[d];
      return 0.5 * gg - p.h * p.y;
    }
    case BASIC_UNCONSTRAINED: {
      // 0.5 * (ubis - u) * (ubis - u) + 0.5 * y * y   (basicunconstrained.jl:22-24)
      S d = p.yd - p.u;
      return 0.5 * d * d + 0.5 * p.y * p.y;
    }
    case KAPPA_POISSON2: {
      // k1*(sol-y)*(sol-y) + 0.5*(k2-1)^2   (poissonmixed2.jl:39-42)
      S d = p.yd - p.y; S b = p.k[1] - 1.0;
      return p.k[0] * d * d + 0.5 * b * b;
    }
    case KAPPA_POISSON: {
      // 0.5*(sol-y)^2 + 0.5*(k1-1)^2 + 0.5*(k2-1)^2   (poissonmixed.jl:37-43)
      S d = p.yd - p.y; S a = p.k[0] - 1.0; S b = p.k[1] - 1.0;
      return 0.5 * d * d + 0.5 * a * a + 0.5 * b * b;
    }
  }
  return cst(0.0, p.y);
}

// weak residual density for ONE test function (value v, physical gradient gv)
template <class S>
static S res_density(const oracle_problem& P, const Pt<S>& p, double v, const double* gv, bool has_u) {
  S gg = cst(0.0, p.y);
  for (int d = 0; d < P.dim; ++d) gg = gg + gv[d] * p.gy[d];   // grad(v) . grad(y)
  S one_dot = cst(0.0, p.y);
  for (int d = 0; d < P.dim; ++d) one_dot = one_dot + p.gy[d]; // grad(y) . one(grad(y))
  switch (P.integrand) {
    case PENALIZED_POISSON:
    case BASIC_UNCONSTRAINED:  // unconstrained problems have no residual
      return cst(0.0, p.y);
    case MEMBRANE:   // grad(v).grad(y) - v*u - v*h
      return gg - v * p.u - v * p.h;
    case PB:         // grad(v).grad(y) + sinh(y)*v - u*v - v*h
      return gg + xsinh(p.y) * v - p.u * v - v * p.h;
    case BURGER_LIN: // -nu*(grad(v).grad(y)) + dt(y,v) - v*u - v*h ; dt(y,v) = (grad(y).1)*v
      return (-P.params[1]) * gg + one_dot * v - v * p.u - v * p.h;
    case BURGER_NL:  // nu*(grad(v).grad(y)) + k1 + c(y,v) - v*u - v*h ; k1(x) = -k[1] ; c = v*(y*(grad(y).1))
      return P.params[1] * gg - p.k[0] + v * (one_dot * p.y) - v * p.u - v * p.h;
    case POISSON_PARAM:    // k1*grad(v).grad(y) - v*f   (poissonparam.jl:35-38)
      return p.k[0] * gg - v * p.h;
    case KAPPA_POISSON2:
    case KAPPA_POISSON: {  // k1*grad(v).grad(y) - v*f*k2 [- v*u]
      S r = p.k[0] * gg - (v * p.h) * p.k[1];
      if (has_u) r = r - v * p.u;
      return r;
    }
  }
  return cst(0.0, p.y);
}

// ------------------------------------------------------------------ cell evaluation
struct CellGeom {
  bool mapped;
  double x0[3];   // lower corner          (plain Cartesian cell)
  double h[3];
  double detJ;
  // mapped cell: per quadrature point
  double xq[27][3], detJq[27], JinvT[27][3][3];
};
static void invert_transpose(int dim, const double J[3][3], double T[3][3], double* det) {
  if (dim == 1) { *det = J[0][0]; T[0][0] = 1.0 / J[0][0]; return; }
  if (dim == 2) {
    const double d = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    *det = d;
    // J^-1 = 1/d [[J11, -J01], [-J10, J00]];  T = (J^-1)^T
    T[0][0] = J[1][1] / d; T[0][1] = -J[1][0] / d; T[1][0] = -J[0][1] / d; T[1][1] = J[0][0] / d;
    return;
  }
  double c[3][3];  // cofactors
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      c[i][j] = J[i1][j1] * J[i2][j2] - J[i1][j2] * J[i2][j1];
    }
  const double d = J[0][0] * c[0][0] + J[0][1] * c[0][1] + J[0][2] * c[0][2];
  *det = d;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i][j] = c[i][j] / d;  // (J^-1)^T = cofactor matrix / det
}
static CellGeom cell_geom(const oracle_problem& P, const RefTables& T, int64_t c_local);
// quadrature weight * |det J| at point q
static inline double geom_w(const CellGeom& G, const double* w, int q) { return w[q] * (G.mapped ? G.detJq[q] : G.detJ); }
// physical gradient of a shape function from its reference gradient
static inline void geom_grad(const oracle_problem& P, const CellGeom& G, int q, const double* ref, double* phys) {
  for (int d = 0; d < P.dim; ++d) {
    if (!G.mapped) { phys[d] = ref[d] / G.h[d]; continue; }
    double a = 0.0;
    for (int e = 0; e < P.dim; ++e) a += G.JinvT[q][d][e] * ref[e];
    phys[d] = a;
  }
}
static inline void geom_point(const oracle_problem& P, const CellGeom& G, const double* xi, int q, double* xq) {
  for (int d = 0; d < 3; ++d) xq[d] = 0.0;
  for (int d = 0; d < P.dim; ++d) xq[d] = G.mapped ? G.xq[q][d] : G.x0[d] + G.h[d] * xi[q * P.dim + d];
}

static CellGeom cell_geom(const oracle_problem& P, const RefTables& T, int64_t c_local) {
  CellGeom G; int64_t c = P.cell0 + c_local; G.detJ = 1; G.mapped = P.vertex_coords != nullptr;
  int64_t idx[3] = {0, 0, 0};
  for (int d = 0; d < 3; ++d) { G.x0[d] = 0; G.h[d] = 1; }
  for (int d = 0; d < P.dim; ++d) {
    int64_t i = c % P.n[d]; c /= P.n[d];
    idx[d] = i;
    G.h[d] = (P.hi[d] - P.lo[d]) / P.n[d];
    G.x0[d] = P.lo[d] + G.h[d] * (double)i;
    G.detJ *= G.h[d];
  }
  if (!G.mapped) return G;
  // vertices of the cell (n-cube order, x fastest) in the x-fastest vertex numbering of the mesh
  const int nv = 1 << P.dim;
  int64_t stride[3] = {1, 0, 0};
  for (int d = 1; d < P.dim; ++d) stride[d] = stride[d - 1] * (P.n[d - 1] + 1);
  double xv[8][3];
  for (int v = 0; v < nv; ++v) {
    int64_t id = 0;
    for (int d = 0; d < P.dim; ++d) id += (idx[d] + ((v >> d) & 1)) * stride[d];
    for (int d = 0; d < P.dim; ++d) xv[v][d] = P.vertex_coords[id * P.dim + d];
  }
  for (int q = 0; q < T.nq; ++q) {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int d = 0; d < 3; ++d) G.xq[q][d] = 0.0;
    for (int v = 0; v < nv; ++v) {
      // Q1 shape function of vertex v and its reference gradient at xi_q
      double N = 1.0, dN[3] = {1.0, 1.0, 1.0};
      for (int d = 0; d < P.dim; ++d) {
        const double xi = T.xi[q * P.dim + d];
        const bool hi = (v >> d) & 1;
        const double val = hi ? xi : 1.0 - xi, der = hi ? 1.0 : -1.0;
        N *= val;
        for (int e = 0; e < P.dim; ++e) dN[e] *= (e == d) ? der : val;
      }
      for (int d = 0; d < P.dim; ++d) {
        G.xq[q][d] += N * xv[v][d];
        for (int e = 0; e < P.dim; ++e) J[d][e] += xv[v][d] * dN[e];
      }
    }
    double det;
    invert_transpose(P.dim, J, G.JinvT[q], &det);
    G.detJq[q] = std::fabs(det);
  }
  return G;
}

// gather the cell-local values [kappa ; y_e ; u_e] (PosNegReindex semantics)
template <int NP, int NY, int NU>
static void gather_cell(const oracle_problem& P, const double* x, int64_t c, double* z, bool unshifted = false) {
  for (int k = 0; k < NP; ++k) z[k] = x[k];
  const double* xy = x + (unshifted ? 0 : NP);
  const double* xu = xy + P.nfree_y;
  for (int a = 0; a < NY; ++a) { int g = P.cell_dofs_y[c * NY + a]; z[NP + a] = g > 0 ? xy[g - 1] : P.dir_y[-g - 1]; }
  for (int b = 0; b < NU; ++b) { int g = P.cell_dofs_u[c * NU + b]; z[NP + NY + b] = g > 0 ? xu[g - 1] : P.dir_u[-g - 1]; }
}

// build the pointwise data at quadrature point q from cell-local scalars zz
template <class S, int NP, int NY, int NU>
static Pt<S> point_data(const oracle_problem& P, const RefTables& T, const CellGeom& G, int q, const S* zz) {
  Pt<S> p;
  S zero = cst(0.0, zz[0]);
  p.y = zero; p.u = zero; for (int d = 0; d < 3; ++d) p.gy[d] = zero;
  p.k[0] = zero; p.k[1] = zero;
  for (int k = 0; k < NP; ++k) p.k[k] = zz[k];
  for (int a = 0; a < NY; ++a) {
    p.y = p.y + T.phiy[q * NY + a] * zz[NP + a];
    double gph[3]; geom_grad(P, G, q, &T.dphiy[(q * NY + a) * P.dim], gph);
    for (int d = 0; d < P.dim; ++d) p.gy[d] = p.gy[d] + gph[d] * zz[NP + a];
  }
  for (int b = 0; b < NU; ++b) p.u = p.u + T.phiu[q * NU + b] * zz[NP + NY + b];
  double xq[3]; geom_point(P, G, T.xi.data(), q, xq);
  p.yd = coef_target(P, xq);
  p.h = coef_source(P, xq);
  return p;
}

// cell objective  F_e(z) = sum_q w_q detJ f(...)
template <class S, int NP, int NY, int NU>
static S cell_obj(const oracle_problem& P, const RefTables& T, const CellGeom& G, const S* zz) {
  S acc = cst(0.0, zz[0]);
  for (int q = 0; q < T.nq; ++q) {
    Pt<S> p = point_data<S, NP, NY, NU>(P, T, G, q, zz);
    acc = acc + geom_w(G, T.w.data(), q) * f_density(P, p);
  }
  return acc;
}

// cell residual vector  c_e[i](z) = sum_q w_q detJ res(..., v_i), i over the NY test functions
template <class S, int NP, int NY, int NU>
static void cell_res(const oracle_problem& P, const RefTables& T, const CellGeom& G, const S* zz, S* out) {
  for (int i = 0; i < NY; ++i) out[i] = cst(0.0, zz[0]);
  for (int q = 0; q < T.nq; ++q) {
    Pt<S> p = point_data<S, NP, NY, NU>(P, T, G, q, zz);
    for (int i = 0; i < NY; ++i) {
      double gv[3];
      geom_grad(P, G, q, &T.dphiy[(q * NY + i) * P.dim], gv);
      out[i] = out[i] + geom_w(G, T.w.data(), q) * res_density(P, p, T.phiy[q * NY + i], gv, NU > 0);
    }
  }
}

// ------------------------------------------------------------------ typed drivers
template <int NP, int NY, int NU>
struct Driver {
  static const int N = NP + NY + NU;
  typedef Dual<double, N> D1;
  typedef Dual<D1, N> D2;

  static void seed1(const double* z, D1* zz) {
    for (int i = 0; i < N; ++i) { zz[i].v = z[i]; for (int j = 0; j < N; ++j) zz[i].d[j] = (i == j) ? 1.0 : 0.0; }
  }
  static void seed2(const double* z, D2* zz) {
    for (int i = 0; i < N; ++i) {
      zz[i].v.v = z[i];
      for (int j = 0; j < N; ++j) zz[i].v.d[j] = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < N; ++k) {
        zz[i].d[k].v = (i == k) ? 1.0 : 0.0;
        for (int j = 0; j < N; ++j) zz[i].d[k].d[j] = 0.0;
      }
    }
  }

  static double obj(const oracle_problem& P, const double* x) {
    if (is_nofe(P)) return nofe_obj<double>(P, x, NP);  // _obj_integral(::NoFETerm, k, x) = f(k)
    RefTables T = make_tables(P);
    CellSums S(1, P.ncells);  // sum(::DomainContribution): pairwise over the cells (see pairwise_sum)
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather_cell<NP, NY, NU>(P, x, c, z);
      CellGeom G = cell_geom(P, T, c);
      S.at(0, c) = cell_obj<double, NP, NY, NU>(P, T, G, z);
    }
    return S.sum(0);
  }

  static void grad(const oracle_problem& P, const double* x, double* g) {
    RefTables T = make_tables(P);
    int64_t nvar = NP + P.nfree_y + P.nfree_u;
    for (int64_t i = 0; i < nvar; ++i) g[i] = 0.0;
    if (is_nofe(P)) {  // g[p+1:] = 0, g[1:p] = ForwardDiff.gradient(f, k)   (additional_obj_terms.jl:69-95)
      typedef Dual<double, (NP > 0 ? NP : 1)> DK;
      DK k[NP > 0 ? NP : 1];
      for (int i = 0; i < NP; ++i) { k[i].v = x[i]; for (int j = 0; j < NP; ++j) k[i].d[j] = i == j; }
      DK F = nofe_obj<DK>(P, k, NP);
      for (int i = 0; i < NP; ++i) g[i] = F.d[i];
      return;
    }
    CellSums S(NP, P.ncells);  // g[1:p] = d(sum over cells)/dkappa: the same pairwise `sum`, on duals
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather_cell<NP, NY, NU>(P, x, c, z);
      D1 zz[N]; seed1(z, zz);
      CellGeom G = cell_geom(P, T, c);
      D1 F = cell_obj<D1, NP, NY, NU>(P, T, G, zz);
      for (int k = 0; k < NP; ++k) S.at(k, c) = F.d[k];
      for (int a = 0; a < NY; ++a) { int gid = P.cell_dofs_y[c * NY + a]; if (gid > 0) g[NP + gid - 1] += F.d[NP + a]; }
      for (int b = 0; b < NU; ++b) { int gid = P.cell_dofs_u[c * NU + b]; if (gid > 0) g[NP + P.nfree_y + gid - 1] += F.d[NP + NY + b]; }
    }
    for (int k = 0; k < NP; ++k) g[k] = S.sum(k);
  }

  static void cons(const oracle_problem& P, const double* x, double* cvec) {
    RefTables T = make_tables(P);
    for (int64_t i = 0; i < P.nfree_y; ++i) cvec[i] = 0.0;
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather_cell<NP, NY, NU>(P, x, c, z);
      CellGeom G = cell_geom(P, T, c);
      double r[NY]; cell_res<double, NP, NY, NU>(P, T, G, z, r);
      for (int i = 0; i < NY; ++i) { int gid = P.cell_dofs_y[c * NY + i]; if (gid > 0) cvec[gid - 1] += r[i]; }
    }
  }

  // ---- per-cell counts (used for the segment offsets) -----------------------
  static int count_jy(const oracle_problem& P, int64_t c) {
    int n = 0;
    for (int j = 0; j < NY; ++j) if (P.cell_dofs_y[c * NY + j] > 0)
      for (int i = 0; i < NY; ++i) if (P.cell_dofs_y[c * NY + i] > 0) ++n;
    return n;
  }
  static int count_ju(const oracle_problem& P, int64_t c) {
    int n = 0;
    for (int j = 0; j < NU; ++j) if (P.cell_dofs_u[c * NU + j] > 0)
      for (int i = 0; i < NY; ++i) if (P.cell_dofs_y[c * NY + i] > 0) ++n;
    return n;
  }
  // multi-field global ids of the cell: y ids as they are, u ids + nfree_y when positive
  static void cell_gids(const oracle_problem& P, int64_t c, int64_t* gid) {
    for (int a = 0; a < NY; ++a) gid[a] = P.cell_dofs_y[c * NY + a];
    for (int b = 0; b < NU; ++b) { int64_t g = P.cell_dofs_u[c * NU + b]; gid[NY + b] = g > 0 ? g + P.nfree_y : g; }
  }
  // Visit the kept Hessian entries of one cell in the reference order.
  // fn(local_row, local_col, gidrow, gidcol); local indices into [y_e ; u_e].
  template <class F>
  static void visit_hess_cell(const oracle_problem& P, int64_t c, F fn) {
    int64_t gid[NY + NU + 1]; cell_gids(P, c, gid);
    const int off[2] = {0, NY}, len[2] = {NY, NU};
    const int nb = NU > 0 ? 2 : 1;
    // eachblockid: column-major over field pairs -> (1,1),(2,1),(1,2),(2,2)
    for (int bj = 0; bj < nb; ++bj)
      for (int bi = 0; bi < nb; ++bi)
        for (int j = 0; j < len[bj]; ++j) {
          int64_t gc = gid[off[bj] + j];
          if (gc > 0)
            for (int i = 0; i < len[bi]; ++i) {
              int64_t gr = gid[off[bi] + i];
              if (gr > 0 && gc <= gr) fn(off[bi] + i, off[bj] + j, gr, gc);
            }
        }
  }
  static int count_h(const oracle_problem& P, int64_t c) {
    int n = 0; visit_hess_cell(P, c, [&](int, int, int64_t, int64_t) { ++n; }); return n;
  }

  static void sizes(const oracle_problem& P, int64_t* out) {
    // out: nnzj_k, nnzj_y, nnzj_u, nnzh_k_obj, nnzh_fe, nnzh_k_con
    int64_t jy = 0, ju = 0, hf = 0;
    for (int64_t c = 0; c < P.ncells; ++c) { jy += count_jy(P, c); ju += count_ju(P, c); hf += count_h(P, c); }
    int64_t n = NP + P.nfree_y + P.nfree_u, p = NP;
    out[0] = P.nfree_y * p; out[1] = jy; out[2] = ju;
    bool inde = (P.integrand == KAPPA_POISSON) || is_nofe(P);  // MixedEnergyFETerm(..., inde = true), poissonmixed.jl:44
    out[3] = p == 0 ? 0 : (inde ? p * (p + 1) / 2 : p * (p + 1) / 2 + (n - p) * p);
    out[4] = hf;
    out[5] = p == 0 ? 0 : p * (p + 1) / 2 + (n - p) * p;
    out[6] = is_nofe(P) ? 0 : hf;  // objective FE block
  }

  static void jac_structure(const oracle_problem& P, int64_t* rows, int64_t* cols) {
    int64_t n = 0;
    // jac_k_structure: ((i,j) for i = 1:ncon, j = 1:p), i fastest
    for (int j = 1; j <= NP; ++j) for (int64_t i = 1; i <= P.nfree_y; ++i) { rows[n] = i; cols[n] = j; ++n; }
    for (int64_t c = 0; c < P.ncells; ++c)
      for (int j = 0; j < NY; ++j) { int gc = P.cell_dofs_y[c * NY + j]; if (gc > 0)
        for (int i = 0; i < NY; ++i) { int gr = P.cell_dofs_y[c * NY + i]; if (gr > 0) { rows[n] = gr; cols[n] = gc + NP; ++n; } } }
    for (int64_t c = 0; c < P.ncells; ++c)
      for (int j = 0; j < NU; ++j) { int gc = P.cell_dofs_u[c * NU + j]; if (gc > 0)
        for (int i = 0; i < NY; ++i) { int gr = P.cell_dofs_y[c * NY + i]; if (gr > 0) { rows[n] = gr; cols[n] = gc + P.nfree_y + NP; ++n; } } }
  }

  static void jac_coord(const oracle_problem& P, const double* x, double* vals) {
    RefTables T = make_tables(P);
    int64_t sz[7]; sizes(P, sz);
    std::vector<int64_t> by(P.ncells + 1, 0), bu(P.ncells + 1, 0);
    for (int64_t c = 0; c < P.ncells; ++c) { by[c + 1] = by[c] + count_jy(P, c); bu[c + 1] = bu[c] + count_ju(P, c); }
    double* vk = vals; double* vy = vals + sz[0]; double* vu = vy + sz[1];
    for (int64_t i = 0; i < sz[0]; ++i) vk[i] = 0.0;
#pragma omp parallel for schedule(static) num_threads(P.nthreads > 0 ? P.nthreads : omp_get_max_threads())
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather_cell<NP, NY, NU>(P, x, c, z);
      D1 zz[N]; seed1(z, zz);
      CellGeom G = cell_geom(P, T, c);
      D1 r[NY]; cell_res<D1, NP, NY, NU>(P, T, G, zz, r);
      int64_t n = by[c];
      for (int j = 0; j < NY; ++j) if (P.cell_dofs_y[c * NY + j] > 0)
        for (int i = 0; i < NY; ++i) if (P.cell_dofs_y[c * NY + i] > 0) vy[n++] = r[i].d[NP + j];
      n = bu[c];
      for (int j = 0; j < NU; ++j) if (P.cell_dofs_u[c * NU + j] > 0)
        for (int i = 0; i < NY; ++i) if (P.cell_dofs_y[c * NY + i] > 0) vu[n++] = r[i].d[NP + NY + j];
      if (NP > 0) {
#pragma omp critical
        for (int k = 0; k < NP; ++k)
          for (int i = 0; i < NY; ++i) { int gr = P.cell_dofs_y[c * NY + i]; if (gr > 0) vk[(int64_t)k * P.nfree_y + gr - 1] += r[i].d[k]; }
      }
    }
  }

  static void hess_structure(const oracle_problem& P, int64_t* rows, int64_t* cols) {
    int64_t sz[7]; sizes(P, sz);
    int64_t n = 0, nvar = NP + P.nfree_y + P.nfree_u;
    bool inde = (P.integrand == KAPPA_POISSON) || is_nofe(P);
    int64_t prows = inde ? NP : nvar;
    // ((i, j) for i = 1:prows, j = 1:p if j <= i), i fastest
    for (int j = 1; j <= NP; ++j) for (int64_t i = j; i <= prows; ++i) { rows[n] = i; cols[n] = j; ++n; }
    if (!is_nofe(P))
    for (int64_t c = 0; c < P.ncells; ++c)
      visit_hess_cell(P, c, [&](int, int, int64_t gr, int64_t gc) { rows[n] = gr + NP; cols[n] = gc + NP; ++n; });
    for (int j = 1; j <= NP; ++j) for (int64_t i = j; i <= nvar; ++i) { rows[n] = i; cols[n] = j; ++n; }
    for (int64_t c = 0; c < P.ncells; ++c)
      visit_hess_cell(P, c, [&](int, int, int64_t gr, int64_t gc) { rows[n] = gr + NP; cols[n] = gc + NP; ++n; });
  }

  // position of (global row i, kappa col j), 1-based, in the trapezoid "i fastest, j <= i" with prows rows
  static int64_t trap_pos(int64_t i, int j, int64_t prows) {
    int64_t pos = 0;
    for (int jj = 1; jj < j; ++jj) pos += prows - jj + 1;
    return pos + (i - j);
  }

  static void hess_coord(const oracle_problem& P, const double* x, const double* lambda, double obj_weight, double* vals) {
    RefTables T = make_tables(P);
    int64_t sz[7]; sizes(P, sz);
    int64_t nvar = NP + P.nfree_y + P.nfree_u;
    const bool nofe = is_nofe(P);
    bool inde = (P.integrand == KAPPA_POISSON) || nofe;
    int64_t prows = inde ? NP : nvar;
    std::vector<int64_t> bh(P.ncells + 1, 0);
    for (int64_t c = 0; c < P.ncells; ++c) bh[c + 1] = bh[c] + count_h(P, c);
    double* vko = vals; double* vfo = vko + sz[3]; double* vkc = vfo + sz[6]; double* vfc = vkc + sz[5];
    int64_t nnzh = sz[3] + sz[6] + sz[5] + sz[4];
    for (int64_t i = 0; i < nnzh; ++i) vals[i] = 0.0;
    if (nofe && NP > 0) {  // vals[1:nnz_hess_k] = LowerTriangular(ForwardDiff.hessian(f, k))   (additional_obj_terms.jl:97-106)
      typedef Dual<double, (NP > 0 ? NP : 1)> DK1; typedef Dual<DK1, (NP > 0 ? NP : 1)> DK2;
      DK2 k[NP > 0 ? NP : 1];
      for (int i = 0; i < NP; ++i) {
        k[i].v.v = x[i];
        for (int j = 0; j < NP; ++j) { k[i].v.d[j] = i == j; k[i].d[j].v = i == j; for (int l = 0; l < NP; ++l) k[i].d[j].d[l] = 0.0; }
      }
      DK2 F = nofe_obj<DK2>(P, k, NP);
      for (int j = 0; j < NP; ++j) for (int i = j; i < NP; ++i) vko[trap_pos(i + 1, j + 1, prows)] = F.d[i].d[j];
    }
    // kappa-kappa corner entries are sums over ALL cells: per-cell values, summed pairwise afterwards (deterministic,
    // independent of the thread count); the cross rows (a dof is touched by a handful of cells) are added in place
    constexpr int NKK = NP * (NP + 1) / 2;
    CellSums Sko(NP > 0 ? NKK : 0, NP > 0 ? P.ncells : 0), Skc(NP > 0 ? NKK : 0, NP > 0 ? P.ncells : 0);
#pragma omp parallel for schedule(static) num_threads(P.nthreads > 0 ? P.nthreads : omp_get_max_threads())
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather_cell<NP, NY, NU>(P, x, c, z);
      D2 zz[N]; seed2(z, zz);
      CellGeom G = cell_geom(P, T, c);
      int64_t gid[NY + NU + 1]; cell_gids(P, c, gid);
      // objective part
      int64_t n = bh[c];
      D2 F = cst(0.0, zz[0]);
      if (!nofe) {
        F = cell_obj<D2, NP, NY, NU>(P, T, G, zz);
        visit_hess_cell(P, c, [&](int li, int lj, int64_t, int64_t) { vfo[n++] = F.d[NP + li].d[NP + lj]; });
      }
      if (NP > 0 && !nofe) {
        for (int j = 0, e = 0; j < NP; ++j)
          for (int i = j; i < NP; ++i, ++e) Sko.at(e, c) = F.d[i].d[j];
#pragma omp critical
        for (int j = 0; j < NP; ++j) {
          if (!inde) for (int a = 0; a < NY + NU; ++a) { int64_t g = gid[a]; if (g > 0) vko[trap_pos(g + NP, j + 1, prows)] += F.d[NP + a].d[j]; }
        }
      }
      if (lambda) {
        // l_e(z) = sum_i lambda_e[i] c_e[i](z), lambda through the TEST space (Dirichlet entries = 0)
        D2 zc[N];
        if (NP > 0 && P.compat_shifted_x) { double z2[N]; gather_cell<NP, NY, NU>(P, x, c, z2, true); seed2(z2, zc); }
        D2 r[NY]; cell_res<D2, NP, NY, NU>(P, T, G, (NP > 0 && P.compat_shifted_x) ? zc : zz, r);
        D2 L = cst(0.0, zz[0]);
        for (int i = 0; i < NY; ++i) { int g = P.cell_dofs_y[c * NY + i]; double li = g > 0 ? lambda[g - 1] : 0.0; L = L + li * r[i]; }
        n = bh[c];
        visit_hess_cell(P, c, [&](int li, int lj, int64_t, int64_t) { vfc[n++] = L.d[NP + li].d[NP + lj]; });
        if (NP > 0) {
          for (int j = 0, e = 0; j < NP; ++j)
            for (int i = j; i < NP; ++i, ++e) Skc.at(e, c) = L.d[i].d[j];
#pragma omp critical
          for (int j = 0; j < NP; ++j) {
            for (int a = 0; a < NY + NU; ++a) { int64_t g = gid[a]; if (g > 0) vkc[trap_pos(g + NP, j + 1, nvar)] += L.d[NP + a].d[j]; }
          }
        }
      }
    }
    if (NP > 0) {
      for (int j = 0, e = 0; j < NP; ++j)
        for (int i = j; i < NP; ++i, ++e) {
          if (!nofe) vko[trap_pos(i + 1, j + 1, prows)] = Sko.sum(e);
          if (lambda) vkc[trap_pos(i + 1, j + 1, nvar)] = Skc.sum(e);
        }
    }
    // vals .*= obj_weight on the objective part only (GridapPDENLPModel.jl:299 runs before the constraint part)
    for (int64_t i = 0; i < sz[3] + sz[6]; ++i) vals[i] *= obj_weight;
  }
};


// ================================================================== multi-field
// Ypde / Ycon are MultiFieldFESpaces of scalar Lagrangian fields (test/problems/controlsir.jl,
// cellincrease.jl, dynamicsir.jl, torebrachistochrone.jl).  Cell-local vector z = [y_1 .. y_nfy ; u_1 .. u_nfu]
// (field-major), consecutive multi-field numbering of the free dofs, one test field per state field.
// Cell matrices are block arrays: blocks are visited in eachblockid order (column-major over field pairs),
// inside a block column outer / row inner (src/fill_matrix_hessian_functions.jl:69-94).
template <class S>
struct PtMF {
  S y[4], gy[4][3], u[4];
  double yd, h;
};

template <class S>
static S f_density_mf(const oracle_problem& P, const PtMF<S>& p) {
  switch (P.integrand) {
    case CONTROL_SIR:   // 0.5 * I * I                           (controlsir.jl:30-33)
      return 0.5 * p.y[0] * p.y[0];
    case CELL_INCREASE: // kp * pf                               (cellincrease.jl:15-18)
      return P.params[0] * p.y[1];
    case DYNAMIC_SIR: { // 0.5*((bf*w0)-w1)^2 + 0.5*((cf*w0)-w2)^2   (dynamicsir.jl:54-58)
      S a = p.yd * p.u[0] - P.params[0];
      S b = p.yd * p.u[1] - P.params[1];
      return 0.5 * a * a + 0.5 * b * b;
    }
    case TORE_BRACHISTOCHRONE: {  // a*a*phi'^2 + (c + a*cos(phi))^2 * theta'^2   (torebrachistochrone.jl:25-28)
      const double a = P.params[0], c = P.params[1];
      S w = c + a * xcos(p.y[0]);
      return (a * a) * p.gy[0][0] * p.gy[0][0] + w * w * p.gy[1][0] * p.gy[1][0];
    }
  }
  return cst(0.0, p.y[0]);
}
// residual density against ONE test function (value v) of test field ft; dt(y, v) = (grad(y).1) * v
template <class S>
static S res_density_mf(const oracle_problem& P, const PtMF<S>& p, int ft, double v) {
  const S& I = p.y[0]; const S& Sv = p.y[1];
  const S& dI = p.gy[0][0]; const S& dS = p.gy[1][0];
  switch (P.integrand) {
    case CONTROL_SIR: {  // dt(I,p) + dt(S,q) - p*(a*S*I - b*I) - q*(mu - b*I - a*S*I)   (controlsir.jl:20-24)
      const double a = P.params[0], b = P.params[1], mu = P.params[2];
      if (ft == 0) return dI * v - v * (a * Sv * I - b * I);
      return dS * v - v * (mu - b * I - a * Sv * I);
    }
    case CELL_INCREASE: {  // cellincrease.jl:33-40 with (cf, pf) = y, u the control
      const double kp = P.params[0], kr = P.params[1];
      const S& cf = I; const S& pf = Sv;
      if (ft == 0) return dI * v - v * (kp * pf * (1.0 - cf) - kr * cf * (1.0 - cf - pf));
      return dS * v + v * (p.u[0] * kr * cf * (1.0 - cf - pf) - kp * pf * pf);
    }
    case SIS: {  // dt(I,p) + dt(S,q) - p*(a*S*I - b*I) - q*(b*I - a*S*I)   (sis.jl:21-25)
      const double a = P.params[0], b = P.params[1];
      if (ft == 0) return dI * v - v * (a * Sv * I - b * I);
      return dS * v - v * (b * I - a * Sv * I);
    }
    case DYNAMIC_SIR: {  // dt(I,p) + dt(S,q) - p*(bf*S*I - cf*I) + q*bf*S*I   (dynamicsir.jl:36-41)
      if (ft == 0) return dI * v - v * (p.u[0] * Sv * I - p.u[1] * I);
      return dS * v + v * (p.u[0] * Sv * I);
    }
  }
  return cst(0.0, p.y[0]);
}

template <int NFY, int NFU, int NLY, int NLU>
struct MFDriver {
  static const int NYT = NFY * NLY, NUT = NFU * NLU, N = NYT + NUT, NF = NFY + NFU;
  typedef Dual<double, N> D1;
  typedef Dual<D1, N> D2;

  static int64_t off_free(const oracle_problem& P, int F) { int64_t o = 0; for (int f = 0; f < F; ++f) o += P.field_nfree[f]; return o; }
  static int64_t off_dir(const oracle_problem& P, int F) {   // within dir_y (state) or dir_u (control)
    int64_t o = 0; const int f0 = F < NFY ? 0 : NFY; for (int f = f0; f < F; ++f) o += P.field_ndir[f]; return o; }
  static int nloc(int F) { return F < NFY ? NLY : NLU; }
  static int zoff(int F) { return F < NFY ? F * NLY : NYT + (F - NFY) * NLU; }
  static int raw_id(const oracle_problem& P, int64_t c, int F, int a) {
    return F < NFY ? P.cell_dofs_y[c * NYT + F * NLY + a] : P.cell_dofs_u[c * NUT + (F - NFY) * NLU + a];
  }
  static void gather(const oracle_problem& P, const double* x, int64_t c, double* z) {
    for (int F = 0; F < NF; ++F)
      for (int a = 0; a < nloc(F); ++a) {
        const int g = raw_id(P, c, F, a);
        const double* dir = F < NFY ? P.dir_y : P.dir_u;
        z[zoff(F) + a] = g > 0 ? x[off_free(P, F) + g - 1] : dir[off_dir(P, F) - g - 1];
      }
  }
  // multi-field global ids (offset positive ids), indexed like z
  static void gids(const oracle_problem& P, int64_t c, int64_t* gid) {
    for (int F = 0; F < NF; ++F)
      for (int a = 0; a < nloc(F); ++a) { const int g = raw_id(P, c, F, a); gid[zoff(F) + a] = g > 0 ? g + off_free(P, F) : g; }
  }
  template <class S>
  static PtMF<S> point(const oracle_problem& P, const RefTables& T, const CellGeom& G, int q, const S* zz) {
    PtMF<S> p; S zero = cst(0.0, zz[0]);
    for (int f = 0; f < 4; ++f) { p.y[f] = zero; p.u[f] = zero; for (int d = 0; d < 3; ++d) p.gy[f][d] = zero; }
    for (int f = 0; f < NFY; ++f)
      for (int a = 0; a < NLY; ++a) {
        p.y[f] = p.y[f] + T.phiy[q * NLY + a] * zz[f * NLY + a];
        { double gph[3]; geom_grad(P, G, q, &T.dphiy[(q * NLY + a) * P.dim], gph);
          for (int d = 0; d < P.dim; ++d) p.gy[f][d] = p.gy[f][d] + gph[d] * zz[f * NLY + a]; }
      }
    for (int g = 0; g < NFU; ++g)
      for (int b = 0; b < NLU; ++b) p.u[g] = p.u[g] + T.phiu[q * NLU + b] * zz[NYT + g * NLU + b];
    double xq[3]; geom_point(P, G, T.xi.data(), q, xq);
    p.yd = coef_target(P, xq); p.h = coef_source(P, xq);
    return p;
  }
  template <class S> static S cell_obj(const oracle_problem& P, const RefTables& T, const CellGeom& G, const S* zz) {
    S acc = cst(0.0, zz[0]);
    for (int q = 0; q < T.nq; ++q) acc = acc + geom_w(G, T.w.data(), q) * f_density_mf(P, point<S>(P, T, G, q, zz));
    return acc;
  }
  // residual vector indexed like the state part of z: (test field ft, local i)
  template <class S> static void cell_res(const oracle_problem& P, const RefTables& T, const CellGeom& G, const S* zz, S* out) {
    for (int i = 0; i < NYT; ++i) out[i] = cst(0.0, zz[0]);
    for (int q = 0; q < T.nq; ++q) {
      PtMF<S> p = point<S>(P, T, G, q, zz);
      for (int ft = 0; ft < NFY; ++ft)
        for (int i = 0; i < NLY; ++i) out[ft * NLY + i] = out[ft * NLY + i] + geom_w(G, T.w.data(), q) * res_density_mf(P, p, ft, T.phiy[q * NLY + i]);
    }
  }
  static void seed1(const double* z, D1* zz) { for (int i = 0; i < N; ++i) { zz[i].v = z[i]; for (int j = 0; j < N; ++j) zz[i].d[j] = i == j; } }
  static void seed2(const double* z, D2* zz) {
    for (int i = 0; i < N; ++i) {
      zz[i].v.v = z[i];
      for (int j = 0; j < N; ++j) zz[i].v.d[j] = i == j;
      for (int k = 0; k < N; ++k) { zz[i].d[k].v = i == k; for (int j = 0; j < N; ++j) zz[i].d[k].d[j] = 0.0; }
    }
  }
  // visit the entries of a block matrix (rows fields [r0,r1), cols fields [c0,c1)) of one cell in the reference order
  template <class Fn> static void visit(const oracle_problem& P, int64_t c, int r0, int r1, int c0, int c1, bool lower, Fn fn) {
    int64_t gid[N + 1]; gids(P, c, gid);
    for (int bj = c0; bj < c1; ++bj)
      for (int bi = r0; bi < r1; ++bi)
        for (int j = 0; j < nloc(bj); ++j) {
          const int64_t gc = gid[zoff(bj) + j];
          if (gc > 0)
            for (int i = 0; i < nloc(bi); ++i) {
              const int64_t gr = gid[zoff(bi) + i];
              if (gr > 0 && (!lower || gc <= gr)) fn(zoff(bi) + i, zoff(bj) + j, gr, gc);
            }
        }
  }
  static int cnt(const oracle_problem& P, int64_t c, int r0, int r1, int c0, int c1, bool lower) {
    int n = 0; visit(P, c, r0, r1, c0, c1, lower, [&](int, int, int64_t, int64_t) { ++n; }); return n; }

  static double obj(const oracle_problem& P, const double* x) {
    RefTables T = make_tables(P); CellSums S(1, P.ncells);  // pairwise over the cells (see pairwise_sum)
    for (int64_t c = 0; c < P.ncells; ++c) { double z[N]; gather(P, x, c, z); S.at(0, c) = cell_obj<double>(P, T, cell_geom(P, T, c), z); }
    return S.sum(0);
  }
  static void grad(const oracle_problem& P, const double* x, double* g) {
    RefTables T = make_tables(P);
    const int64_t nvar = P.nfree_y + P.nfree_u;
    for (int64_t i = 0; i < nvar; ++i) g[i] = 0.0;
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather(P, x, c, z); D1 zz[N]; seed1(z, zz);
      D1 F = cell_obj<D1>(P, T, cell_geom(P, T, c), zz);
      int64_t gid[N + 1]; gids(P, c, gid);
      for (int a = 0; a < N; ++a) if (gid[a] > 0) g[gid[a] - 1] += F.d[a];
    }
  }
  static void cons(const oracle_problem& P, const double* x, double* cv) {
    RefTables T = make_tables(P);
    for (int64_t i = 0; i < P.nfree_y; ++i) cv[i] = 0.0;
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather(P, x, c, z); double r[NYT > 0 ? NYT : 1];
      cell_res<double>(P, T, cell_geom(P, T, c), z, r);
      int64_t gid[N + 1]; gids(P, c, gid);
      for (int i = 0; i < NYT; ++i) if (gid[i] > 0) cv[gid[i] - 1] += r[i];
    }
  }
  static void sizes(const oracle_problem& P, int64_t* out) {
    int64_t jy = 0, ju = 0, hf = 0;
    for (int64_t c = 0; c < P.ncells; ++c) {
      jy += cnt(P, c, 0, NFY, 0, NFY, false); ju += cnt(P, c, 0, NFY, NFY, NF, false); hf += cnt(P, c, 0, NF, 0, NF, true);
    }
    out[0] = 0; out[1] = jy; out[2] = ju; out[3] = 0; out[4] = hf; out[5] = 0; out[6] = is_nofe(P) ? 0 : hf;
  }
  static void jac_structure(const oracle_problem& P, int64_t* rows, int64_t* cols) {
    int64_t n = 0;
    for (int64_t c = 0; c < P.ncells; ++c) visit(P, c, 0, NFY, 0, NFY, false, [&](int, int, int64_t gr, int64_t gc) { rows[n] = gr; cols[n] = gc; ++n; });
    for (int64_t c = 0; c < P.ncells; ++c) visit(P, c, 0, NFY, NFY, NF, false, [&](int, int, int64_t gr, int64_t gc) { rows[n] = gr; cols[n] = gc; ++n; });
  }
  static void jac_coord(const oracle_problem& P, const double* x, double* vals) {
    RefTables T = make_tables(P);
    int64_t sz[7]; sizes(P, sz);
    double* vy = vals; double* vu = vals + sz[1];
    int64_t ny = 0, nu = 0;
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather(P, x, c, z); D1 zz[N]; seed1(z, zz);
      D1 r[NYT > 0 ? NYT : 1]; cell_res<D1>(P, T, cell_geom(P, T, c), zz, r);
      visit(P, c, 0, NFY, 0, NFY, false, [&](int li, int lj, int64_t, int64_t) { vy[ny++] = r[li].d[lj]; });
      visit(P, c, 0, NFY, NFY, NF, false, [&](int li, int lj, int64_t, int64_t) { vu[nu++] = r[li].d[lj]; });
    }
  }
  static void hess_structure(const oracle_problem& P, int64_t* rows, int64_t* cols) {
    int64_t n = 0;
    for (int rep = is_nofe(P) ? 1 : 0; rep < 2; ++rep)
      for (int64_t c = 0; c < P.ncells; ++c) visit(P, c, 0, NF, 0, NF, true, [&](int, int, int64_t gr, int64_t gc) { rows[n] = gr; cols[n] = gc; ++n; });
  }
  static void hess_coord(const oracle_problem& P, const double* x, const double* lambda, double obj_weight, double* vals) {
    RefTables T = make_tables(P);
    int64_t sz[7]; sizes(P, sz);
    const bool nofe = is_nofe(P);
    double* vo = vals; double* vc = vals + sz[6];
    for (int64_t i = 0; i < sz[6] + sz[4]; ++i) vals[i] = 0.0;
    int64_t no = 0, nc = 0;
    for (int64_t c = 0; c < P.ncells; ++c) {
      double z[N]; gather(P, x, c, z); D2 zz[N]; seed2(z, zz);
      CellGeom G = cell_geom(P, T, c);
      if (!nofe) {
        D2 F = cell_obj<D2>(P, T, G, zz);
        visit(P, c, 0, NF, 0, NF, true, [&](int li, int lj, int64_t, int64_t) { vo[no++] = F.d[li].d[lj]; });
      }
      if (lambda) {
        D2 r[NYT > 0 ? NYT : 1]; cell_res<D2>(P, T, G, zz, r);
        int64_t gid[N + 1]; gids(P, c, gid);
        D2 L = cst(0.0, zz[0]);
        for (int i = 0; i < NYT; ++i) { const double li = gid[i] > 0 ? lambda[gid[i] - 1] : 0.0; L = L + li * r[i]; }
        visit(P, c, 0, NF, 0, NF, true, [&](int li, int lj, int64_t, int64_t) { vc[nc++] = L.d[li].d[lj]; });
      }
    }
    for (int64_t i = 0; i < sz[6]; ++i) vals[i] *= obj_weight;
  }
};

#define MF_DISPATCH(CALL)                                                                      \
  do {                                                                                         \
    const int nfy = P->nfy, nfu = P->nfu;                                                      \
    const int nly = nloc_of(P->order_y, P->dim), nlu = nloc_of(P->order_u, P->dim);            \
    if (nfy == 2 && nfu == 0 && nly == 2) { typedef MFDriver<2, 0, 2, 1> D; CALL; }            \
    else if (nfy == 2 && nfu == 1 && nly == 2 && nlu == 2) { typedef MFDriver<2, 1, 2, 2> D; CALL; } \
    else if (nfy == 2 && nfu == 2 && nly == 2 && nlu == 2) { typedef MFDriver<2, 2, 2, 2> D; CALL; } \
    else return -2;                                                                            \
    return 0;                                                                                  \
  } while (0)
static bool is_mf(const oracle_problem* P) { return (P->integrand >= CONTROL_SIR && P->integrand <= TORE_BRACHISTOCHRONE) || P->integrand == SIS; }

// ------------------------------------------------------------------ dispatch
#define DISPATCH(CALL)                                                             \
  do {                                                                             \
    int ny = nloc_of(P->order_y, P->dim), nu = nloc_of(P->order_u, P->dim);        \
    int np = P->nparam;                                                            \
    if (np == 0 && ny == 9 && nu == 4) { typedef Driver<0, 9, 4> D; CALL; }        \
    else if (np == 0 && ny == 4 && nu == 4) { typedef Driver<0, 4, 4> D; CALL; }   \
    else if (np == 0 && ny == 4 && nu == 0) { typedef Driver<0, 4, 0> D; CALL; }   \
    else if (np == 0 && ny == 2 && nu == 2) { typedef Driver<0, 2, 2> D; CALL; }   \
    else if (np == 1 && ny == 2 && nu == 2) { typedef Driver<1, 2, 2> D; CALL; }   \
    else if (np == 2 && ny == 4 && nu == 0) { typedef Driver<2, 4, 0> D; CALL; }   \
    else if (np == 2 && ny == 4 && nu == 4) { typedef Driver<2, 4, 4> D; CALL; }   \
    else if (np == 2 && ny == 8 && nu == 8) { typedef Driver<2, 8, 8> D; CALL; }   \
    else if (np == 2 && ny == 8 && nu == 0) { typedef Driver<2, 8, 0> D; CALL; }   \
    else if (np == 0 && ny == 8 && nu == 8) { typedef Driver<0, 8, 8> D; CALL; }   \
    else if (np == 1 && ny == 4 && nu == 0) { typedef Driver<1, 4, 0> D; CALL; }   \
    else return -2;                                                                \
  } while (0)

extern "C" {

// out7: nnzj_k, nnzj_y, nnzj_u, nnzh_k_obj, nnzh_fe (constraint FE block), nnzh_k_con, nnzh_fe_obj (objective FE block)
int oracle_sizes(const oracle_problem* P, int64_t* out7) { if (is_mf(P)) MF_DISPATCH(D::sizes(*P, out7)); DISPATCH(D::sizes(*P, out7)); return 0; }
int oracle_obj(const oracle_problem* P, const double* x, double* f) { if (is_mf(P)) MF_DISPATCH(*f = D::obj(*P, x)); DISPATCH(*f = D::obj(*P, x)); return 0; }
int oracle_grad(const oracle_problem* P, const double* x, double* g) { if (is_mf(P)) MF_DISPATCH(D::grad(*P, x, g)); DISPATCH(D::grad(*P, x, g)); return 0; }
int oracle_cons(const oracle_problem* P, const double* x, double* c) { if (is_mf(P)) MF_DISPATCH(D::cons(*P, x, c)); DISPATCH(D::cons(*P, x, c)); return 0; }
int oracle_jac_structure(const oracle_problem* P, int64_t* rows, int64_t* cols) { if (is_mf(P)) MF_DISPATCH(D::jac_structure(*P, rows, cols)); DISPATCH(D::jac_structure(*P, rows, cols)); return 0; }
int oracle_jac_coord(const oracle_problem* P, const double* x, double* vals) { if (is_mf(P)) MF_DISPATCH(D::jac_coord(*P, x, vals)); DISPATCH(D::jac_coord(*P, x, vals)); return 0; }
int oracle_hess_structure(const oracle_problem* P, int64_t* rows, int64_t* cols) { if (is_mf(P)) MF_DISPATCH(D::hess_structure(*P, rows, cols)); DISPATCH(D::hess_structure(*P, rows, cols)); return 0; }
int oracle_hess_coord(const oracle_problem* P, const double* x, const double* lambda, double obj_weight, double* vals) {
  if (is_mf(P)) MF_DISPATCH(D::hess_coord(*P, x, lambda, obj_weight, vals));
  DISPATCH(D::hess_coord(*P, x, lambda, obj_weight, vals)); return 0;
}

// NLPModels coo_prod!(rows, cols, vals, v, Av): fill!(Av,0); Av[rows[k]] += vals[k]*v[cols[k]]
void oracle_coo_prod(int64_t nnz, const int64_t* rows, const int64_t* cols, const double* vals,
                     const double* v, int64_t nout, double* Av) {
  for (int64_t i = 0; i < nout; ++i) Av[i] = 0.0;
  for (int64_t k = 0; k < nnz; ++k) Av[rows[k] - 1] += vals[k] * v[cols[k] - 1];
}
// NLPModels coo_sym_prod!
void oracle_coo_sym_prod(int64_t nnz, const int64_t* rows, const int64_t* cols, const double* vals,
                         const double* v, int64_t nout, double* Av) {
  for (int64_t i = 0; i < nout; ++i) Av[i] = 0.0;
  for (int64_t k = 0; k < nnz; ++k) {
    int64_t i = rows[k] - 1, j = cols[k] - 1;
    Av[i] += vals[k] * v[j];
    if (i != j) Av[j] += vals[k] * v[i];
  }
}

// expose the oracle's own tables/coefs so the tests can cross-check the host flattening layer
int oracle_tables(const oracle_problem* P, int32_t* nq, double* xi, double* w, double* phiy, double* dphiy, double* phiu) {
  RefTables T = make_tables(*P);
  *nq = T.nq;
  if (xi) memcpy(xi, T.xi.data(), T.xi.size() * 8);
  if (w) memcpy(w, T.w.data(), T.w.size() * 8);
  if (phiy) memcpy(phiy, T.phiy.data(), T.phiy.size() * 8);
  if (dphiy) memcpy(dphiy, T.dphiy.data(), T.dphiy.size() * 8);
  if (phiu && T.nu) memcpy(phiu, T.phiu.data(), T.phiu.size() * 8);
  return 0;
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
