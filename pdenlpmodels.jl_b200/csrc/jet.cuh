// jet.cuh — in-register forward-mode dual numbers.
//
// The reference differentiates each cell functional with ForwardDiff duals as
// wide as the cell-local dof vector (Chunk{length(x)}, nested for Hessians:
// /root/reference/src/test_autodiff.jl:8,26).  Every integrand of the closed set
// depends on the dofs only through the M = 2 + DIM (+ NPARAM) pointwise fields
//   s = (y, d_1 y .. d_DIM y, u [, kappa_1 ..])
// which are LINEAR in the dofs, so the same derivatives are obtained exactly by
// differentiating the pointwise integrand w.r.t. s with M-wide duals and
// contracting with the shape tables (chain rule).  Jet1 carries value+gradient
// (Jacobian kernels); Jet2 adds the packed symmetric Hessian (a second-order
// "hyper-dual", Lagrangian-Hessian kernels).  Everything is fully unrolled so a
// jet lives in registers.
#pragma once

#ifndef PDENLP_HD
#ifdef __CUDACC__
#define PDENLP_HD __host__ __device__ __forceinline__
#else
#define PDENLP_HD inline
#endif
#endif

namespace pdenlp {

// Products with a structurally zero factor.  The seeds (and everything affine in them) carry the literal constants
// 0.0 / 1.0 in their derivative entries after inlining, but IEEE arithmetic forbids folding 0.0 * v (v may be inf or
// NaN).  That keeps dead multiplications alive — and the values they read, e.g. the gather of x in kernels that only
// consume second derivatives — and it turns every structural zero of a derivative into a run-time value which the
// kernels can no longer skip at compile time (the reference-tensor kernels of kvector.cuh branch on them).
// zmul(h, v) = h * v with "h == 0 => 0": folds at compile time for literal operands; for run-time operands the test
// runs on the integer pipe (on B200 the FP64 pipe is the scarce one).  -DPDENLP_NO_ZMUL: plain products (A/B switch).
PDENLP_HD bool jzero(double h) {
#ifdef PDENLP_NO_ZMUL
  (void)h;
  return false;
#elif defined(__CUDA_ARCH__)
  return ((unsigned long long)__double_as_longlong(h) << 1) == 0ull;
#else
  return h == 0.0;
#endif
}
PDENLP_HD double zmul(double h, double v) { return jzero(h) ? 0.0 : h * v; }
PDENLP_HD double zmul2(double a, double b) { return (jzero(a) || jzero(b)) ? 0.0 : a * b; }

// ---------------------------------------------------------------- Jet1
template <int M>
struct Jet1 {
  double v;
  double g[M];
};

template <int M> PDENLP_HD Jet1<M> jet_const(double a, const Jet1<M>&) {
  Jet1<M> r; r.v = a;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = 0.0;
  return r;
}
template <int M> PDENLP_HD void jet_seed(Jet1<M>& r, double val, int k0) {
  r.v = val;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = (k == k0) ? 1.0 : 0.0;
}
// seed with a linear combination of n consecutive independent variables starting at k0 (per-cell geometry: a physical
// gradient component is a combination of the reference gradient slots)
template <int M> PDENLP_HD void jet_seed_lin(Jet1<M>& r, double val, int k0, int n, const double* row) {
  r.v = val;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = (k >= k0 && k < k0 + n) ? row[k - k0 < 0 ? 0 : k - k0] : 0.0;
}
template <int M> PDENLP_HD Jet1<M> operator+(const Jet1<M>& a, const Jet1<M>& b) {
  Jet1<M> r; r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a.g[k] + b.g[k];
  return r;
}
template <int M> PDENLP_HD Jet1<M> operator-(const Jet1<M>& a, const Jet1<M>& b) {
  Jet1<M> r; r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a.g[k] - b.g[k];
  return r;
}
template <int M> PDENLP_HD Jet1<M> operator-(const Jet1<M>& a) {
  Jet1<M> r; r.v = -a.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = -a.g[k];
  return r;
}
template <int M> PDENLP_HD Jet1<M> operator*(const Jet1<M>& a, const Jet1<M>& b) {
  Jet1<M> r; r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(a.g[k], b.v) + zmul(b.g[k], a.v);
  return r;
}
template <int M> PDENLP_HD Jet1<M> operator*(double a, const Jet1<M>& b) {
  Jet1<M> r; r.v = a * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(b.g[k], a);
  return r;
}
template <int M> PDENLP_HD Jet1<M> operator*(const Jet1<M>& b, double a) { return a * b; }
template <int M> PDENLP_HD Jet1<M> operator+(const Jet1<M>& a, double b) { Jet1<M> r = a; r.v += b; return r; }
template <int M> PDENLP_HD Jet1<M> operator+(double b, const Jet1<M>& a) { return a + b; }
template <int M> PDENLP_HD Jet1<M> operator-(const Jet1<M>& a, double b) { Jet1<M> r = a; r.v -= b; return r; }
template <int M> PDENLP_HD Jet1<M> operator-(double b, const Jet1<M>& a) { return (-a) + b; }
template <int M> PDENLP_HD Jet1<M> jsinh(const Jet1<M>& a) {
  Jet1<M> r; r.v = sinh(a.v); const double c = cosh(a.v);
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(a.g[k], c);
  return r;
}
template <int M> PDENLP_HD Jet1<M> jcos(const Jet1<M>& a) {
  Jet1<M> r; r.v = cos(a.v); const double ms = -sin(a.v);
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(a.g[k], ms);
  return r;
}

// ---------------------------------------------------------------- Jet2
template <int M>
struct Jet2 {
  static constexpr int NH = M * (M + 1) / 2;
  double v;
  double g[M];
  double h[NH];  // packed lower triangle: (k,l), k>=l at k(k+1)/2 + l
  PDENLP_HD static constexpr int idx(int k, int l) { return k >= l ? k * (k + 1) / 2 + l : l * (l + 1) / 2 + k; }
  PDENLP_HD double hess(int k, int l) const { return h[idx(k, l)]; }
};

template <int M> PDENLP_HD Jet2<M> jet_const(double a, const Jet2<M>&) {
  Jet2<M> r; r.v = a;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = 0.0;
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = 0.0;
  return r;
}
template <int M> PDENLP_HD void jet_seed(Jet2<M>& r, double val, int k0) {
  r.v = val;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = (k == k0) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = 0.0;
}
template <int M> PDENLP_HD void jet_seed_lin(Jet2<M>& r, double val, int k0, int n, const double* row) {
  r.v = val;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = (k >= k0 && k < k0 + n) ? row[k - k0 < 0 ? 0 : k - k0] : 0.0;
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = 0.0;
}
template <int M> PDENLP_HD Jet2<M> operator+(const Jet2<M>& a, const Jet2<M>& b) {
  Jet2<M> r; r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a.g[k] + b.g[k];
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = a.h[k] + b.h[k];
  return r;
}
template <int M> PDENLP_HD Jet2<M> operator-(const Jet2<M>& a, const Jet2<M>& b) {
  Jet2<M> r; r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a.g[k] - b.g[k];
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = a.h[k] - b.h[k];
  return r;
}
template <int M> PDENLP_HD Jet2<M> operator-(const Jet2<M>& a) {
  Jet2<M> r; r.v = -a.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = -a.g[k];
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = -a.h[k];
  return r;
}
template <int M> PDENLP_HD Jet2<M> operator*(const Jet2<M>& a, const Jet2<M>& b) {
  Jet2<M> r; r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(a.g[k], b.v) + zmul(b.g[k], a.v);
#pragma unroll
  for (int k = 0; k < M; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l)
      r.h[k * (k + 1) / 2 + l] = zmul(a.h[k * (k + 1) / 2 + l], b.v) + zmul(b.h[k * (k + 1) / 2 + l], a.v) +
                                 zmul2(a.g[k], b.g[l]) + zmul2(a.g[l], b.g[k]);
  return r;
}
template <int M> PDENLP_HD Jet2<M> operator*(double a, const Jet2<M>& b) {
  Jet2<M> r; r.v = a * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(b.g[k], a);
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = zmul(b.h[k], a);
  return r;
}
template <int M> PDENLP_HD Jet2<M> operator*(const Jet2<M>& b, double a) { return a * b; }
template <int M> PDENLP_HD Jet2<M> operator+(const Jet2<M>& a, double b) { Jet2<M> r = a; r.v += b; return r; }
template <int M> PDENLP_HD Jet2<M> operator+(double b, const Jet2<M>& a) { return a + b; }
template <int M> PDENLP_HD Jet2<M> operator-(const Jet2<M>& a, double b) { Jet2<M> r = a; r.v -= b; return r; }
template <int M> PDENLP_HD Jet2<M> operator-(double b, const Jet2<M>& a) { return (-a) + b; }
template <int M> PDENLP_HD Jet2<M> jsinh(const Jet2<M>& a) {
  Jet2<M> r; const double s = sinh(a.v), c = cosh(a.v); r.v = s;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(a.g[k], c);
#pragma unroll
  for (int k = 0; k < M; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l)
      r.h[k * (k + 1) / 2 + l] = zmul(a.h[k * (k + 1) / 2 + l], c) + zmul(zmul2(a.g[k], a.g[l]), s);
  return r;
}
template <int M> PDENLP_HD Jet2<M> jcos(const Jet2<M>& a) {
  Jet2<M> r; const double c = cos(a.v), ms = -sin(a.v); r.v = c;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = zmul(a.g[k], ms);
#pragma unroll
  for (int k = 0; k < M; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l)
      r.h[k * (k + 1) / 2 + l] = zmul(a.h[k * (k + 1) / 2 + l], ms) - zmul(zmul2(a.g[k], a.g[l]), c);
  return r;
}

// plain doubles take part in the same generic integrand code
PDENLP_HD double jet_const(double a, const double&) { return a; }
PDENLP_HD double jsinh(double a) { return sinh(a); }
PDENLP_HD double jcos(double a) { return cos(a); }
PDENLP_HD void jet_seed(double& r, double val, int) { r = val; }
PDENLP_HD void jet_seed_lin(double& r, double val, int, int, const double*) { r = val; }

}  // namespace pdenlp
