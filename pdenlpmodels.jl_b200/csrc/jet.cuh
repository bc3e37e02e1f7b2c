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
  for (int k = 0; k < M; ++k) r.g[k] = a.g[k] * b.v + a.v * b.g[k];
  return r;
}
template <int M> PDENLP_HD Jet1<M> operator*(double a, const Jet1<M>& b) {
  Jet1<M> r; r.v = a * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a * b.g[k];
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
  for (int k = 0; k < M; ++k) r.g[k] = c * a.g[k];
  return r;
}
template <int M> PDENLP_HD Jet1<M> jcos(const Jet1<M>& a) {
  Jet1<M> r; r.v = cos(a.v); const double ms = -sin(a.v);
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = ms * a.g[k];
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
// h * v with a structurally zero second derivative: the seeds (and everything affine in them) carry the literal
// constant 0.0 in h after inlining, but IEEE arithmetic forbids folding 0.0 * v, which would keep the VALUE of the
// other factor alive — and with it the gather of x — in kernels that only consume second derivatives (the Hessian
// of a quadratic objective or of a bilinear residual does not depend on x at all).  The explicit test folds at
// compile time for such operands and is a select otherwise.
#ifdef PDENLP_NO_ZMUL  // A/B switch for measurements
PDENLP_HD double zmul(double h, double v) { return h * v; }
#else
PDENLP_HD double zmul(double h, double v) { return h == 0.0 ? 0.0 : h * v; }
#endif

template <int M> PDENLP_HD Jet2<M> operator*(const Jet2<M>& a, const Jet2<M>& b) {
  Jet2<M> r; r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a.g[k] * b.v + a.v * b.g[k];
#pragma unroll
  for (int k = 0; k < M; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l)
      r.h[k * (k + 1) / 2 + l] = zmul(a.h[k * (k + 1) / 2 + l], b.v) + zmul(b.h[k * (k + 1) / 2 + l], a.v) +
                                 a.g[k] * b.g[l] + a.g[l] * b.g[k];
  return r;
}
template <int M> PDENLP_HD Jet2<M> operator*(double a, const Jet2<M>& b) {
  Jet2<M> r; r.v = a * b.v;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = a * b.g[k];
#pragma unroll
  for (int k = 0; k < Jet2<M>::NH; ++k) r.h[k] = a * b.h[k];
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
  for (int k = 0; k < M; ++k) r.g[k] = c * a.g[k];
#pragma unroll
  for (int k = 0; k < M; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l)
      r.h[k * (k + 1) / 2 + l] = c * a.h[k * (k + 1) / 2 + l] + s * (a.g[k] * a.g[l]);
  return r;
}
template <int M> PDENLP_HD Jet2<M> jcos(const Jet2<M>& a) {
  Jet2<M> r; const double c = cos(a.v), ms = -sin(a.v); r.v = c;
#pragma unroll
  for (int k = 0; k < M; ++k) r.g[k] = ms * a.g[k];
#pragma unroll
  for (int k = 0; k < M; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l)
      r.h[k * (k + 1) / 2 + l] = ms * a.h[k * (k + 1) / 2 + l] - c * (a.g[k] * a.g[l]);
  return r;
}

// plain doubles take part in the same generic integrand code
PDENLP_HD double jet_const(double a, const double&) { return a; }
PDENLP_HD double jsinh(double a) { return sinh(a); }
PDENLP_HD double jcos(double a) { return cos(a); }
PDENLP_HD void jet_seed(double& r, double val, int) { r = val; }

}  // namespace pdenlp
