// integrands.cuh — the closed, named set of weak-form integrands.
//
// Each integrand is written ONCE, generically over the scalar type T (double,
// Jet1, Jet2), as a function of the pointwise fields
//     s[SY] = y, s[SG+d] = d_d y, s[SU] = u, s[SK+k] = kappa_k
// (several state / control fields: see MFIdx at the end of the file)
// f(...)  : objective density
// res(...) : weak residual density, linear in the test function (v, grad v):
//            res = R[0]*v + sum_d R[1+d]*d_d v  (+ R[1+DIM]*1 when HAS_CONST:
//            a term of the integrand that does not multiply v is added to every
//            test component by Gridap's broadcasting — burger1d_param.jl:45).
// Formulas follow the reference problem files cited in include/pdenlp_cuda.h.
#pragma once
#include "jet.cuh"

namespace pdenlp {

struct PointCtx {
  double target;      // yd / sol at the quadrature point
  double source;      // h / f at the quadrature point
  const double* par;  // scalar parameters (alpha, nu, ...)
};

template <int DIM_, int NP_>
struct FieldIdx {
  static constexpr int DIM = DIM_, NP = NP_;
  static constexpr int NFY = 1, NFU = 1;  // one state field, (at most) one control field
  static constexpr bool NO_FE_OBJ = false;  // true: NoFETerm objective fk(kappa), no integral (see PoissonParam)
  // true: d res/d(y, grad y, u) and the FE-field Hessian of f are the same in EVERY cell of a uniform mesh — they depend
  // neither on the state nor on the coefficient fields (an affine residual with constant coefficients, a quadratic
  // objective); kappa may enter.  Selects the template-tile path of the coordinate kernels (kernels.cuh); like the
  // other shortcut traits it is verified by the create-time self-check.
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = false;
  static constexpr int SY = 0, SG = 1, SU = 1 + DIM_, SK = 2 + DIM_;
  static constexpr int M = 2 + DIM_ + NP_;
};

// 0.5*(yd - y)*(yd - y) + 0.5*alpha*u*u — shared by MEMBRANE / PB / BURGER
template <class T, class I>
PDENLP_HD T tracking_f(const PointCtx& c, const T* s) {
  T d = c.target - s[I::SY];
  return 0.5 * d * d + (0.5 * c.par[0]) * s[I::SU] * s[I::SU];
}

template <int DIM_>
struct Membrane : FieldIdx<DIM_, 0> {
  using I = FieldIdx<DIM_, 0>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = false;  // derivatives do not read coefficient fields
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = true;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) { return tracking_f<T, I>(c, s); }
  // grad(v).grad(y) - v*u - v*h
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = -s[I::SU] - c.source;
#pragma unroll
    for (int d = 0; d < DIM_; ++d) R[1 + d] = s[I::SG + d];
  }
};

template <int DIM_>
struct PoissonBoltzmann : FieldIdx<DIM_, 0> {
  using I = FieldIdx<DIM_, 0>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = false;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = false;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) { return tracking_f<T, I>(c, s); }
  // grad(v).grad(y) + sinh(y)*v - u*v - v*h
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = jsinh(s[I::SY]) - s[I::SU] - c.source;
#pragma unroll
    for (int d = 0; d < DIM_; ++d) R[1 + d] = s[I::SG + d];
  }
};

// TEST ONLY (PDENLP_SELFTEST_WRONG_TRAIT): the Poisson-Boltzmann integrand with both shortcut traits WRONGLY set — the
// sinh term makes the residual non-affine, so its constraint FE Hessian block is not zero and its derivatives differ
// from point to point.  pdenlp_create must refuse it (create-time self-check), tests/test_gpu_error_paths.py.
template <int DIM_>
struct WrongTraitPB : PoissonBoltzmann<DIM_> {
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;
  static constexpr bool RES_FE_HESSIAN_ZERO = true;
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
};

struct BurgerLin : FieldIdx<1, 0> {
  using I = FieldIdx<1, 0>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = true;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) { return tracking_f<T, I>(c, s); }
  // -nu*grad(v).grad(y) + (grad(y).1)*v - v*u - v*h
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = s[I::SG] - s[I::SU] - c.source;
    R[1] = (-c.par[1]) * s[I::SG];
  }
};

struct BurgerNl : FieldIdx<1, 1> {
  using I = FieldIdx<1, 1>;
  static constexpr bool HAS_CONST = true;
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = false;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = false;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    T t = s[I::SK] - c.par[2];
    return tracking_f<T, I>(c, s) + 0.5 * t * t;
  }
  // nu*grad(v).grad(y) - k1 + v*y*(grad(y).1) - v*u - v*h
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = s[I::SY] * s[I::SG] - s[I::SU] - c.source;
    R[1] = c.par[1] * s[I::SG];
    R[2] = -s[I::SK];
  }
};

template <int DIM_, bool HAS_U>
struct KappaPoisson : FieldIdx<DIM_, 2> {
  using I = FieldIdx<DIM_, 2>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = true;  // d res / d k2 = -v*f
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = true;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    T d = c.target - s[I::SY];
    T a = s[I::SK] - 1.0, b = s[I::SK + 1] - 1.0;
    return 0.5 * d * d + 0.5 * a * a + 0.5 * b * b;
  }
  // k1*grad(v).grad(y) - v*f*k2 [- v*u]
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = (-c.source) * s[I::SK + 1];
    if (HAS_U) R[0] = R[0] - s[I::SU];
#pragma unroll
    for (int d = 0; d < DIM_; ++d) R[1 + d] = s[I::SK] * s[I::SG + d];
  }
};

// NoFETerm objective (src/additional_obj_terms.jl:59-106): obj = fk(kappa), not an integral.  The density f is
// identically zero (so the FE parts of grad!/hprod! vanish), the objective Hessian has NO FE block and its
// kappa-block is the p x p lower triangle; fk is differentiated by a one-thread kernel (k_nofe).
// test/problems/poissonparam.jl:35-44: fk = 0.5*dot(k .- 1, k .- 1), res = k1*grad(v).grad(y) - v*f
template <int DIM_>
struct PoissonParam : FieldIdx<DIM_, 1> {
  using I = FieldIdx<DIM_, 1>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;
  static constexpr bool RES_FE_HESSIAN_ZERO = true;
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
  static constexpr bool NO_FE_OBJ = true;
  template <class T> PDENLP_HD static T f(const PointCtx&, const T* s) { return jet_const(0.0, s[0]); }
  template <class T> PDENLP_HD static T fk(const double*, const T* k) {
    T d = k[0] - 1.0;
    return 0.5 * d * d;
  }
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = jet_const(-c.source, s[0]);
#pragma unroll
    for (int d = 0; d < DIM_; ++d) R[1 + d] = s[I::SK] * s[I::SG + d];
  }
};

// test/problems/poissonmixed2.jl:35-43: f = k1*(sol - y)^2 + 0.5*(k2 - 1)^2 (inde = false), res as KappaPoisson
template <int DIM_>
struct KappaPoisson2 : FieldIdx<DIM_, 2> {
  using I = FieldIdx<DIM_, 2>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = true;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = false;  // d2f/dy2 = 2*k1 is, but the kappa-y cross terms are not
  static constexpr bool RES_FE_HESSIAN_ZERO = true;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    T d = c.target - s[I::SY];
    T b = s[I::SK + 1] - 1.0;
    return s[I::SK] * d * d + 0.5 * b * b;
  }
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    R[0] = (-c.source) * s[I::SK + 1];
#pragma unroll
    for (int d = 0; d < DIM_; ++d) R[1 + d] = s[I::SK] * s[I::SG + d];
  }
};

// ---- unconstrained problems: objective only, the residual is identically zero (ncon = 0 on the
// host side; the constraint entry points are never called for these models)
template <int DIM_>
struct PenalizedPoisson : FieldIdx<DIM_, 0> {
  using I = FieldIdx<DIM_, 0>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = true;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
  // 0.5 * grad(y).grad(y) - w * y
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    T a = s[I::SG] * s[I::SG];
#pragma unroll
    for (int d = 1; d < DIM_; ++d) a = a + s[I::SG + d] * s[I::SG + d];
    return 0.5 * a - c.source * s[I::SY];
  }
  template <class T> PDENLP_HD static void res(const PointCtx&, const T* s, T* R) {
#pragma unroll
    for (int t = 0; t < 1 + DIM_; ++t) R[t] = jet_const(0.0, s[0]);
  }
};

template <int DIM_>
struct BasicUnconstrained : FieldIdx<DIM_, 0> {
  using I = FieldIdx<DIM_, 0>;
  static constexpr bool HAS_CONST = false;
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = true;  // d res/d(y,grad y,u) and the FE-field Hessians are the same at every point of a cell
  static constexpr bool RES_FE_HESSIAN_ZERO = true;  // res is affine in (y, grad y, u): the constraint FE Hessian block is identically zero
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = true;
  // 0.5 * (ubis - u) * (ubis - u) + 0.5 * y * y
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    T d = c.target - s[I::SU];
    return 0.5 * d * d + 0.5 * s[I::SY] * s[I::SY];
  }
  template <class T> PDENLP_HD static void res(const PointCtx&, const T* s, T* R) {
#pragma unroll
    for (int t = 0; t < 1 + DIM_; ++t) R[t] = jet_const(0.0, s[0]);
  }
};

// ---- multi-field state / control spaces (MultiFieldFESpace).
// Pointwise fields: state field f occupies slots f*NB .. f*NB+DIM (value, gradient), NB = 1 + DIM;
// control field g occupies slot SU + g.  Residual slots are laid out the same way over the TEST
// fields: R[f*NB] multiplies v_f, R[f*NB+1+d] multiplies d_d v_f.
template <int DIM_, int NFY_, int NFU_>
struct MFIdx {
  static constexpr int DIM = DIM_, NP = 0, NFY = NFY_, NFU = NFU_;
  static constexpr int NB = 1 + DIM_;
  static constexpr int SY = 0, SG = 1, SU = NFY_ * NB, SK = SU + (NFU_ > 0 ? NFU_ : 1);
  static constexpr int M = SK;
  static constexpr bool HAS_CONST = false;
  static constexpr bool FE_DERIVS_POINT_INDEPENDENT = false;
  static constexpr bool FE_DERIVS_CELL_INDEPENDENT = false;
  static constexpr bool NO_FE_OBJ = false;
  PDENLP_HD static constexpr int Y(int f) { return f * NB; }
  PDENLP_HD static constexpr int G(int f, int d = 0) { return f * NB + 1 + d; }
  PDENLP_HD static constexpr int U(int g) { return SU + g; }
  template <class T> PDENLP_HD static void zero_res(const T* s, T* R) {
#pragma unroll
    for (int t = 0; t < NFY_ * NB; ++t) R[t] = jet_const(0.0, s[0]);
  }
};

// test/problems/controlsir.jl:20-33: fields (I, S); par = (a, b, mu)
struct ControlSir : MFIdx<1, 2, 0> {
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool RES_FE_HESSIAN_ZERO = false;
  using X = MFIdx<1, 2, 0>;
  template <class T> PDENLP_HD static T f(const PointCtx&, const T* s) { return 0.5 * s[X::Y(0)] * s[X::Y(0)]; }
  // dt(I,p) + dt(S,q) - p*(a*S*I - b*I) - q*(mu - b*I - a*S*I)
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    const double a = c.par[0], b = c.par[1], mu = c.par[2];
    const T& I = s[X::Y(0)]; const T& S = s[X::Y(1)];
    X::zero_res(s, R);
    const T aSI = a * S * I;
    R[X::Y(0)] = s[X::G(0)] - (aSI - b * I);
    R[X::Y(1)] = s[X::G(1)] - (mu - b * I - aSI);
  }
};

// test/problems/cellincrease.jl:15-40: fields (cf, pf), control u in L2; par = (kp, kr)
struct CellIncrease : MFIdx<1, 2, 1> {
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool RES_FE_HESSIAN_ZERO = false;
  using X = MFIdx<1, 2, 1>;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) { return c.par[0] * s[X::Y(1)]; }
  // dt(cf,p) + dt(pf,q) - p*(kp*pf*(1-cf) - kr*cf*(1-cf-pf)) + q*(u*kr*cf*(1-cf-pf) - kp*pf*pf)
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    const double kp = c.par[0], kr = c.par[1];
    const T& cf = s[X::Y(0)]; const T& pf = s[X::Y(1)]; const T& u = s[X::U(0)];
    X::zero_res(s, R);
    const T g = kr * cf * (1.0 - cf - pf);
    R[X::Y(0)] = s[X::G(0)] - (kp * pf * (1.0 - cf) - g);
    R[X::Y(1)] = s[X::G(1)] + (u * g - kp * pf * pf);
  }
};

// test/problems/dynamicsir.jl:36-58: fields (I, S), controls (bf, cf) in H1; target = w0(t), par = (w1, w2)
struct DynamicSir : MFIdx<1, 2, 2> {
  static constexpr bool NEEDS_COEF_DERIV = true;  // d f / d(bf, cf) reads w0 = coef target
  static constexpr bool RES_FE_HESSIAN_ZERO = false;
  using X = MFIdx<1, 2, 2>;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    T a = c.target * s[X::U(0)] - c.par[0];
    T b = c.target * s[X::U(1)] - c.par[1];
    return 0.5 * a * a + 0.5 * b * b;
  }
  // dt(I,p) + dt(S,q) - p*(bf*S*I - cf*I) + q*bf*S*I
  template <class T> PDENLP_HD static void res(const PointCtx&, const T* s, T* R) {
    const T& I = s[X::Y(0)]; const T& S = s[X::Y(1)];
    X::zero_res(s, R);
    const T bSI = s[X::U(0)] * S * I;
    R[X::Y(0)] = s[X::G(0)] - (bSI - s[X::U(1)] * I);
    R[X::Y(1)] = s[X::G(1)] + bSI;
  }
};

// test/problems/torebrachistochrone.jl:25-28: fields (phi, theta), unconstrained; par = (a, c)
struct ToreBrachistochrone : MFIdx<1, 2, 0> {
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool RES_FE_HESSIAN_ZERO = true;
  using X = MFIdx<1, 2, 0>;
  template <class T> PDENLP_HD static T f(const PointCtx& c, const T* s) {
    const double a = c.par[0];
    T w = c.par[1] + a * jcos(s[X::Y(0)]);
    return (a * a) * s[X::G(0)] * s[X::G(0)] + w * w * s[X::G(1)] * s[X::G(1)];
  }
  template <class T> PDENLP_HD static void res(const PointCtx&, const T* s, T* R) { X::zero_res(s, R); }
};

// test/problems/sis.jl:21-34: fields (I, S), objective NoFETerm() = 0; par = (a, b)
struct Sis : MFIdx<1, 2, 0> {
  static constexpr bool NEEDS_COEF_DERIV = false;
  static constexpr bool RES_FE_HESSIAN_ZERO = false;
  static constexpr bool NO_FE_OBJ = true;
  using X = MFIdx<1, 2, 0>;
  template <class T> PDENLP_HD static T f(const PointCtx&, const T* s) { return jet_const(0.0, s[0]); }
  template <class T> PDENLP_HD static T fk(const double*, const T* k) { return jet_const(0.0, k[0]); }
  // dt(I,p) + dt(S,q) - p*(a*S*I - b*I) - q*(b*I - a*S*I)
  template <class T> PDENLP_HD static void res(const PointCtx& c, const T* s, T* R) {
    const double a = c.par[0], b = c.par[1];
    const T& I = s[X::Y(0)]; const T& S = s[X::Y(1)];
    X::zero_res(s, R);
    const T t = a * S * I - b * I;
    R[X::Y(0)] = s[X::G(0)] - t;
    R[X::Y(1)] = s[X::G(1)] + t;
  }
};

}  // namespace pdenlp
