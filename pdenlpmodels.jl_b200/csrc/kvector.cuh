// kvector.cuh — reference-tensor ("K-path") versions of the dof-indexed callbacks and of the dense kappa blocks.
//
// For an element with several quadrature points the generic kernels of kernels.cuh evaluate the integrand's jets
// at every point (config 5, 3-D Q1 with 8 points and 7 pointwise fields: ~1200 FP64 FMAs per cell for jprod!,
// several thousand for hprod!) — on B200 that is FP64-pipe bound (64 FMA/clk/SM), far below the HBM roofline.
// Integrands whose residual is AFFINE in the FE fields with a point-independent linear part,
//     R_t(s, kappa; q) = A_t(kappa; q) + sum_k D_tk(kappa) s_k ,        s = (y, grad y, u),
// (traits FE_DERIVS_POINT_INDEPENDENT && RES_FE_HESSIAN_ZERO: membrane, burger1d, kappa-Poisson, poissonparam)
// have element residuals  c = a(kappa) + G(kappa) z  with
//     G[i][a] = sum_{t,k} D_tk K[t][k][a][i],   K = sum_q w_q T_i[t](q) B_a[k](q)   (KTab, pre-integrated once),
//     a_i     = sum_q w_q sum_t T_i[t](q) A_t(kappa; q),
// so every product is a handful of 8x8 mat-vecs with tensors held in shared memory:
//     cons!   c          = a + G z
//     jprod!  J v        = G v_z + sum_m v_km (da/dk_m + dG/dk_m z)
//     jtprod! J' w       = (G' w ; w.(da/dk_m + dG/dk_m z))
//     hprod!  H v        = obj_weight * C-weighted mass/stiffness mat-vecs (C = point-independent Hessian of f)
//                          + (sum_m v_km dG/dk_m' lambda ; lambda.(dG/dk_m v_z) + sum_n v_kn lambda.d2a/dk_m dk_n)
//     kappa blocks of jac_coord!/hess_coord!: da/dk_m + dG/dk_m z   and   dG/dk_m' lambda, lambda.d2a/dk dk
// D, dD/dkappa, A and its kappa-derivatives come from ONE evaluation of the generic integrand code with jets seeded
// at s = 0 (exact because R is affine in s; D is assumed at most linear in kappa, true for the closed set).
// The generic per-point kernels stay compiled in: pdenlp_create runs both on a sample of cells and refuses the
// model if they disagree (a wrong trait cannot silently change results).
#pragma once
#include "kernels.cuh"

namespace pdenlp {

template <class E>
__host__ __device__ constexpr bool kvec_enabled() {
  using I = typename E::Integrand;
  return KTab<E>::ENABLED && I::FE_DERIVS_POINT_INDEPENDENT && I::RES_FE_HESSIAN_ZERO;
}
// which ops have a K-path instance
__host__ __device__ constexpr bool kvec_op(int op) {
  return op == OP_CONS || op == OP_JPROD || op == OP_JTPROD || op == OP_HPROD || op == OP_JAC_KAPPA || op == OP_HESS_KAPPA;
}

// The reference tensors of the K-path vector kernels travel BY VALUE as a __grid_constant__ kernel parameter (up to
// 13.6 KB for the 3-D Q1 element; parameters may be 32 KB since CUDA 12.1): like the shape tables they then live in the
// constant bank and every entry the (fully unrolled, compile-time indexed) mat-vecs use is an immediate c[0x0][..]
// operand of a DFMA — no LDS, no address arithmetic (a shared-memory copy costs one LDS per FMA: 160 of the
// 1500 instructions per tile of config 5's jprod!, on top of the staging and the barrier).
template <class E>
struct KConst {
  double v[KTab<E>::TOTAL > 0 ? KTab<E>::TOTAL : 1];
};

__host__ __device__ constexpr bool kvec_tensors_in_smem(int op) {
#ifdef PDENLP_KVEC_SMEM
  return true;
#else
  return op == OP_HPROD || op == OP_HESS_KAPPA;
#endif
}

template <class E>
struct AffineRes {
  static constexpr int NB = E::NB, NK = E::NB + 1, NPS = E::NP > 0 ? E::NP : 1;
  double D[NB][NK];        // dR_t / ds_k, k = (y, grad y, u)
  double D1[NB][NK][NPS];  // d2R_t / ds_k dkappa_m
};

// one jet evaluation of res at s = 0 (kappa seeded at its value)
template <class E, bool WITH_D1>
__device__ __forceinline__ void eval_affine(const PointCtx& pc, const double* __restrict__ x, AffineRes<E>& A) {
  using I = typename E::Integrand;
  constexpr int NB = E::NB, NK = NB + 1;
  if (WITH_D1 && E::NP > 0) {
    using J2 = Jet2<E::M>;
    J2 s[E::M], R[E::NR];
#pragma unroll
    for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + E::NP) ? __ldg(x + (k - E::SK)) : 0.0, k);
    I::res(pc, s, R);
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        A.D[t][k] = R[t].g[k];
#pragma unroll
        for (int m = 0; m < E::NP; ++m) A.D1[t][k][m] = R[t].h[J2::idx(E::SK + m, k)];
      }
  } else {
    Jet1<E::M> s[E::M], R[E::NR];
#pragma unroll
    for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + E::NP) ? __ldg(x + (k - E::SK)) : 0.0, k);
    I::res(pc, s, R);
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        A.D[t][k] = R[t].g[k];
#pragma unroll
        for (int m = 0; m < AffineRes<E>::NPS; ++m) A.D1[t][k][m] = 0.0;
      }
  }
}

// grad-grad sub-block of a coefficient matrix is c * identity (Laplace-type forms): the pre-summed tensor `lap` serves it
template <class E>
__device__ __forceinline__ bool gg_iso(const double (&C)[E::NB][E::NB + 1]) {
  bool iso = E::DIM >= 1;
#pragma unroll
  for (int t = 1; t < E::NB; ++t)
#pragma unroll
    for (int k = 1; k < E::NB; ++k) iso = iso && (t == k ? (t == 1 || C[t][k] == C[1][1]) : jzero(C[t][k]));
  return iso;
}

// out[i] += sum_{t,k} sum_j yy(t,k,j,i) * (C0[t][k] * v0[j] + C1[t][k] * v1[j])         (rows i: test dofs)
template <class E, bool TWO>
__device__ __forceinline__ void apply_yy(const double* __restrict__ Ks, const double (&C0)[E::NB][E::NB + 1], const double* v0,
                                         const double (&C1)[E::NB][E::NB + 1], const double* v1, double* out) {
  using KT = KTab<E>;
  const bool iso = gg_iso<E>(C0) && (!TWO || gg_iso<E>(C1));
  if (iso && (!jzero(C0[1][1]) || (TWO && !jzero(C1[1][1])))) {
#pragma unroll
    for (int j = 0; j < E::NY; ++j) {
      const double w = TWO ? zmul(C0[1][1], v0[j]) + zmul(C1[1][1], v1[j]) : C0[1][1] * v0[j];
#pragma unroll
      for (int i = 0; i < E::NY; ++i) out[i] = fma(Ks[KT::lap(j, i)], w, out[i]);
    }
  }
#pragma unroll
  for (int t = 0; t < E::NB; ++t)
#pragma unroll
    for (int k = 0; k < E::NB; ++k) {
      const bool gg = t >= 1 && k >= 1;
      if ((!gg || !iso) && (!jzero(C0[t][k]) || (TWO && !jzero(C1[t][k])))) {
#pragma unroll
        for (int j = 0; j < E::NY; ++j) {
          const double w = TWO ? zmul(C0[t][k], v0[j]) + zmul(C1[t][k], v1[j]) : C0[t][k] * v0[j];
#pragma unroll
          for (int i = 0; i < E::NY; ++i) out[i] = fma(Ks[KT::yy(t, k, j, i)], w, out[i]);
        }
      }
    }
}
// the control column k = SU:  out[i] += sum_t sum_b yu(t,b,i) * (C0[t][SU] * u0[b] + C1[t][SU] * u1[b])
template <class E, bool TWO>
__device__ __forceinline__ void apply_yu(const double* __restrict__ Ks, const double (&C0)[E::NB][E::NB + 1], const double* u0,
                                         const double (&C1)[E::NB][E::NB + 1], const double* u1, double* out) {
  using KT = KTab<E>;
  if (E::NU > 0) {
#pragma unroll
    for (int t = 0; t < E::NB; ++t)
      if (!jzero(C0[t][E::NB]) || (TWO && !jzero(C1[t][E::NB]))) {
#pragma unroll
        for (int b = 0; b < E::NU; ++b) {
          const double w = TWO ? zmul(C0[t][E::NB], u0[b]) + zmul(C1[t][E::NB], u1[b]) : C0[t][E::NB] * u0[b];
#pragma unroll
          for (int i = 0; i < E::NY; ++i) out[i] = fma(Ks[KT::yu(t, b, i)], w, out[i]);
        }
      }
  }
}

// Transposed application.  For every active term (t,k):  p = K[t][k]' w  (p[j] = sum_i yy(t,k,j,i) w[i]);
//   oy[j] += C0[t][k] * p[j];   kacc[m] += C1[t][k][m] * (p . zy)
// and for the control column:  pu[b] = sum_i yu(t,b,i) w[i];  ou[b] += C0[t][SU] * pu[b];  kacc[m] += C1[t][SU][m] * (pu . zu)
// WITH_OUT / WITH_K select which of the two uses are wanted.
template <class E, bool WITH_OUT, bool WITH_K>
__device__ __forceinline__ void apply_T(const double* __restrict__ Ks, const double (&C0)[E::NB][E::NB + 1],
                                        const double (&C1)[E::NB][E::NB + 1][AffineRes<E>::NPS], const double* w,
                                        const double* zy, const double* zu, double* oy, double* ou, double* kacc) {
  using KT = KTab<E>;
  constexpr int NPS = AffineRes<E>::NPS;
  auto use_y = [&](const double* p, double c0, const double* c1) {
    if (WITH_OUT && !jzero(c0)) {
#pragma unroll
      for (int j = 0; j < E::NY; ++j) oy[j] = fma(c0, p[j], oy[j]);
    }
    if (WITH_K && E::NP > 0) {
      double dot = 0.0;
#pragma unroll
      for (int j = 0; j < E::NY; ++j) dot = fma(p[j], zy[j], dot);
#pragma unroll
      for (int m = 0; m < E::NP; ++m)
        if (!jzero(c1[m])) kacc[m] = fma(c1[m], dot, kacc[m]);
    }
  };
  bool iso = gg_iso<E>(C0);
  bool any1 = false;
  if (WITH_K && E::NP > 0) {
#pragma unroll
    for (int m = 0; m < E::NP; ++m) {
      bool im = E::DIM >= 1;
#pragma unroll
      for (int t = 1; t < E::NB; ++t)
#pragma unroll
        for (int k = 1; k < E::NB; ++k) im = im && (t == k ? (t == 1 || C1[t][k][m] == C1[1][1][m]) : jzero(C1[t][k][m]));
      iso = iso && im;
      any1 = any1 || !jzero(C1[1][1][m]);
    }
  }
  if (iso && ((WITH_OUT && !jzero(C0[1][1])) || any1)) {
    double p[E::NY];
#pragma unroll
    for (int j = 0; j < E::NY; ++j) {
      double a = 0.0;
#pragma unroll
      for (int i = 0; i < E::NY; ++i) a = fma(Ks[KT::lap(j, i)], w[i], a);
      p[j] = a;
    }
    use_y(p, C0[1][1], C1[1][1]);
  }
#pragma unroll
  for (int t = 0; t < E::NB; ++t)
#pragma unroll
    for (int k = 0; k < E::NB; ++k) {
      const bool gg = t >= 1 && k >= 1;
      bool act = WITH_OUT && !jzero(C0[t][k]);
      if (WITH_K) {
#pragma unroll
        for (int m = 0; m < E::NP; ++m) act = act || !jzero(C1[t][k][m]);
      }
      if ((!gg || !iso) && act) {
        double p[E::NY];
#pragma unroll
        for (int j = 0; j < E::NY; ++j) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < E::NY; ++i) a = fma(Ks[KT::yy(t, k, j, i)], w[i], a);
          p[j] = a;
        }
        use_y(p, C0[t][k], C1[t][k]);
      }
    }
  if (E::NU > 0) {
#pragma unroll
    for (int t = 0; t < E::NB; ++t) {
      bool act = WITH_OUT && !jzero(C0[t][E::NB]);
      if (WITH_K) {
#pragma unroll
        for (int m = 0; m < E::NP; ++m) act = act || !jzero(C1[t][E::NB][m]);
      }
      if (act) {
        double dot = 0.0;
#pragma unroll
        for (int b = 0; b < E::NU; ++b) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < E::NY; ++i) a = fma(Ks[KT::yu(t, b, i)], w[i], a);
          if (WITH_OUT && !jzero(C0[t][E::NB])) ou[b] = fma(C0[t][E::NB], a, ou[b]);
          if (WITH_K && E::NP > 0) dot = fma(a, zu[b], dot);
        }
        if (WITH_K) {
#pragma unroll
          for (int m = 0; m < NPS; ++m)
            if (!jzero(C1[t][E::NB][m])) kacc[m] = fma(C1[t][E::NB][m], dot, kacc[m]);
        }
      }
    }
  }
}

// a-type element vectors:  acc[c][i] += sum_q w_q sum_t T_i[t](q) X_t^c(q),  X^c = the c-th requested component of
// res(s = 0; q): ORDER 0: the value (1 component); ORDER 1: d/dkappa_m (NP components); ORDER 2: d2/dkappa_m dkappa_n,
// m >= n (NP(NP+1)/2 components).  A (q, t, c) term that vanishes on every lane of the warp is skipped.
// `valid` masks lanes without a cell.
template <class E, int ORDER>
struct AOrder {
  static constexpr int NC = ORDER == 0 ? 1 : (ORDER == 1 ? (E::NP > 0 ? E::NP : 1) : (E::NP > 0 ? E::NP * (E::NP + 1) / 2 : 1));
};
template <class E, int ORDER>
__device__ __forceinline__ void accum_A(const Tables<E>& tab, const DevModel& D, const double* __restrict__ x, int64_t cell, bool valid,
                                        double (&acc)[AOrder<E, ORDER>::NC][E::NY]) {
  using I = typename E::Integrand;
  constexpr int NC = AOrder<E, ORDER>::NC;
  if (ORDER > 0 && E::NP == 0) return;
  // coefficient values of all points up front: independent loads (one memory latency, not NQ dependent ones)
  double tgt[E::NQ], src[E::NQ];
#pragma unroll
  for (int q = 0; q < E::NQ; ++q) {
    tgt[q] = valid ? __ldg(D.coef + (cell * E::NQ + q)) : 0.0;
    src[q] = valid ? __ldg(D.coef + ((D.ncells_coef + cell) * E::NQ + q)) : 0.0;
  }
  double kap[E::NP > 0 ? E::NP : 1];
#pragma unroll
  for (int m = 0; m < E::NP; ++m) kap[m] = __ldg(x + m);
  // (fully unrolled: the table entries By[q][i][t] are immediate operands, and components that are structurally zero
  //  — literal zeros after the jets' compile-time folding — disappear from the code)
#pragma unroll
  for (int q = 0; q < E::NQ; ++q) {
    PointCtx pc;
    pc.par = tab.par; pc.target = tgt[q]; pc.source = src[q];
    const double w = valid ? tab.wdet[q] : 0.0;
    double X[E::NB][NC];
    if (ORDER == 0) {
      double s[E::M], R[E::NR];
#pragma unroll
      for (int k = 0; k < E::M; ++k) s[k] = (k >= E::SK && k < E::SK + E::NP) ? kap[k - E::SK] : 0.0;
      I::res(pc, s, R);
#pragma unroll
      for (int t = 0; t < E::NB; ++t) X[t][0] = zmul(R[t], w);
    } else if (ORDER == 1) {
      Jet1<E::M> s[E::M], R[E::NR];
#pragma unroll
      for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + E::NP) ? kap[k - E::SK] : 0.0, k);
      I::res(pc, s, R);
#pragma unroll
      for (int t = 0; t < E::NB; ++t)
#pragma unroll
        for (int m = 0; m < E::NP; ++m) X[t][m] = zmul(R[t].g[E::SK + m], w);
    } else {
      using J2 = Jet2<E::M>;
      J2 s[E::M], R[E::NR];
#pragma unroll
      for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + E::NP) ? kap[k - E::SK] : 0.0, k);
      I::res(pc, s, R);
#pragma unroll
      for (int t = 0; t < E::NB; ++t) {
        int c = 0;
#pragma unroll
        for (int n = 0; n < E::NP; ++n)
#pragma unroll
          for (int m = n; m < E::NP; ++m) { X[t][c] = zmul(R[t].h[J2::idx(E::SK + m, E::SK + n)], w); ++c; }
      }
    }
    // (structural zeros are literal zeros after the jets' compile-time folding: their terms vanish from the code)
#pragma unroll
    for (int t = 0; t < E::NB; ++t)
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (!jzero(X[t][c])) {
#pragma unroll
          for (int i = 0; i < E::NY; ++i) acc[c][i] = fma(tab.By[q][i][t], X[t][c], acc[c][i]);
        }
  }
}

#ifndef PDENLP_KVEC_MINB
#define PDENLP_KVEC_MINB 2
#endif
template <class E, int OP>
__global__ void __launch_bounds__(kThreads, PDENLP_KVEC_MINB)
k_vector_K(const __grid_constant__ Tables<E> tab, const __grid_constant__ KConst<E> kc, const DevModel D, const double* __restrict__ x,
           const double* __restrict__ lambda, const double* __restrict__ v, double obj_weight, int64_t prows,
           double* __restrict__ out, double* __restrict__ partials) {
  using I = typename E::Integrand;
  using AR = AffineRes<E>;
  constexpr int NB = E::NB, NK = NB + 1, NPS = AR::NPS, NP = E::NP;
  // Operand source of the reference tensors, per op by measurement (config 5, ms): constant bank (LDCU + DFMA with a
  // uniform-register operand) jprod! 0.139 / jtprod! 0.156 / hprod! 0.238; shared memory (LDS per FMA) 0.180 / 0.174 / 0.122
  extern __shared__ __align__(16) double smem[];
  const double* Ks = kc.v;
  if constexpr (kvec_tensors_in_smem(OP)) Ks = load_ktab<E>(D, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NKK = NP > 0 ? NP * (NP + 1) / 2 : 1;
  double ksum[NKK];  // kappa entries (JTPROD/HPROD: NP; HESS_KAPPA: the kappa-kappa corner)
#pragma unroll
  for (int k = 0; k < NKK; ++k) ksum[k] = 0.0;
  constexpr bool NEED_Z = OP == OP_CONS || ((OP == OP_JPROD || OP == OP_JTPROD || OP == OP_JAC_KAPPA) && NP > 0);
  constexpr bool NEED_D1 = NP > 0 && (OP == OP_JPROD || OP == OP_JTPROD || OP == OP_HPROD || OP == OP_JAC_KAPPA || OP == OP_HESS_KAPPA);
  double vk[NPS];
#pragma unroll
  for (int m = 0; m < NPS; ++m) vk[m] = (NP > 0 && (OP == OP_JPROD || OP == OP_HPROD)) ? __ldg(v + m) : 0.0;

  const int nwarps = blockDim.x >> 5;
  const int64_t vstride = (int64_t)gridDim.x * nwarps;
  Cell<E> c, nx;
  {
    const int64_t t0 = (int64_t)blockIdx.x * nwarps + warp;
    const int64_t c0 = t0 * 32 + lane;
    const bool v0 = t0 < D.ntiles && c0 < D.ncells;
    load_ids<E>(D, v0 ? c0 : 0, v0, nx);
  }
  for (int64_t tile = (int64_t)blockIdx.x * nwarps + warp; tile < D.ntiles; tile += vstride) {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) c.gy[a] = nx.gy[a];
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) c.gu[b] = nx.gu[b];
    {
      const int64_t c1 = (tile + vstride) * 32 + lane;
      const bool v1 = (tile + vstride) < D.ntiles && c1 < D.ncells;
      load_ids<E>(D, v1 ? c1 : 0, v1, nx);
    }
    if (NEED_Z) gather_state<E>(D, x, valid, c);
    AR A;
    {
      const PointGeo<E> pc0 = point_ctx<E>(D, tab, cell, valid, 0, I::NEEDS_COEF_DERIV);
      eval_affine<E, NEED_D1>(pc0, x, A);
    }
    double ey[E::NY], eu[E::NUS];
#pragma unroll
    for (int a = 0; a < E::NY; ++a) ey[a] = 0.0;
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) eu[b] = 0.0;

    if (OP == OP_CONS) {
      double a0[1][E::NY];
#pragma unroll
      for (int i = 0; i < E::NY; ++i) a0[0][i] = 0.0;
      accum_A<E, 0>(tab, D, x, cell, valid, a0);
#pragma unroll
      for (int i = 0; i < E::NY; ++i) ey[i] = a0[0][i];
      apply_yy<E, false>(Ks, A.D, c.zy, A.D, c.zy, ey);
      apply_yu<E, false>(Ks, A.D, c.zu, A.D, c.zu, ey);
    } else if (OP == OP_JPROD || OP == OP_JAC_KAPPA) {
      // J v = G v_z + sum_m v_km (a1_m + G1_m z);   JAC_KAPPA: column m = a1_m + G1_m z
      double a1[NPS][E::NY];
#pragma unroll
      for (int m = 0; m < NPS; ++m)
#pragma unroll
        for (int i = 0; i < E::NY; ++i) a1[m][i] = 0.0;
      accum_A<E, 1>(tab, D, x, cell, valid, a1);
      if (OP == OP_JPROD) {
        double vy[E::NY], vu[E::NUS];
        gather_free<E::NY>(v + D.nparam, c.gy, vy);
        if (E::NU > 0) gather_free<E::NUS>(v + D.nparam + D.nfree_y, c.gu, vu);
        double C1[NB][NK];  // sum_m v_km dD/dk_m
#pragma unroll
        for (int t = 0; t < NB; ++t)
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            double a = 0.0;
#pragma unroll
            for (int m = 0; m < NP; ++m) a += zmul(A.D1[t][k][m], vk[m]);
            C1[t][k] = a;
          }
        if (NP > 0) {
          apply_yy<E, true>(Ks, A.D, vy, C1, c.zy, ey);
          apply_yu<E, true>(Ks, A.D, vu, C1, c.zu, ey);
#pragma unroll
          for (int m = 0; m < NP; ++m)
#pragma unroll
            for (int i = 0; i < E::NY; ++i) ey[i] += zmul(a1[m][i], vk[m]);
        } else {
          apply_yy<E, false>(Ks, A.D, vy, A.D, vy, ey);
          apply_yu<E, false>(Ks, A.D, vu, A.D, vu, ey);
        }
      } else if (NP > 0) {
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          double C1[NB][NK], col[E::NY];
#pragma unroll
          for (int t = 0; t < NB; ++t)
#pragma unroll
            for (int k = 0; k < NK; ++k) C1[t][k] = A.D1[t][k][m];
#pragma unroll
          for (int i = 0; i < E::NY; ++i) col[i] = a1[m][i];
          apply_yy<E, false>(Ks, C1, c.zy, C1, c.zy, col);
          apply_yu<E, false>(Ks, C1, c.zu, C1, c.zu, col);
          if (valid) {
#pragma unroll
            for (int i = 0; i < E::NY; ++i)
              if (c.gy[i] > 0) atomicAdd(out + ((int64_t)m * D.nfree_y + (c.gy[i] - 1)), col[i]);
          }
        }
      }
    } else if (OP == OP_JTPROD) {
      double w_e[E::NY];
      gather_free<E::NY>(v, c.gy, w_e);
      double kacc[NPS];
#pragma unroll
      for (int m = 0; m < NPS; ++m) kacc[m] = 0.0;
      apply_T<E, true, (NP > 0)>(Ks, A.D, A.D1, w_e, c.zy, c.zu, ey, eu, kacc);
      if (NP > 0) {
        double a1[NPS][E::NY];
#pragma unroll
        for (int m = 0; m < NPS; ++m)
#pragma unroll
          for (int i = 0; i < E::NY; ++i) a1[m][i] = 0.0;
        accum_A<E, 1>(tab, D, x, cell, valid, a1);
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          double a = kacc[m];
#pragma unroll
          for (int i = 0; i < E::NY; ++i) a += zmul(a1[m][i], w_e[i]);
          if (valid) ksum[m] += a;
        }
      }
    } else if (OP == OP_HPROD || OP == OP_HESS_KAPPA) {
      double vy[E::NY], vu[E::NUS];
      if (OP == OP_HPROD) {
        gather_free<E::NY>(v + D.nparam, c.gy, vy);
        if (E::NU > 0) gather_free<E::NUS>(v + D.nparam + D.nfree_y, c.gu, vu);
      }
      // ---- constraint part: (sum_m v_km G1_m' lambda ; lambda.(G1_m v_z) + sum_n v_kn lambda.a2_mn)
      if (lambda != nullptr && NP > 0) {
        double lam_e[E::NY];
        gather_free<E::NY>(lambda, c.gy, lam_e);
        double a2[AOrder<E, 2>::NC][E::NY];
#pragma unroll
        for (int cc = 0; cc < AOrder<E, 2>::NC; ++cc)
#pragma unroll
          for (int i = 0; i < E::NY; ++i) a2[cc][i] = 0.0;
        accum_A<E, 2>(tab, D, x, cell, valid, a2);
        double la2[AOrder<E, 2>::NC];  // lambda . a2_mn  (m >= n, n outer)
#pragma unroll
        for (int cc = 0; cc < AOrder<E, 2>::NC; ++cc) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < E::NY; ++i) a += zmul(a2[cc][i], lam_e[i]);
          la2[cc] = valid ? a : 0.0;
        }
        if (OP == OP_HPROD) {
          double Cv[NB][NK], kacc[NPS];
#pragma unroll
          for (int t = 0; t < NB; ++t)
#pragma unroll
            for (int k = 0; k < NK; ++k) {
              double a = 0.0;
#pragma unroll
              for (int m = 0; m < NP; ++m) a += zmul(A.D1[t][k][m], vk[m]);
              Cv[t][k] = a;
            }
#pragma unroll
          for (int m = 0; m < NPS; ++m) kacc[m] = 0.0;
          apply_T<E, true, true>(Ks, Cv, A.D1, lam_e, vy, vu, ey, eu, kacc);
          int cc = 0;
#pragma unroll
          for (int n = 0; n < NP; ++n)
#pragma unroll
            for (int m = n; m < NP; ++m) {
              kacc[m] += zmul(la2[cc], vk[n]);
              if (m != n) kacc[n] += zmul(la2[cc], vk[m]);
              ++cc;
            }
#pragma unroll
          for (int m = 0; m < NP; ++m)
            if (valid) ksum[m] += kacc[m];
        } else {
          // HESS_KAPPA (constraint trapezoid): cross rows (dof, kappa_m) = (G1_m' lambda); corner = lambda.a2
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            double C0[NB][NK], ry[E::NY], ru[E::NUS], dummy[NPS];
#pragma unroll
            for (int t = 0; t < NB; ++t)
#pragma unroll
              for (int k = 0; k < NK; ++k) C0[t][k] = A.D1[t][k][m];
#pragma unroll
            for (int a = 0; a < E::NY; ++a) ry[a] = 0.0;
#pragma unroll
            for (int b = 0; b < E::NUS; ++b) ru[b] = 0.0;
            apply_T<E, true, false>(Ks, C0, A.D1, lam_e, vy, vu, ry, ru, dummy);
            if (valid) {
#pragma unroll
              for (int a = 0; a < E::NY; ++a)
                if (c.gy[a] > 0) atomicAdd(out + trap_pos((int64_t)NP + c.gy[a], m + 1, prows), ry[a]);
#pragma unroll
              for (int b = 0; b < E::NU; ++b)
                if (c.gu[b] > 0) atomicAdd(out + trap_pos((int64_t)NP + D.nfree_y + c.gu[b], m + 1, prows), ru[b]);
            }
          }
#pragma unroll
          for (int cc = 0; cc < AOrder<E, 2>::NC; ++cc) ksum[cc] += la2[cc];
        }
      }
      // ---- objective part of hprod!: the Hessian C of f w.r.t. the pointwise fields is point- and state-independent
      if (OP == OP_HPROD && !jzero(obj_weight)) {
        using J2 = Jet2<E::M>;
        J2 s[E::M];
#pragma unroll
        for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + NP) ? __ldg(x + (k - E::SK)) : 0.0, k);
        const PointGeo<E> pc0 = point_ctx<E>(D, tab, cell, valid, 0, I::NEEDS_COEF_DERIV);
        const J2 F = I::f(pc0, s);
        double C[NB][NK];
#pragma unroll
        for (int t = 0; t < NB; ++t)
#pragma unroll
          for (int k = 0; k < NK; ++k) C[t][k] = zmul(k < NB ? F.h[J2::idx(t, k)] : 0.0, obj_weight);
        double hy[E::NY], hu[E::NUS];
#pragma unroll
        for (int a = 0; a < E::NY; ++a) hy[a] = 0.0;
#pragma unroll
        for (int b = 0; b < E::NUS; ++b) hu[b] = 0.0;
        apply_yy<E, false>(Ks, C, vy, C, vy, hy);   // block (y,y): hy[a] += sum_b yy(k,l,b,a) C_kl v_b
        if (E::NU > 0) {
          using KT = KTab<E>;
#pragma unroll
          for (int l = 0; l < NB; ++l) {
            const double cul = zmul(F.h[J2::idx(E::SU, l)], obj_weight);
            if (!jzero(cul)) {
#pragma unroll
              for (int j = 0; j < E::NY; ++j) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < E::NU; ++i) {
                  hu[i] = fma(Ks[KT::uy(l, j, i)] * cul, vy[j], hu[i]);
                  a = fma(Ks[KT::uy(l, j, i)], vu[i], a);
                }
                hy[j] = fma(cul, a, hy[j]);
              }
            }
          }
          const double cuu = zmul(F.h[J2::idx(E::SU, E::SU)], obj_weight);
          if (!jzero(cuu)) {
#pragma unroll
            for (int j = 0; j < E::NU; ++j) {
              const double wv = cuu * vu[j];
#pragma unroll
              for (int i = 0; i < E::NU; ++i) hu[i] = fma(Ks[KT::uu(j, i)], wv, hu[i]);
            }
          }
        }
        if (NP > 0) {  // kappa rows/columns: first-order tensors y1, u1 and the cell volume
          using KT = KTab<E>;
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            double a = 0.0;
#pragma unroll
            for (int n = 0; n < NP; ++n) a += zmul(F.h[J2::idx(E::SK + m, E::SK + n)], Ks[KT::OFF_VOL] * vk[n]);
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              const double ck = F.h[J2::idx(E::SK + m, k)];
              if (!jzero(ck)) {
#pragma unroll
                for (int j = 0; j < E::NY; ++j) {
                  a = fma(ck * Ks[KT::y1(k, j)], vy[j], a);
                  hy[j] = fma(ck * Ks[KT::y1(k, j)] * obj_weight, vk[m], hy[j]);
                }
              }
            }
            if (E::NU > 0) {
              const double cu = F.h[J2::idx(E::SK + m, E::SU)];
              if (!jzero(cu)) {
#pragma unroll
                for (int b = 0; b < E::NU; ++b) {
                  a = fma(cu * Ks[KT::u1(b)], vu[b], a);
                  hu[b] = fma(cu * Ks[KT::u1(b)] * obj_weight, vk[m], hu[b]);
                }
              }
            }
            if (valid) ksum[m] = fma(obj_weight, a, ksum[m]);
          }
        }
#pragma unroll
        for (int a = 0; a < E::NY; ++a) ey[a] += hy[a];
#pragma unroll
        for (int b = 0; b < E::NUS; ++b) eu[b] += hu[b];
      }
    }
    if (valid) {
      if (OP == OP_JTPROD || OP == OP_HPROD) scatter_element<E>(D, c, ey, eu, out);
      if (OP == OP_CONS || OP == OP_JPROD) {
#pragma unroll
        for (int i = 0; i < E::NY; ++i)
          if (c.gy[i] > 0) atomicAdd(out + (c.gy[i] - 1), ey[i]);
      }
    }
  }
  __shared__ double red[kWarpsPerCta];
  constexpr int NPART = kappa_partials<E>(OP);
  if (NPART > 0) {
#pragma unroll
    for (int k = 0; k < NPART; ++k) cta_partial(ksum[k], partials + gridDim.x, NPART, k, red);
  }
}

}  // namespace pdenlp
