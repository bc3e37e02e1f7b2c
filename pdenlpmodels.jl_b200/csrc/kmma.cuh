// kmma.cuh — FP64 tensor-core ("DMMA") versions of the dof-indexed callbacks and of the dense kappa blocks.
//
// For an integrand whose residual is affine in the FE fields with CELL-INDEPENDENT linear part (traits
// FE_DERIVS_POINT_INDEPENDENT && RES_FE_HESSIAN_ZERO && FE_DERIVS_CELL_INDEPENDENT) every callback is a handful of
// small dense mat-vecs per cell with the SAME element matrices in every cell (kvector.cuh derives them):
//     c = a(kappa) + G z,   J v,   J' w,   H v,   kappa blocks.
// With lane = cell (kvector.cuh) each FMA of those mat-vecs needs its own uniform operand fetch and the kernels are
// instruction-issue bound (config 5 jprod!: 1500 instructions per 32 cells, 0.15-0.30 ms against 0.03 ms of HBM time).
// Here the cells are the M dimension of a GEMM:  out[cell][o] = sum_in vec[cell][in] * M[in][o]  runs on the FP64
// tensor cores as mma.sync.m8n8k4 (SASS DMMA.8x8x4; FP64 has no tcgen05 form — this is Blackwell's FP64 tensor path):
//   A fragment (8 cells x 4 inputs)  : lane holds vec[cell = 8g + lane/4][in = 4ks + lane%4]   — gathered straight from
//                                      the dof vectors in this layout (no shuffles, no transposition)
//   B fragment (4 inputs x 8 outputs): lane holds M[in = 4ks + lane%4][o = 8nt + lane/4]        — loop-invariant registers,
//                                      built once per kernel from the reference tensors in the constant bank
//   C fragment (8 cells x 8 outputs) : lane holds out[cell = 8g + lane/4][o = 8nt + 2(lane%4) + {0,1}] — scattered with
//                                      red.global.add.f64 through dof ids loaded in the same layout.
// A warp still owns a tile of 32 consecutive cells (4 groups of 8).  The quadrature-point sums of the a-type vectors
// (kvector.cuh) are GEMMs as well: A[cell][q] = the integrand's value / kappa-derivative at (cell, q) evaluated by the
// lane, B_t[q][i] = w_q T_i[t](q).  Where only  w . a  is wanted (kappa entries) the multiplier is first interpolated to
// the quadrature points (Lam_t = B_t' w, one more GEMM) and contracted there.
// Structural zeros: matrices / a-terms that vanish are skipped through warp-uniform flags computed once per kernel
// (votes over the lanes' fragment entries; a-terms: probed with 32 different pseudo-random coefficient values).
// Like every trait-selected path these kernels are compared with the generic per-point kernels by pdenlp_create.
#pragma once
#include "kvector.cuh"

namespace pdenlp {

template <class E>
__host__ __device__ constexpr bool kmma_enabled() {
#ifdef PDENLP_NO_KMMA
  return false;
#else
  return kvec_enabled<E>() && E::Integrand::FE_DERIVS_CELL_INDEPENDENT && E::NY <= 16 && E::NU <= 16 && E::NQ <= 16;
#endif
}

// cell groups (of 8 cells = one DMMA row block) a warp processes per iteration: fewer groups = fewer live registers
// (more resident warps to hide the gather latency), more groups = more independent loads and DMMA chains per warp
#ifndef PDENLP_KMMA_G
#define PDENLP_KMMA_G 2
#endif
constexpr int KG = PDENLP_KMMA_G;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// B fragments of one NIN x NOUT matrix (entries M[in][o]); `any` is warp-uniform
template <int NIN, int NOUT>
struct MFrag {
  static constexpr int KS = (NIN + 3) / 4, NT = (NOUT + 7) / 8;
  double b[KS > 0 ? KS : 1][NT > 0 ? NT : 1];
  bool any;
};
template <int NIN, int NOUT, class F>
__device__ __forceinline__ MFrag<NIN, NOUT> make_frag(int lane, F&& entry /* (o, in) -> M[in][o] */) {
  MFrag<NIN, NOUT> m;
  bool nz = false;
#pragma unroll
  for (int ks = 0; ks < MFrag<NIN, NOUT>::KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < MFrag<NIN, NOUT>::NT; ++nt) {
      const int in = 4 * ks + (lane & 3), o = 8 * nt + (lane >> 2);
      const bool inside = in < NIN && o < NOUT;
      const double e = entry(inside ? o : 0, inside ? in : 0);
      m.b[ks][nt] = inside ? e : 0.0;
      nz = nz || (inside && !jzero(e));
    }
  m.any = __any_sync(0xffffffffu, nz);
  return m;
}
// c[g][nt][.] += a[g][ks] * M   (all 32 lanes must call; the caller tests m.any, which is warp-uniform)
template <int NIN, int NOUT>
__device__ __forceinline__ void mv(const MFrag<NIN, NOUT>& m, const double (&a)[KG][MFrag<NIN, NOUT>::KS],
                                   double (&c)[KG][MFrag<NIN, NOUT>::NT][2]) {
#pragma unroll
  for (int g = 0; g < KG; ++g)
#pragma unroll
    for (int nt = 0; nt < MFrag<NIN, NOUT>::NT; ++nt)
#pragma unroll
      for (int ks = 0; ks < MFrag<NIN, NOUT>::KS; ++ks) dmma884(c[g][nt][0], c[g][nt][1], a[g][ks], m.b[ks][nt]);
}

// dof ids of a tile in the A-fragment ("in") and C-fragment ("out") layouts; 0 = no dof (padding / lane without a cell)
template <int N>
__device__ __forceinline__ void ids_in(const int32_t* __restrict__ dofs, int64_t ldc, int64_t cell0, int64_t ncells, int lane,
                                       int (&id)[KG][(N + 3) / 4]) {
#pragma unroll
  for (int g = 0; g < KG; ++g)
#pragma unroll
    for (int ks = 0; ks < (N + 3) / 4; ++ks) {
      const int in = 4 * ks + (lane & 3);
      const int64_t cell = cell0 + 8 * g + (lane >> 2);
      id[g][ks] = (in < N && cell < ncells) ? __ldg(dofs + (int64_t)in * ldc + cell) : 0;
    }
}
template <int N>
__device__ __forceinline__ void ids_out(const int32_t* __restrict__ dofs, int64_t ldc, int64_t cell0, int64_t ncells, int lane,
                                        int (&id)[KG][(N + 7) / 8][2]) {
#pragma unroll
  for (int g = 0; g < KG; ++g)
#pragma unroll
    for (int nt = 0; nt < (N + 7) / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int o = 8 * nt + 2 * (lane & 3) + e;
        const int64_t cell = cell0 + 8 * g + (lane >> 2);
        id[g][nt][e] = (o < N && cell < ncells) ? __ldg(dofs + (int64_t)o * ldc + cell) : 0;
      }
}
// values of a dof vector through ids (PosNegReindex when `dir` is given, zero Dirichlet values otherwise).
// Branch-free: the address is selected, the load is unconditional (entry 0 for "no dof"), the value is selected —
// so the loads of a whole fragment (and of several fragments) are issued back to back.
template <int A, int B>
__device__ __forceinline__ void gather2(const double* __restrict__ vec, const double* __restrict__ dir, const int (&id)[A][B], double (&val)[A][B]) {
  // (`dir` is nullptr as a literal at the call sites without Dirichlet values: the test folds at compile time)
  double raw[A][B];
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int g = id[i][j];
      const double* p = dir != nullptr ? (g > 0 ? vec + (g - 1) : dir + (g < 0 ? -g - 1 : 0)) : vec + (g > 0 ? g - 1 : 0);
      raw[i][j] = __ldg(p);
    }
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int j = 0; j < B; ++j) val[i][j] = (dir != nullptr ? id[i][j] != 0 : id[i][j] > 0) ? raw[i][j] : 0.0;
}
// the same with Dirichlet values, for call sites where the Dirichlet array is a run-time pointer that is never null
template <int A, int B>
__device__ __forceinline__ void gather2d(const double* __restrict__ vec, const double* __restrict__ dir, const int (&id)[A][B], double (&val)[A][B]) {
  double raw[A][B];
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int g = id[i][j];
      raw[i][j] = __ldg(g > 0 ? vec + (g - 1) : dir + (g < 0 ? -g - 1 : 0));
    }
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int j = 0; j < B; ++j) val[i][j] = id[i][j] != 0 ? raw[i][j] : 0.0;
}
template <int A, int B>
__device__ __forceinline__ void gather3(const double* __restrict__ vec, const double* __restrict__ dir, const int (&id)[KG][A][B], double (&val)[KG][A][B]) {
#pragma unroll
  for (int g = 0; g < KG; ++g) gather2<A, B>(vec, dir, id[g], val[g]);
}
template <int A, int B>
__device__ __forceinline__ void gather3d(const double* __restrict__ vec, const double* __restrict__ dir, const int (&id)[KG][A][B], double (&val)[KG][A][B]) {
#pragma unroll
  for (int g = 0; g < KG; ++g) gather2d<A, B>(vec, dir, id[g], val[g]);
}
template <int NT>
__device__ __forceinline__ void scatter_out(double* __restrict__ out, const int (&id)[KG][NT][2], const double (&c)[KG][NT][2]) {
#pragma unroll
  for (int g = 0; g < KG; ++g)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (id[g][nt][e] > 0) atomicAdd(out + (id[g][nt][e] - 1), c[g][nt][e]);
}
template <int NT>
__device__ __forceinline__ void zero_out(double (&c)[KG][NT][2]) {
#pragma unroll
  for (int g = 0; g < KG; ++g)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) c[g][nt][0] = c[g][nt][1] = 0.0;
}
template <int NT>
__device__ __forceinline__ double dot_out(const double (&a)[KG][NT][2], const double (&b)[KG][NT][2]) {
  double s = 0.0;
#pragma unroll
  for (int g = 0; g < KG; ++g)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s = fma(a[g][nt][0], b[g][nt][0], fma(a[g][nt][1], b[g][nt][1], s));
  return s;
}

// The kappa-derivative components of res(s = 0; coefficients at one quadrature point) an op needs:
//   CONS: the value;  JPROD: sum_m vk_m d/dk_m;  JAC_KAPPA, JTPROD: d/dk_m (component m);
//   HPROD: sum_n vk_n d2/dk_m dk_n (component m);  HESS_KAPPA: d2/dk_m dk_n, m >= n (n outer)
template <class E, int OP>
struct XComp {
  static constexpr int NP = E::NP;
  static constexpr int NC = (OP == OP_CONS || OP == OP_JPROD) ? 1
                          : (OP == OP_HESS_KAPPA ? (NP > 0 ? NP * (NP + 1) / 2 : 1) : (NP > 0 ? NP : 1));
};
template <class E, int OP>
__device__ __forceinline__ void eval_X(const Tables<E>& tab, double target, double source, const double* kap, const double* vk,
                                       double (&X)[E::NB][XComp<E, OP>::NC]) {
  using I = typename E::Integrand;
  constexpr int NP = E::NP;
  PointCtx pc;
  pc.par = tab.par; pc.target = target; pc.source = source;
#pragma unroll
  for (int t = 0; t < E::NB; ++t)
#pragma unroll
    for (int c = 0; c < XComp<E, OP>::NC; ++c) X[t][c] = 0.0;
  if (OP == OP_CONS) {
    double s[E::M], R[E::NR];
#pragma unroll
    for (int k = 0; k < E::M; ++k) s[k] = (k >= E::SK && k < E::SK + NP) ? kap[k - E::SK] : 0.0;
    I::res(pc, s, R);
#pragma unroll
    for (int t = 0; t < E::NB; ++t) X[t][0] = R[t];
  } else if (OP == OP_JPROD || OP == OP_JAC_KAPPA || OP == OP_JTPROD) {
    if (NP > 0) {
      Jet1<E::M> s[E::M], R[E::NR];
#pragma unroll
      for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + NP) ? kap[k - E::SK] : 0.0, k);
      I::res(pc, s, R);
#pragma unroll
      for (int t = 0; t < E::NB; ++t)
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          if (OP == OP_JPROD) X[t][0] += zmul(R[t].g[E::SK + m], vk[m]);
          else X[t][m] = R[t].g[E::SK + m];
        }
    }
  } else if (NP > 0) {  // HPROD, HESS_KAPPA
    using J2 = Jet2<E::M>;
    J2 s[E::M], R[E::NR];
#pragma unroll
    for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + NP) ? kap[k - E::SK] : 0.0, k);
    I::res(pc, s, R);
#pragma unroll
    for (int t = 0; t < E::NB; ++t) {
      int c = 0;
#pragma unroll
      for (int n = 0; n < NP; ++n)
#pragma unroll
        for (int m = n; m < NP; ++m) {
          const double h = R[t].h[J2::idx(E::SK + m, E::SK + n)];
          if (OP == OP_HESS_KAPPA) { X[t][c] = h; ++c; }
          else {
            X[t][m] += zmul(h, vk[n]);
            if (m != n) X[t][n] += zmul(h, vk[m]);
          }
        }
    }
  }
}

#ifndef PDENLP_KMMA_MINB
#define PDENLP_KMMA_MINB 3
#endif
template <class E, int OP>
__global__ void __launch_bounds__(kThreads, PDENLP_KMMA_MINB)
k_vector_M(const __grid_constant__ Tables<E> tab, const __grid_constant__ KConst<E> kc, const DevModel D, const double* __restrict__ x,
           const double* __restrict__ lambda, const double* __restrict__ v, double obj_weight, int64_t prows,
           double* __restrict__ out, double* __restrict__ partials) {
  using I = typename E::Integrand;
  using AR = AffineRes<E>;
  using KT = KTab<E>;
  constexpr int NB = E::NB, NK = NB + 1, NPS = AR::NPS, NP = E::NP, NY = E::NY, NU = E::NU, NQ = E::NQ;
  constexpr int NUS = NU > 0 ? NU : 1;
  constexpr int KSY = (NY + 3) / 4, KSU = (NUS + 3) / 4, NTY = (NY + 7) / 8, NTU = (NUS + 7) / 8;
  constexpr int KSQ = (NQ + 3) / 4, NTQ = (NQ + 7) / 8;
  constexpr int NC = XComp<E, OP>::NC;
  // The fragment entries are read with LANE-DEPENDENT indices: from global memory (L1/L2 hits after the first CTA), not
  // from the constant bank — a constant load whose address diverges across the warp is replayed once per distinct
  // address (measured: the prologue took ~25 us per CTA with c[][] operands).  Uniformly indexed entries use kc.
  const double* const Ks = D.ktab;
  const double* const Kc = kc.v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int NKK = NP > 0 ? NP * (NP + 1) / 2 : 1;
  double ksum[NKK];
#pragma unroll
  for (int k = 0; k < NKK; ++k) ksum[k] = 0.0;

  // ---------------- once per kernel: the cell-independent coefficient tables and the B fragments
  double kap[NPS], vk[NPS];
#pragma unroll
  for (int m = 0; m < NPS; ++m) {
    kap[m] = NP > 0 ? __ldg(x + m) : 0.0;
    vk[m] = (NP > 0 && (OP == OP_JPROD || OP == OP_HPROD)) ? __ldg(v + m) : 0.0;
  }
  AR A;
  {
    PointCtx pc0;
    pc0.par = tab.par; pc0.target = 0.0; pc0.source = 0.0;  // (D does not depend on the coefficient fields: trait)
    eval_affine<E, (NP > 0)>(pc0, x, A);
  }
  double Cv[NB][NK];  // sum_m vk_m dD/dk_m
#pragma unroll
  for (int t = 0; t < NB; ++t)
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      double a = 0.0;
#pragma unroll
      for (int m = 0; m < NP; ++m) a += zmul(A.D1[t][k][m], vk[m]);
      Cv[t][k] = a;
    }
  // element-matrix entries from a coefficient table c[t][k]:  Y: rows = test dof i, cols = state dof j;  U: cols = control dof b
  auto MY = [&](const double (&c)[NB][NK], int i, int j) {
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (!jzero(c[t][k])) a = fma(c[t][k], __ldg(Ks + KT::yy(t, k, 0, 0) + j * NY + i), a);
    return a;
  };
  auto MU = [&](const double (&c)[NB][NK], int i, int b) {
    double a = 0.0;
    if (NU > 0) {
#pragma unroll
      for (int t = 0; t < NB; ++t)
        if (!jzero(c[t][NB])) a = fma(c[t][NB], __ldg(Ks + KT::yu(t, 0, 0) + b * NY + i), a);
    }
    return a;
  };
  auto table_m = [&](int m, double (&c)[NB][NK]) {
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int k = 0; k < NK; ++k) c[t][k] = A.D1[t][k][m < NPS ? m : 0];
  };
  // B_t[q][i] = w_q T_i[t](q)  (vector mode: in = q, out = i)  /  its transpose (dot mode: in = i, out = q)
  auto BT = [&](int t, int q, int i) { return __ldg(Ks + KT::bt(t, 0, 0) + q * NY + i); };

  // structural flags of the a-term components: probed with lane-dependent pseudo-random coefficient values
  bool xflag[NB][NC];
  {
    double X[NB][NC];
    eval_X<E, OP>(tab, 0.3711 + 0.1173 * lane, 0.7093 - 0.0719 * lane, kap, vk, X);
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int c = 0; c < NC; ++c) xflag[t][c] = __any_sync(0xffffffffu, !jzero(X[t][c]));
  }
  bool xany[NB];  // some component of slot t is present
#pragma unroll
  for (int t = 0; t < NB; ++t) {
    xany[t] = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) xany[t] = xany[t] || xflag[t][c];
  }

  // ---- all B fragments of the op (loop-invariant registers).  in -> out:  yy: y -> y,  uy: u -> y,  yu: y -> u,  uu: u -> u
  constexpr bool VECMODE = OP == OP_CONS || OP == OP_JPROD || OP == OP_JAC_KAPPA;          // a-vectors enter the output
  constexpr bool DOTMODE = NP > 0 && (OP == OP_JTPROD || OP == OP_HPROD || OP == OP_HESS_KAPPA);  // only w . a is wanted
  constexpr bool PER_M = NP > 0 && (OP == OP_JAC_KAPPA || OP == OP_JTPROD || OP == OP_HPROD || OP == OP_HESS_KAPPA);
  constexpr bool PER_M_T = OP == OP_JTPROD || OP == OP_HESS_KAPPA;  // per-kappa matrices applied transposed (test -> trial)
  MFrag<NY, NY> yy0, yy1, yym[NPS];
  MFrag<NUS, NY> uy0, uy1, uym[NPS];
  MFrag<NY, NUS> yu0, yu1, yum[NPS];
  MFrag<NUS, NUS> uu0;
  MFrag<NQ, NY> bt[NB];    // vector mode: in = q, out = test dof
  MFrag<NY, NQ> btT[NB];   // dot mode: in = test dof, out = q
  yy0.any = yy1.any = uy0.any = uy1.any = yu0.any = yu1.any = uu0.any = false;
#pragma unroll
  for (int m = 0; m < NPS; ++m) yym[m].any = uym[m].any = yum[m].any = false;
#pragma unroll
  for (int t = 0; t < NB; ++t) bt[t].any = btT[t].any = false;
  if (OP == OP_CONS || OP == OP_JPROD) {
    yy0 = make_frag<NY, NY>(lane, [&](int i, int j) { return MY(A.D, i, j); });
    if (NU > 0) uy0 = make_frag<NUS, NY>(lane, [&](int i, int b) { return MU(A.D, i, b); });
  }
  if (OP == OP_JPROD && NP > 0) {
    yy1 = make_frag<NY, NY>(lane, [&](int i, int j) { return MY(Cv, i, j); });
    if (NU > 0) uy1 = make_frag<NUS, NY>(lane, [&](int i, int b) { return MU(Cv, i, b); });
  }
  if (OP == OP_JTPROD) {
    yy0 = make_frag<NY, NY>(lane, [&](int j, int i) { return MY(A.D, i, j); });
    if (NU > 0) yu0 = make_frag<NY, NUS>(lane, [&](int b, int i) { return MU(A.D, i, b); });
  }
  if (OP == OP_HPROD && NP > 0) {
    yy0 = make_frag<NY, NY>(lane, [&](int j, int i) { return MY(Cv, i, j); });
    if (NU > 0) yu0 = make_frag<NY, NUS>(lane, [&](int b, int i) { return MU(Cv, i, b); });
  }
  if (PER_M) {
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      double c1[NB][NK];
      table_m(m, c1);
      if (PER_M_T) {
        yym[m] = make_frag<NY, NY>(lane, [&](int j, int i) { return MY(c1, i, j); });
        if (NU > 0) yum[m] = make_frag<NY, NUS>(lane, [&](int b, int i) { return MU(c1, i, b); });
      } else {
        yym[m] = make_frag<NY, NY>(lane, [&](int i, int j) { return MY(c1, i, j); });
        if (NU > 0) uym[m] = make_frag<NUS, NY>(lane, [&](int i, int b) { return MU(c1, i, b); });
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NB; ++t) {
    if (VECMODE) bt[t] = make_frag<NQ, NY>(lane, [&](int i, int q) { return BT(t, q, i); });
    if (DOTMODE) btT[t] = make_frag<NY, NQ>(lane, [&](int q, int i) { return BT(t, q, i); });
  }
  // objective part of hprod!: C = Hessian of f w.r.t. the pointwise fields (cell- and state-independent), times obj_weight
  double kk_obj[NPS], cy_in[NPS][KSY], cu_in[NPS][KSU], cy_out[NPS][NTY][2], cu_out[NPS][NTU][2];
  bool cross_obj[NPS];
#pragma unroll
  for (int m = 0; m < NPS; ++m) { kk_obj[m] = 0.0; cross_obj[m] = false; }
  if (OP == OP_HPROD && !jzero(obj_weight)) {
    using J2 = Jet2<E::M>;
    J2 s[E::M];
#pragma unroll
    for (int k = 0; k < E::M; ++k) jet_seed(s[k], (k >= E::SK && k < E::SK + NP) ? kap[k - E::SK] : 0.0, k);
    PointCtx pc0;
    pc0.par = tab.par; pc0.target = 0.0; pc0.source = 0.0;
    const J2 F = I::f(pc0, s);
    auto Ch = [&](int k, int l) { return zmul(F.h[J2::idx(k, l)], obj_weight); };
    yy1 = make_frag<NY, NY>(lane, [&](int a, int b) {
      double r = 0.0;
#pragma unroll
      for (int k = 0; k < NB; ++k)
#pragma unroll
        for (int l = 0; l < NB; ++l)
          if (!jzero(Ch(k, l))) r = fma(Ch(k, l), __ldg(Ks + KT::yy(k, l, 0, 0) + b * NY + a), r);
      return r;
    });
    if (NU > 0) {
      auto huy_e = [&](int iu_, int jy_) {  // rows u_i, cols y_j
        double r = 0.0;
#pragma unroll
        for (int l = 0; l < NB; ++l)
          if (!jzero(Ch(E::SU, l))) r = fma(Ch(E::SU, l), __ldg(Ks + KT::uy(l, 0, 0) + jy_ * NU + iu_), r);
        return r;
      };
      uy0 = make_frag<NUS, NY>(lane, [&](int a, int i) { return huy_e(i, a); });  // out y_a, in u_i
      yu1 = make_frag<NY, NUS>(lane, [&](int i, int a) { return huy_e(i, a); });  // out u_i, in y_a
      uu0 = make_frag<NUS, NUS>(lane, [&](int i, int j) { return jzero(Ch(E::SU, E::SU)) ? 0.0 : Ch(E::SU, E::SU) * __ldg(Ks + KT::uu(0, 0) + j * NU + i); });
    }
    // kappa rows / columns: the kappa-kappa term (times the cell volume) and the cross terms through the first-order
    // tensors  c_y[j] = sum_k C[km][k] y1(k, j),  c_u[b] = C[km][SU] u1(b)   (in both fragment layouts)
#pragma unroll
    for (int m = 0; m < NP; ++m) {
#pragma unroll
      for (int n = 0; n < NP; ++n) kk_obj[m] += zmul(Ch(E::SK + m, E::SK + n), Kc[KT::OFF_VOL] * vk[n]);
      bool anyc = false;
#pragma unroll
      for (int k = 0; k <= E::SU; ++k) anyc = anyc || !jzero(Ch(E::SK + m, k));
      cross_obj[m] = anyc;  // (uniform: C does not depend on the lane)
      auto cyf = [&](int j) {
        double cy = 0.0;
#pragma unroll
        for (int k = 0; k < NB; ++k) cy = fma(Ch(E::SK + m, k), __ldg(Ks + KT::y1(k, 0) + (j < NY ? j : 0)), cy);
        return j < NY ? cy : 0.0;
      };
      auto cuf = [&](int b) { return (NU > 0 && b < NU) ? Ch(E::SK + m, E::SU) * __ldg(Ks + KT::u1(0) + (b < NU ? b : 0)) : 0.0; };
#pragma unroll
      for (int ks = 0; ks < KSY; ++ks) cy_in[m][ks] = anyc ? cyf(4 * ks + (lane & 3)) : 0.0;
#pragma unroll
      for (int ks = 0; ks < KSU; ++ks) cu_in[m][ks] = anyc ? cuf(4 * ks + (lane & 3)) : 0.0;
#pragma unroll
      for (int nt = 0; nt < NTY; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) cy_out[m][nt][e] = anyc ? cyf(8 * nt + 2 * (lane & 3) + e) : 0.0;
#pragma unroll
      for (int nt = 0; nt < NTU; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) cu_out[m][nt][e] = anyc ? cuf(8 * nt + 2 * (lane & 3) + e) : 0.0;
    }
  }

  const int64_t vstride = (int64_t)gridDim.x * nwarps;
  const int64_t nblk = (D.ncells + 8 * KG - 1) / (8 * KG);  // blocks of KG groups of 8 consecutive cells
  const double* const xy = x + D.nparam;
  const double* const xu = xy + D.nfree_y;
  // the dof ids of the next block are loaded one iteration ahead (ids -> values is a dependent load chain)
  int iy[KG][KSY], iu[KG][KSU], niy[KG][KSY], niu[KG][KSU];
  {
    const int64_t c0 = ((int64_t)blockIdx.x * nwarps + warp) * (8 * KG);
    ids_in<NY>(D.dofs_y, D.ldc, c0, D.ncells, lane, niy);
    if (NU > 0) ids_in<NUS>(D.dofs_u, D.ldc, c0, D.ncells, lane, niu);
  }
  for (int64_t blk = (int64_t)blockIdx.x * nwarps + warp; blk < nblk; blk += vstride) {
    const int64_t cell0 = blk * (8 * KG);
#pragma unroll
    for (int g = 0; g < KG; ++g) {
#pragma unroll
      for (int ks = 0; ks < KSY; ++ks) iy[g][ks] = niy[g][ks];
#pragma unroll
      for (int ks = 0; ks < KSU; ++ks) iu[g][ks] = niu[g][ks];
    }
    ids_in<NY>(D.dofs_y, D.ldc, cell0 + vstride * (8 * KG), D.ncells, lane, niy);
    if (NU > 0) ids_in<NUS>(D.dofs_u, D.ldc, cell0 + vstride * (8 * KG), D.ncells, lane, niu);
    double ey[KG][NTY][2], eu[KG][NTU][2];
    zero_out<NTY>(ey);
    zero_out<NTU>(eu);

    // a-type vectors in vector mode: acc += sum_t B_t' X_t   (X evaluated at (cell, q) in the A layout).
    // Two phases so that the coefficient loads are issued together with the tile's other gathers.
    auto a_load = [&](int comp, double (&xa)[NB][KG][KSQ]) {
#pragma unroll
      for (int g = 0; g < KG; ++g)
#pragma unroll
        for (int ks = 0; ks < KSQ; ++ks) {
          const int q = 4 * ks + (lane & 3);
          const int64_t cell = cell0 + 8 * g + (lane >> 2);
          const bool ok = q < NQ && cell < D.ncells;
          const int64_t at = ok ? cell * NQ + q : 0;
          const double tg = __ldg(D.coef + at);
          const double sr = __ldg(D.coef + (D.ncells_coef * NQ + at));
          double X[NB][NC];
          eval_X<E, OP>(tab, tg, sr, kap, vk, X);
#pragma unroll
          for (int t = 0; t < NB; ++t) xa[t][g][ks] = ok ? X[t][comp] : 0.0;
        }
    };
    auto a_mma = [&](int comp, const double (&xa)[NB][KG][KSQ], double (&acc)[KG][NTY][2]) {
#pragma unroll
      for (int t = 0; t < NB; ++t)
        if (xflag[t][comp]) mv<NQ, NY>(bt[t], xa[t], acc);
    };
    // a-type vectors in dot mode: this lane's share of  sum_cells w . a_c  for every component c (w in the A layout);
    // the integrand values X are evaluated at (cell, q) in the C layout (load phase), then contracted with B_t' w
    auto d_load = [&](double (&xo)[NB][NC][KG][NTQ][2]) {
#pragma unroll
      for (int g = 0; g < KG; ++g)
#pragma unroll
        for (int nt = 0; nt < NTQ; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int q = 8 * nt + 2 * (lane & 3) + e;
            const int64_t cell = cell0 + 8 * g + (lane >> 2);
            const bool ok = q < NQ && cell < D.ncells;
            const int64_t at = ok ? cell * NQ + q : 0;
            const double tg = __ldg(D.coef + at);
            const double sr = __ldg(D.coef + (D.ncells_coef * NQ + at));
            double X[NB][NC];
            eval_X<E, OP>(tab, tg, sr, kap, vk, X);
#pragma unroll
            for (int t = 0; t < NB; ++t)
#pragma unroll
              for (int c = 0; c < NC; ++c) xo[t][c][g][nt][e] = ok ? X[t][c] : 0.0;
          }
    };
    auto d_mma = [&](const double (&xo)[NB][NC][KG][NTQ][2], const double (&w_in)[KG][KSY], double (&res)[NC]) {
#pragma unroll
      for (int c = 0; c < NC; ++c) res[c] = 0.0;
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        if (!xany[t]) continue;
        double lam_q[KG][NTQ][2];
        zero_out<NTQ>(lam_q);
        mv<NY, NQ>(btT[t], w_in, lam_q);
#pragma unroll
        for (int c = 0; c < NC; ++c) res[c] += dot_out<NTQ>(xo[t][c], lam_q);
      }
    };

    if (OP == OP_CONS) {
      double zy[KG][KSY], zu[KG][KSU], xa[NB][KG][KSQ];
      // loads first (one memory round trip), then the DMMAs
      if (yy0.any) gather2d<KG, KSY>(xy, D.dir_y, iy, zy);
      if (NU > 0 && uy0.any) gather2d<KG, KSU>(xu, D.dir_u, iu, zu);
      a_load(0, xa);
      a_mma(0, xa, ey);
      if (yy0.any) mv<NY, NY>(yy0, zy, ey);
      if (NU > 0 && uy0.any) mv<NUS, NY>(uy0, zu, ey);
    } else if (OP == OP_JPROD) {
      double a0[KG][KSY], b0[KG][KSU], a1[KG][KSY], b1[KG][KSU], xa[NB][KG][KSQ];
      if (yy0.any) gather2<KG, KSY>(v + D.nparam, nullptr, iy, a0);
      if (NU > 0 && uy0.any) gather2<KG, KSU>(v + D.nparam + D.nfree_y, nullptr, iu, b0);
      if (NP > 0) {
        if (yy1.any) gather2d<KG, KSY>(xy, D.dir_y, iy, a1);
        if (NU > 0 && uy1.any) gather2d<KG, KSU>(xu, D.dir_u, iu, b1);
        a_load(0, xa);
      }
      if (yy0.any) mv<NY, NY>(yy0, a0, ey);
      if (NU > 0 && uy0.any) mv<NUS, NY>(uy0, b0, ey);
      if (NP > 0) {
        if (yy1.any) mv<NY, NY>(yy1, a1, ey);
        if (NU > 0 && uy1.any) mv<NUS, NY>(uy1, b1, ey);
        a_mma(0, xa, ey);
      }
    } else if (OP == OP_JAC_KAPPA) {
      if (NP > 0) {
        double zy[KG][KSY], zu[KG][KSU];
        gather2d<KG, KSY>(xy, D.dir_y, iy, zy);
        if (NU > 0) gather2d<KG, KSU>(xu, D.dir_u, iu, zu);
        int oy[KG][NTY][2];
        ids_out<NY>(D.dofs_y, D.ldc, cell0, D.ncells, lane, oy);
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          double col[KG][NTY][2], xa[NB][KG][KSQ];
          zero_out<NTY>(col);
          bool anyx = false;
#pragma unroll
          for (int t = 0; t < NB; ++t) anyx = anyx || xflag[t][m];
          if (anyx) { a_load(m, xa); a_mma(m, xa, col); }
          if (yym[m].any) mv<NY, NY>(yym[m], zy, col);
          if (NU > 0 && uym[m].any) mv<NUS, NY>(uym[m], zu, col);
          scatter_out<NTY>(out + (int64_t)m * D.nfree_y, oy, col);
        }
      }
    } else if (OP == OP_JTPROD) {
      // kappa_m entry: w . (a1_m + G1_m z)  =  w . a1_m  +  z_y . (G1y_m' w)  +  z_u . (G1u_m' w)
      double w_in[KG][KSY];
      gather2<KG, KSY>(v, nullptr, iy, w_in);
      bool need_y = false, need_u = false;
#pragma unroll
      for (int m = 0; m < NP; ++m) { need_y = need_y || yym[m].any; need_u = need_u || (NU > 0 && yum[m].any); }
      int oy[KG][NTY][2], ou[KG][NTU][2];
      double zyo[KG][NTY][2], zuo[KG][NTU][2], xo[NB][NC][KG][NTQ][2];
      ids_out<NY>(D.dofs_y, D.ldc, cell0, D.ncells, lane, oy);
      if (NU > 0) ids_out<NUS>(D.dofs_u, D.ldc, cell0, D.ncells, lane, ou);
      if (NP > 0) {
        d_load(xo);
        if (need_y) gather3d<NTY, 2>(xy, D.dir_y, oy, zyo);
        if (need_u) gather3d<NTU, 2>(xu, D.dir_u, ou, zuo);
      }
      if (yy0.any) mv<NY, NY>(yy0, w_in, ey);
      if (NU > 0 && yu0.any) mv<NY, NUS>(yu0, w_in, eu);
      if (NP > 0) {
        double res[NC];
        d_mma(xo, w_in, res);
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          double acc = res[m];
          if (yym[m].any) {
            double p[KG][NTY][2];
            zero_out<NTY>(p);
            mv<NY, NY>(yym[m], w_in, p);
            acc += dot_out<NTY>(p, zyo);
          }
          if (NU > 0 && yum[m].any) {
            double p[KG][NTU][2];
            zero_out<NTU>(p);
            mv<NY, NUS>(yum[m], w_in, p);
            acc += dot_out<NTU>(p, zuo);
          }
          ksum[m] += acc;
        }
      }
      scatter_out<NTY>(out + D.nparam, oy, ey);
      if (NU > 0) scatter_out<NTU>(out + D.nparam + D.nfree_y, ou, eu);
    } else if (OP == OP_HPROD || OP == OP_HESS_KAPPA) {
      double vy[KG][KSY], vu[KG][KSU];
      if (OP == OP_HPROD) {
        gather2<KG, KSY>(v + D.nparam, nullptr, iy, vy);
        if (NU > 0) gather2<KG, KSU>(v + D.nparam + D.nfree_y, nullptr, iu, vu);
      }
      // ---- constraint part
      if (lambda != nullptr && NP > 0) {
        double lam_in[KG][KSY], xo[NB][NC][KG][NTQ][2];
        gather2<KG, KSY>(lambda, nullptr, iy, lam_in);
        d_load(xo);
        double res[NC];
        d_mma(xo, lam_in, res);  // HPROD: component m = sum_n vk_n lambda.a2_mn;  HESS_KAPPA: lambda.a2_mn (m >= n)
        if (OP == OP_HPROD) {
          if (yy0.any) mv<NY, NY>(yy0, lam_in, ey);
          if (NU > 0 && yu0.any) mv<NY, NUS>(yu0, lam_in, eu);
          // kappa_m: lambda . (G1y_m v_y + G1u_m v_u)
          bool need = false;
#pragma unroll
          for (int m = 0; m < NP; ++m) need = need || yym[m].any || (NU > 0 && uym[m].any);
          int oy[KG][NTY][2];
          double lam_o[KG][NTY][2];
          if (need) { ids_out<NY>(D.dofs_y, D.ldc, cell0, D.ncells, lane, oy); gather3<NTY, 2>(lambda, nullptr, oy, lam_o); }
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            double acc = res[m];
            if (yym[m].any || (NU > 0 && uym[m].any)) {
              double p[KG][NTY][2];
              zero_out<NTY>(p);
              if (yym[m].any) mv<NY, NY>(yym[m], vy, p);
              if (NU > 0 && uym[m].any) mv<NUS, NY>(uym[m], vu, p);
              acc += dot_out<NTY>(p, lam_o);
            }
            ksum[m] += acc;
          }
        } else {
          // HESS_KAPPA: cross rows (dof, kappa_m) = G1_m' lambda; the kappa-kappa corner = lambda . a2
          int oy[KG][NTY][2], ou[KG][NTU][2];
          ids_out<NY>(D.dofs_y, D.ldc, cell0, D.ncells, lane, oy);
          if (NU > 0) ids_out<NUS>(D.dofs_u, D.ldc, cell0, D.ncells, lane, ou);
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            if (yym[m].any) {
              double ry[KG][NTY][2];
              zero_out<NTY>(ry);
              mv<NY, NY>(yym[m], lam_in, ry);
#pragma unroll
              for (int g = 0; g < KG; ++g)
#pragma unroll
                for (int nt = 0; nt < NTY; ++nt)
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (oy[g][nt][e] > 0) atomicAdd(out + trap_pos((int64_t)NP + oy[g][nt][e], m + 1, prows), ry[g][nt][e]);
            }
            if (NU > 0 && yum[m].any) {
              double ru[KG][NTU][2];
              zero_out<NTU>(ru);
              mv<NY, NUS>(yum[m], lam_in, ru);
#pragma unroll
              for (int g = 0; g < KG; ++g)
#pragma unroll
                for (int nt = 0; nt < NTU; ++nt)
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (ou[g][nt][e] > 0) atomicAdd(out + trap_pos((int64_t)NP + D.nfree_y + ou[g][nt][e], m + 1, prows), ru[g][nt][e]);
            }
          }
#pragma unroll
          for (int cc = 0; cc < NC; ++cc) ksum[cc] += res[cc];
        }
      }
      // ---- objective part of hprod!
      if (OP == OP_HPROD && !jzero(obj_weight)) {
        if (yy1.any) mv<NY, NY>(yy1, vy, ey);
        if (NU > 0) {
          if (uy0.any) mv<NUS, NY>(uy0, vu, ey);
          if (yu1.any) mv<NY, NUS>(yu1, vy, eu);
          if (uu0.any) mv<NUS, NUS>(uu0, vu, eu);
        }
#pragma unroll
        for (int m = 0; m < NP; ++m) {
#pragma unroll
          for (int g = 0; g < KG; ++g) {
            const bool okc = cell0 + 8 * g + (lane >> 2) < D.ncells;
            if (okc && (lane & 3) == 0) ksum[m] += kk_obj[m];  // one lane per cell adds the kappa-kappa term
            if (cross_obj[m] && okc) {
#pragma unroll
              for (int ks = 0; ks < KSY; ++ks) ksum[m] = fma(cy_in[m][ks], vy[g][ks], ksum[m]);
#pragma unroll
              for (int ks = 0; ks < KSU; ++ks) ksum[m] = fma(cu_in[m][ks], vu[g][ks], ksum[m]);
#pragma unroll
              for (int nt = 0; nt < NTY; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) ey[g][nt][e] = fma(cy_out[m][nt][e], vk[m], ey[g][nt][e]);
#pragma unroll
              for (int nt = 0; nt < NTU; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) eu[g][nt][e] = fma(cu_out[m][nt][e], vk[m], eu[g][nt][e]);
            }
          }
        }
      }
    }

    // ---- scatter (dof-indexed outputs)
    if (OP == OP_CONS || OP == OP_JPROD) {
      int oy[KG][NTY][2];
      ids_out<NY>(D.dofs_y, D.ldc, cell0, D.ncells, lane, oy);
      scatter_out<NTY>(out, oy, ey);
    } else if (OP == OP_HPROD) {
      int oy[KG][NTY][2];
      ids_out<NY>(D.dofs_y, D.ldc, cell0, D.ncells, lane, oy);
      scatter_out<NTY>(out + D.nparam, oy, ey);
      if (NU > 0) {
        int ou[KG][NTU][2];
        ids_out<NUS>(D.dofs_u, D.ldc, cell0, D.ncells, lane, ou);
        scatter_out<NTU>(out + D.nparam + D.nfree_y, ou, eu);
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
