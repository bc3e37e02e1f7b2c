// kernels.cuh — sm_100a FP64 kernels of the derivative-evaluation hot path.
//
// Data layout in HBM (owned by a pdenlp_model):
//   dofs_y / dofs_u : int32, SoA [local dof][ldc]  (ldc = ncells rounded up to 32) so that a
//                     warp (lane = cell) reads 128 B lines; signed, 1-based, <0 = Dirichlet
//   dir_y / dir_u   : FP64 Dirichlet values, index -id-1
//   coef            : FP64 [field][cell][q]
//   base_jy/ju/h    : int64 exclusive scans of the per-tile (32 cells) COO entry counts
//   tables          : passed BY VALUE as a __grid_constant__ kernel parameter => they live in
//                     the constant bank and are consumed as immediate operands of DFMA
//                     (zero load instructions; every lane reads the same entry)
//
// The coordinate kernels (jac_coord!, hess_coord!) are structured streaming writers:
//   lane = cell, warp = tile of 32 consecutive cells = ONE contiguous run of each COO segment.
//   Each lane evaluates its element derivatives in registers (forward-mode jets on the
//   pointwise fields + contraction with the tables), ranks its kept entries with the
//   reference's filter (gid>0, and gidcol<=gidrow for the Hessian — on GLOBAL ids), and
//   writes them at (warp-scan offset + rank) into the warp's shared-memory staging buffer.
//   One elected lane then streams the whole run to HBM with a single TMA bulk store
//   (cp.async.bulk.global.shared::cta, SASS UBLKCP) — fully coalesced 16 B-aligned traffic,
//   the SM's LSUs never touch the 10 GB output stream.
//
// Map of the file:
//   Elem / Tables / KTab / DevModel            element description, constant tables, reference tensors, device model
//   load_ids, gather_state, point_state, ...    per-lane cell data and pointwise fields (single- and multi-field)
//   StageLayout / SegStage                     the staged segment writer (pieces, parity, bulk stores)
//   emit_*_K, load_ktab                        reference-tensor emitters for nq > 1 with point-independent derivatives
//   k_jac_coord, k_hess_coord                  the two streaming writers (register pipeline over tiles; nq > 1 elements
//                                              take the reference-tensor path, the register-accumulate path for small
//                                              nonlinear elements, or per-point accumulation in shared memory)
//   k_copy_chunks, k_jac_cells_cg, k_hess_cells_cg   cell-granular templates: affine problems on identical cells copy the
//                                              image of a template cell (pure 16 B-store stream) and evaluate only
//                                              the boundary cells (lane = entry emission)
//   k_jac_coord_mf, k_hess_coord_mf            the same writers for multi-field spaces (block-ordered emission)
//   k_structure, k_count                       setup: rows / cols and per-tile entry counts
//   k_vector<OP>                               obj, grad!, cons!, jprod!, jtprod!, hprod! (matrix-free, atomics)
//   k_zero_fill, k_sum_partials, k_nofe        zero fill of structurally zero blocks, objective reduction, NoFETerm
//   last_cta, finish_objective / finish_partials   second pass of a reduction by the last CTA to arrive (no extra launch)
//   k_jac_kappa, k_hess_kappa(_kk)             dense kappa blocks
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "integrands.cuh"

namespace pdenlp {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

// ------------------------------------------------------------------ element description
// NLY_/NLU_ are the local dofs of ONE state / control field; a multi-field space (NFY state fields,
// NFU control fields, all fields of a kind sharing one reference element) is handled as one element
// with NY = NFY*NLY state dofs, field-major, whose dof ids were renumbered at create time to the
// concatenated (ConsecutiveMultiFieldStyle) ids of the whole state / control space.
// GEO_ = true: per-cell geometry (desc.cell_jinvT): the tables hold REFERENCE gradients / weights and every cell applies
// its own J^-T and det(J) (any Gridap triangulation); false: the uniform fast path with physical gradients folded into
// the tables (identical affine cells, e.g. a CartesianDiscreteModel without `map`).
template <class Integrand_, int NLY_, int NLU_, int NQ_, bool GEO_ = false>
struct Elem {
  using Integrand = Integrand_;
  static constexpr bool GEO = GEO_;
  static constexpr int DIM = Integrand::DIM, NP = Integrand::NP, M = Integrand::M;
  static constexpr int SY = Integrand::SY, SG = Integrand::SG, SU = Integrand::SU, SK = Integrand::SK;
  static constexpr int NFY = Integrand::NFY, NFU = NLU_ > 0 ? Integrand::NFU : 0;
  static constexpr int NLY = NLY_, NLU = NLU_, NLUS = NLU_ > 0 ? NLU_ : 1;
  static constexpr bool MULTI = NFY > 1 || NFU > 1;
  static constexpr int NY = NFY * NLY_, NU = NFU * NLU_, NQ = NQ_;
  static constexpr int NUS = NU > 0 ? NU : 1;
  static constexpr int NB = 1 + DIM;                                  // (phi, grad phi)
  static constexpr bool HAS_CONST = Integrand::HAS_CONST;
  static constexpr int NR = NFY * NB + (HAS_CONST ? 1 : 0);           // residual slots
  static constexpr int RUN_JY = NY * NY, RUN_JU = NY * NU;
  static constexpr int RUN_H = NY * (NY + 1) / 2 + NU * NY + NU * (NU + 1) / 2;  // distinct ids per cell
  static constexpr int RUN_J = RUN_JY > RUN_JU ? RUN_JY : RUN_JU;
};

template <class E>
struct Tables {
  double wdet[E::NQ];
  double By[E::NQ][E::NLY][E::NB];
  double Bu[E::NQ][E::NLUS];
  double par[8];
};

// Pre-integrated reference tensors for elements with nq > 1 (uniform geometry):
//   yy[k][l][j][i] = sum_q w_q By_i[k](q) By_j[l](q)   (Jacobian y-block with k = test slot; Hessian block (1,1))
//   uy[l][j][i]    = sum_q w_q psi_i(q)  By_j[l](q)    (Hessian block (2,1): rows u_i, cols y_j)
//   yu[t][b][i]    = sum_q w_q By_i[t](q) psi_b(q)     (Jacobian u-block: rows y_i, cols u_b)
//   uu[j][i]       = sum_q w_q psi_i(q)  psi_j(q)      (Hessian block (2,2))
// When the pointwise derivative data D of a cell is the same at all its quadrature points (bilinear /
// constant-coefficient forms; detected at run time, warp-uniformly) the element blocks are
// sum_{t,k} D[t][k] * K[t][k][.][.]: one pass, no per-quadrature-point read-modify-write in shared memory.
template <class E>
struct KTab {
  static constexpr bool ENABLED = E::NQ > 1 && !E::HAS_CONST && !E::MULTI && !E::GEO;  // (reference tensors need identical cells)
  static constexpr int NB = E::NB, NY = E::NY, NU = E::NU;
  static constexpr int OFF_YY = 0;
  static constexpr int OFF_UY = NB * NB * NY * NY;
  static constexpr int OFF_YU = OFF_UY + NB * NY * NU;
  static constexpr int OFF_UU = OFF_YU + NB * NU * NY;
  // used by the K-path vector kernels (kvector.cuh): lap[j][i] = sum_{d>=1} yy[d][d][j][i] (Laplace-type forms),
  // first-order tensors y1[k][j] = sum_q w_q By_j[k](q), u1[b] = sum_q w_q psi_b(q), and vol = sum_q w_q
  static constexpr int OFF_LAP = (OFF_UU + NU * NU + 1) & ~1;
  static constexpr int OFF_Y1 = OFF_LAP + NY * NY;
  static constexpr int OFF_U1 = OFF_Y1 + NB * NY;
  static constexpr int OFF_VOL = OFF_U1 + NU;
  // used by the tensor-core vector kernels (kmma.cuh): bt[t][q][i] = w_q By_i[t](q)
  static constexpr int OFF_BT = (OFF_VOL + 1 + 1) & ~1;
  static constexpr int TOTAL = ENABLED ? ((OFF_BT + NB * E::NQ * NY + 1) & ~1) : 0;
  __host__ __device__ static constexpr int bt(int t, int q, int i) { return OFF_BT + (t * E::NQ + q) * NY + i; }
  __host__ __device__ static constexpr int lap(int j, int i) { return OFF_LAP + j * NY + i; }
  __host__ __device__ static constexpr int y1(int k, int j) { return OFF_Y1 + k * NY + j; }
  __host__ __device__ static constexpr int u1(int b) { return OFF_U1 + b; }
  __host__ __device__ static constexpr int yy(int k, int l, int j, int i) { return OFF_YY + ((k * NB + l) * NY + j) * NY + i; }
  __host__ __device__ static constexpr int uy(int l, int j, int i) { return OFF_UY + (l * NY + j) * NU + i; }
  __host__ __device__ static constexpr int yu(int t, int b, int i) { return OFF_YU + (t * NU + b) * NY + i; }
  __host__ __device__ static constexpr int uu(int j, int i) { return OFF_UU + j * NU + i; }
};

// chunk of the cell-granular template path: first entry in the Jacobian y- / u-block and the FE Hessian block, cells (1..32)
struct __align__(32) ChunkRec { long long jy, ju, h, n; };

struct DevModel {
  const int32_t* dofs_y;
  const int32_t* dofs_u;
  const double* dir_y;
  const double* dir_u;
  const double* coef;      // [2][ncells][nq] (a lane's nq values share a line: one DRAM access, then L1 hits)
  int64_t ncells_coef;     // cells of the coefficient arrays (= ncells, except in sub-models: self-check sample, template tile)
  const int64_t* base_jy;  // [ntiles+1]
  const int64_t* base_ju;
  const int64_t* base_h;
  const double* ktab;      // pre-integrated reference tensors (KTab layout), nq > 1 elements only
  const double* jinvT;     // per-cell geometry, SoA [geom_nq][dim][dim][ldc]: grad_x = jinvT * grad_xi  (GEO elements only)
  const double* detj;      // [geom_nq][ldc]
  int32_t geom_nq;         // 1 (affine cells) or nq
  int64_t ncells, ldc;
  int64_t nfree_y, nfree_u;
  int32_t ntiles;
  int32_t nparam;
  int32_t notrait;         // 1: ignore the trait-selected shortcuts (self-check: generic per-point evaluation)
  // ---- deterministic mode (PDENLP_FLAG_DETERMINISTIC): the dof-indexed kernels run once per COLOUR of a greedy cell
  // colouring (cells of one colour share no free dof), over the colour's cell list and with plain read-modify-write
  // instead of atomics: every dof receives its contributions in colour order — bitwise reproducible run to run
  const int32_t* cell_list;  // cells of the colour being processed (nullptr: all cells, lane = cell, atomics)
  int64_t nlist;
  // ---- template tiles (see "template tiles" below); tpl_list == nullptr: every tile takes the generic path
  const int32_t* tpl_list; // [ntpl] tiles whose image equals the template's, ascending
  const int32_t* gen_list; // [ngen] all other tiles, ascending
  int32_t ntpl, ngen;
  const double* tpl_jy;    // [32*RUN_JY] image of one template tile of the Jacobian y-block (model-owned scratch)
  const double* tpl_ju;    // [32*RUN_JU]
  const double* tpl_h;     // [32*RUN_H]  objective FE Hessian block
  // ---- cell-granular templates (see "cell-granular templates" below); ch == nullptr: not in use.  The slab's cells
  // are either members of a CHUNK (<= 32 consecutive template cells = one contiguous run of every COO segment, stored
  // from the shared-memory image with one bulk copy) or GENERIC cells (evaluated one per lane, stored directly).
  const ChunkRec* ch;      // [nchunk] one 32 B record per chunk: a copy CTA reads ONE sector before its first store
  const int32_t* gc_cell;  // [ngc] generic cells, ascending
  const int64_t* gc_jy;    // [ngc] first entry of the cell in each segment
  const int64_t* gc_ju;
  const int64_t* gc_h;
  int32_t nchunk, ngc;
  // ---- last-CTA reductions (see finish_partials): arrival counter of the running kernel (reset by the last CTA),
  // device scalar of the objective and its pinned host mirror (written by the kernel: no copy after the launch)
  unsigned* ticket;
  double* fsum;
  double* fsum_host;
};

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// Programmatic dependent launch (all no-ops when the kernel was not launched with the stream-serialization attribute):
// wait = every earlier kernel of the stream has completed and its writes are visible; launch_dependents = the next
// kernel of the stream may start its prologue.  Protocol of the coord kernels: read model-owned data (dof ids,
// reference tensors) -> wait -> launch_dependents -> touch the caller's vectors.  A kernel therefore never reads or
// writes caller memory before its predecessors are done, and its completion implies theirs.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <int N>  // at most N of this thread's most recent bulk groups may still be reading shared memory
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// cell processed by lane `lane` of tile `tile`: tile*32 + lane, or the entry of the colour's cell list (deterministic mode)
__device__ __forceinline__ int64_t lane_cell(const DevModel& D, int64_t tile, int lane, bool& valid) {
  const int64_t idx = tile * 32 + lane;
  if (D.cell_list == nullptr) { valid = tile < D.ntiles && idx < D.ncells; return valid ? idx : 0; }
  valid = tile < D.ntiles && idx < D.nlist;
  return valid ? (int64_t)__ldg(D.cell_list + idx) : 0;
}
// out[i] += v: atomically, or as a plain read-modify-write when the launch covers one colour (no two cells of the
// launch touch the same dof)
__device__ __forceinline__ void accumulate(const DevModel& D, double* p, double v) {
  if (D.cell_list == nullptr) atomicAdd(p, v);
  else *p += v;
}

// ------------------------------------------------------------------ per-lane cell data
template <class E>
struct Cell {
  int gy[E::NY];
  int gu[E::NUS];
  double zy[E::NY];
  double zu[E::NUS];
};

template <class E>
__device__ __forceinline__ void load_ids(const DevModel& D, int64_t cell, bool valid, Cell<E>& c) {
#pragma unroll
  for (int a = 0; a < E::NY; ++a) c.gy[a] = valid ? __ldg(D.dofs_y + (int64_t)a * D.ldc + cell) : -1;
#pragma unroll
  for (int b = 0; b < E::NUS; ++b) c.gu[b] = (valid && E::NU > 0) ? __ldg(D.dofs_u + (int64_t)b * D.ldc + cell) : -1;
}
// PosNegReindex(free, dirichlet): /root/reference/src/PDENLPModels.jl:27-43
template <class E>
__device__ __forceinline__ void gather_state(const DevModel& D, const double* __restrict__ x, bool valid, Cell<E>& c) {
  const double* xy = x + D.nparam;
  const double* xu = xy + D.nfree_y;
#pragma unroll
  for (int a = 0; a < E::NY; ++a) {
    const int g = c.gy[a];
    c.zy[a] = !valid ? 0.0 : (g > 0 ? __ldg(xy + (g - 1)) : __ldg(D.dir_y + (-g - 1)));
  }
#pragma unroll
  for (int b = 0; b < E::NUS; ++b) {
    const int g = c.gu[b];
    c.zu[b] = (!valid || E::NU == 0) ? 0.0 : (g > 0 ? __ldg(xu + (g - 1)) : __ldg(D.dir_u + (-g - 1)));
  }
}
// values of a free-dof vector through a space whose Dirichlet values are zero
// (lambda through the TEST space, /root/reference/src/GridapPDENLPModel.jl:383; directions v)
template <int N>
__device__ __forceinline__ void gather_free(const double* __restrict__ vec, const int* gid, double* out) {
#pragma unroll
  for (int a = 0; a < N; ++a) {  // unconditional load (entry 0 for a non-free dof), then select: no branch
    const int g = gid[a];
    const double v = __ldg(vec + (g > 0 ? g - 1 : 0));
    out[a] = g > 0 ? v : 0.0;
  }
}

// quadrature point q of a cell: coefficient values, quadrature weight * det(J) and (GEO elements) the cell's J^-T
template <class E>
struct PointGeo : PointCtx {
  double w;                                   // quadrature weight * |det J|
  double T[E::GEO ? E::DIM : 1][E::GEO ? E::DIM : 1];  // J^-T at the point: grad_x = T * grad_xi
};
template <class E>
__device__ __forceinline__ PointGeo<E> point_ctx(const DevModel& D, const Tables<E>& tab, int64_t cell, bool valid, int q, bool need) {
  PointGeo<E> pc;
  pc.par = tab.par;
  pc.target = 0.0;
  pc.source = 0.0;
  if (need && valid) {
    pc.target = __ldg(D.coef + (cell * E::NQ + q));
    pc.source = __ldg(D.coef + ((D.ncells_coef + cell) * E::NQ + q));
  }
  pc.w = tab.wdet[q];
  if (E::GEO) {
    const int g = D.geom_nq > 1 ? q : 0;
    const int64_t c = valid ? cell : 0;
#pragma unroll
    for (int d = 0; d < E::DIM; ++d)
#pragma unroll
      for (int e = 0; e < E::DIM; ++e) pc.T[E::GEO ? d : 0][E::GEO ? e : 0] = __ldg(D.jinvT + ((int64_t)((g * E::DIM + d) * E::DIM + e) * D.ldc + c));
    pc.w *= __ldg(D.detj + ((int64_t)g * D.ldc + c));
  }
  return pc;
}

// pointwise fields at quadrature point q, seeded as jets (or plain doubles).  GEO elements: the gradient slots carry
// the PHYSICAL gradient T * grad_xi y as value and the row of T as seed, so the jets differentiate w.r.t. the
// reference slots (the chain rule through the cell map is done by the jet arithmetic).
template <class E, class T>
__device__ __forceinline__ void point_state(const Tables<E>& tab, const PointGeo<E>& pc, const Cell<E>& c, const double* __restrict__ x, int q, T* s) {
  double val[E::M];
#pragma unroll
  for (int k = 0; k < E::M; ++k) val[k] = 0.0;
#pragma unroll
  for (int f = 0; f < E::NFY; ++f)
#pragma unroll
    for (int a = 0; a < E::NLY; ++a)
#pragma unroll
      for (int k = 0; k < E::NB; ++k) val[f * E::NB + k] = fma(tab.By[q][a][k], c.zy[f * E::NLY + a], val[f * E::NB + k]);
#pragma unroll
  for (int f = 0; f < E::NFU; ++f)
#pragma unroll
    for (int b = 0; b < E::NLU; ++b) val[E::SU + f] = fma(tab.Bu[q][b], c.zu[f * E::NLU + b], val[E::SU + f]);
#pragma unroll
  for (int k = 0; k < E::NP; ++k) val[E::SK + k] = __ldg(x + k);
  if (E::GEO) {
#pragma unroll
    for (int k = 0; k < E::M; ++k) {
      const int f = k / E::NB, d = k % E::NB - 1;  // state field and gradient component of slot k
      if (k < E::NFY * E::NB && d >= 0) {
        double row[E::DIM], a = 0.0;
#pragma unroll
        for (int e = 0; e < E::DIM; ++e) { row[e] = pc.T[E::GEO ? d : 0][E::GEO ? e : 0]; a = fma(row[e], val[f * E::NB + 1 + e], a); }
        jet_seed_lin(s[k], a, f * E::NB + 1, E::DIM, row);
      } else {
        jet_seed(s[k], val[k], k);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < E::M; ++k) jet_seed(s[k], val[k], k);
  }
}
// the integrand's residual slots; GEO elements: the gradient slots multiply the PHYSICAL test gradient
// T * grad_xi v, so they are pulled back to the reference slots: R_ref[1+e] = sum_d T[d][e] R[1+d]
template <class E, class S>
__device__ __forceinline__ void eval_res(const PointGeo<E>& pc, const S* s, S* R) {
  E::Integrand::res(pc, s, R);
  if (E::GEO) {
#pragma unroll
    for (int f = 0; f < E::NFY; ++f) {
      S P[E::DIM];
#pragma unroll
      for (int e = 0; e < E::DIM; ++e) {
        P[e] = pc.T[0][E::GEO ? e : 0] * R[f * E::NB + 1];
#pragma unroll
        for (int d = 1; d < E::DIM; ++d) P[e] = P[e] + pc.T[E::GEO ? d : 0][E::GEO ? e : 0] * R[f * E::NB + 1 + d];
      }
#pragma unroll
      for (int e = 0; e < E::DIM; ++e) R[f * E::NB + 1 + e] = P[e];
    }
  }
}

// test-function slot t of local test function i: (phi_i, grad phi_i, [1])  (single-field elements)
template <class E>
__device__ __forceinline__ double test_slot(const Tables<E>& tab, int q, int i, int t) {
  static_assert(!E::MULTI, "single-field helper");
  return t < E::NB ? tab.By[q][i][t] : 1.0;
}
// sum_t T_i[t] * vec[t] for local test function i of test field f; vec is laid out like the residual
// slots: [f*NB + (value, gradient)] per test field, then the constant slot
template <class E>
__device__ __forceinline__ double test_contract(const Tables<E>& tab, int q, int f, int i, const double* vec) {
  double a = 0.0;
#pragma unroll
  for (int t = 0; t < E::NB; ++t) a = fma(tab.By[q][i][t], vec[f * E::NB + t], a);
  if (E::HAS_CONST) a += vec[E::NFY * E::NB];
  return a;
}

// ------------------------------------------------------------------ entry counts (reference filters)
template <class E>
__device__ __forceinline__ void count_entries(const Cell<E>& c, int& njy, int& nju, int& nh) {
  int fy = 0, fu = 0;
#pragma unroll
  for (int a = 0; a < E::NY; ++a) fy += c.gy[a] > 0;
#pragma unroll
  for (int b = 0; b < E::NU; ++b) fu += c.gu[b] > 0;
  njy = fy * fy;
  nju = fy * fu;
  int h = 0;
#pragma unroll
  for (int j = 0; j < E::NY; ++j)
#pragma unroll
    for (int i = 0; i < E::NY; ++i) h += (c.gy[j] > 0) & (c.gy[j] <= c.gy[i]);
  h += fu * fy;  // block (2,1): u ids are offset by nfree_y, so gidcol <= gidrow always holds
#pragma unroll
  for (int j = 0; j < E::NU; ++j)
#pragma unroll
    for (int i = 0; i < E::NU; ++i) h += (c.gu[j] > 0) & (c.gu[j] <= c.gu[i]);
  nh = h;
}

// Closed-form counts, valid when the dof ids of a cell are pairwise distinct (checked once at
// create time by k_count): among distinct ids exactly one of (i,j),(j,i) passes gidcol<=gidrow.
template <class E>
__device__ __forceinline__ void closed_counts(const Cell<E>& c, int& fy, int& fu, int& nh) {
  fy = 0; fu = 0;
#pragma unroll
  for (int a = 0; a < E::NY; ++a) fy += c.gy[a] > 0;
#pragma unroll
  for (int b = 0; b < E::NU; ++b) fu += c.gu[b] > 0;
  nh = fy * (fy + 1) / 2 + fu * fy + fu * (fu + 1) / 2;
}

// per-tile counts, used once at create time to build the segment offsets
template <class E>
__global__ void __launch_bounds__(kThreads) k_count(DevModel D, int32_t* __restrict__ counts /*[ntiles][3]*/, int32_t* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= D.ntiles) return;
  const int64_t cell = tile * 32 + lane;
  const bool valid = cell < D.ncells;
  Cell<E> c;
  load_ids<E>(D, valid ? cell : 0, valid, c);
  int njy, nju, nh;
  count_entries<E>(c, njy, nju, nh);
  int fy, fu, nh_closed;
  closed_counts<E>(c, fy, fu, nh_closed);
  if (nh != nh_closed) atomicExch(flag, 1);  // repeated dof ids inside one cell are not supported
  for (int d = 16; d > 0; d >>= 1) {
    njy += __shfl_xor_sync(0xffffffffu, njy, d);
    nju += __shfl_xor_sync(0xffffffffu, nju, d);
    nh += __shfl_xor_sync(0xffffffffu, nh, d);
  }
  if (lane == 0) { counts[tile * 3 + 0] = njy; counts[tile * 3 + 1] = nju; counts[tile * 3 + 2] = nh; }
}

// ------------------------------------------------------------------ staged segment writer
// One per warp.  The tile's run of one COO segment is staged in shared memory and streamed to HBM
// with TMA bulk stores.  Lane l writes its k-th kept entry at (exclusive scan + k), i.e. lanes are
// RUN doubles apart: with an even RUN that is a 2..16-way bank conflict on every STS.64 (measured:
// 3.6e9 conflicts per launch for the 3-D Q1 element).  The image is therefore cut into P pieces of
// CPP = 32/P consecutive cells; pieces are PSTRIDE doubles apart (see below), which spreads the lanes
// over the banks, and every piece is one contiguous global range streamed by its own bulk store,
// issued by the piece's first lane.  `delta` keeps each piece congruent (mod 16 B) with its global
// destination so the bulk copy is legal; unaligned head/tail elements go by plain 8 B stores.
// RUN odd => P = 1: one piece, one bulk store per tile, no conflicts to begin with.
template <int RUN>
struct StageLayout {
  // Pieces cost TMA efficiency (8 stores of 1.1 KB were measured 11% slower than one of 9.2 KB on the
  // HBM-bound membrane u-block), so they are used only where the conflict is >= 8-way (RUN % 8 == 0).
  static constexpr int P = (RUN % 8 == 0) ? 16 : 1;
  static constexpr int CPP = 32 / P;                       // cells (lanes) per piece
  static constexpr int PR = CPP * RUN + 2;                 // piece capacity incl. parity slack
#ifdef PDENLP_EXP_ODD_PSTRIDE
  static constexpr int PSTRIDE = (P == 1) ? PR : (PR | 1); // odd stride between pieces
#else
  // Every piece image starts at a double offset congruent (mod 2) with its global range, and the ranges of one tile
  // almost always share their parity (even entry counts), so the piece starts can only spread over the 8 even (or
  // odd) bank pairs: the stride must make p*PSTRIDE/2 run through all residues mod 8 for the 8 pieces of a half-warp,
  // i.e. PSTRIDE = 2 (mod 4).  PR = CPP*RUN + 2 with RUN % 8 == 0 already is.  (An odd stride looks conflict-free on
  // paper but the parity shift folds pieces p and p+5 onto the same banks: 8 wavefronts per STS.64 measured, 4 now.)
  static constexpr int PSTRIDE = PR;
  static_assert(P == 1 || PR % 4 == 2, "piece stride must be 2 mod 4");
#endif
  static constexpr int WARP_DOUBLES = (P * PSTRIDE + 2) & ~1;  // even => 16 B aligned warp buffers
};

template <int RUN>
struct SegStage {
  using L = StageLayout<RUN>;
  double* wbuf;    // this warp's staging buffer (16 B aligned)
  double* piece;   // start of this lane's piece image (wbuf + piece*PSTRIDE + delta)
  // make sure the previous bulk stores have finished READING the buffer before refilling it
#ifdef PDENLP_COPYOUT_STG
  // Experiment (-DPDENLP_COPYOUT_STG): copy-out with posted vector stores instead of TMA.  Measured
  // 33% slower on B200 (step 2.54 vs 1.91 ms at config 2), so TMA bulk stores are the default.
  template <int NBUF = 1> __device__ __forceinline__ void acquire(int) {}
#else
  // (lanes wait on their own bulk groups; the issuing lanes are the first lanes of the finest
  // piece layout, 16 pieces = even lanes, which covers every coarser layout sharing the buffer)
  // NBUF > 1: the warp rotates over NBUF staging buffers (small elements), so only the store issued NBUF
  // uses ago must have left shared memory
  template <int NBUF = 1>
  __device__ __forceinline__ void acquire(int lane) {
    if ((lane & 1) == 0) bulk_wait_read<NBUF - 1>();
    __syncwarp();
  }
#endif
  // where this lane writes its first entry; gbase = global address of the tile's first entry
  __device__ __forceinline__ double* begin(const double* gbase, int excl, int lane) {
    if (L::P == 1) {
      piece = wbuf + ((reinterpret_cast<uintptr_t>(gbase) >> 3) & 1);
      return piece + excl;
    }
    const int l0 = lane & ~(L::CPP - 1);
    const int pe = __shfl_sync(0xffffffffu, excl, l0);
    double* region = wbuf + (lane / L::CPP) * L::PSTRIDE;
    const unsigned gpar = (unsigned)(reinterpret_cast<uintptr_t>(gbase + pe) >> 3) & 1u;
    const unsigned spar = (smem_u32(region) >> 3) & 1u;
    piece = region + (gpar ^ spar);
    return piece + (excl - pe);
  }
  // all lanes have written their entries; stream every piece to its global range
#ifdef PDENLP_COPYOUT_STG
  // Cooperative copy-out: the (sub-)warp owning a piece reads its image with LDS.128 and writes
  // it with coalesced 16 B stores.
  __device__ __forceinline__ void release(double* gbase, int excl, int incl, int lane) {
    __syncwarp();
    const int l0 = lane & ~(L::CPP - 1);
    const int pe = L::P == 1 ? 0 : __shfl_sync(0xffffffffu, excl, l0);
    const int pend = __shfl_sync(0xffffffffu, incl, l0 + L::CPP - 1);
    const int n = pend - pe;
    if (n > 0) {
      double* g = gbase + pe;
      const int head = (int)((reinterpret_cast<uintptr_t>(g) >> 3) & 1);  // unaligned head element
      const int sub = lane - l0;
      if (head && sub == 0) g[0] = piece[0];
      const int n2 = (n - head) >> 1;  // 16 B granules
      const double2* src = reinterpret_cast<const double2*>(piece + head);
      double2* dst = reinterpret_cast<double2*>(g + head);
      for (int i = sub; i < n2; i += L::CPP) __stcs(dst + i, src[i]);
      if (((n - head) & 1) && sub == 0) g[n - 1] = piece[n - 1];  // tail element
    }
    __syncwarp();
  }
#else
  __device__ __forceinline__ void release(double* gbase, int excl, int incl, int lane) {
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the async proxy (TMA)
    __syncwarp();
    const int l0 = lane & ~(L::CPP - 1);
    const int pe = L::P == 1 ? 0 : __shfl_sync(0xffffffffu, excl, l0);
    const int pend = __shfl_sync(0xffffffffu, incl, l0 + L::CPP - 1);
    if (lane == l0 && pend > pe) {
      const int n = pend - pe;
      double* g = gbase + pe;
      int done = 0;
      if ((reinterpret_cast<uintptr_t>(g) >> 3) & 1) { g[0] = piece[0]; done = 1; }  // unaligned head element
      const int nb = (n - done) & ~1;                                                   // 16 B granules
#ifndef PDENLP_EXP_NO_STORE  // experiment: SM-side work only
      if (nb > 0) {
        bulk_store_s2g(g + done, piece + done, (uint32_t)nb * 8u);
        bulk_commit();
      }
#endif
      if ((n - done) & 1) g[n - 1] = piece[n - 1];  // tail element
    }
  }
#endif
#ifdef PDENLP_COPYOUT_STG
  __device__ __forceinline__ void drain(int) {}
#else
  __device__ __forceinline__ void drain(int lane) { if ((lane & 1) == 0) bulk_wait_all(); }
#endif
  // the whole tile image is zero (see fe_hess_is_zero): vectorised cooperative fill instead of
  // per-lane ranked emission
  __device__ __forceinline__ void zero_fill(int total, int lane) {
    double2* b = reinterpret_cast<double2*>(wbuf);
    const int n2 = L::P == 1 ? (total + 3) / 2 : L::WARP_DOUBLES / 2;
    for (int i = lane; i < n2; i += 32) b[i] = make_double2(0.0, 0.0);
  }
};

// ---- K-path emitters (see KTab).  D entries that are zero on every lane of the warp are skipped
// through warp-uniform branches, so e.g. a Poisson-type form costs 3 of the 16 (t,k) terms.
// Common sparsity patterns of the NB x NB pointwise derivative data, as bit masks over (t*NB + k): the
// emitters are instantiated for them with the pattern as a compile-time constant (straight-line code, no
// per-column uniform tests); any other pattern takes the generic instance (CMASK = 0) with run-time tests.
template <class E> __host__ __device__ constexpr unsigned kmask_mass() { return 1u; }                    // (0,0): reaction / L2 terms
template <class E> __host__ __device__ constexpr unsigned kmask_diag_grad() {                            // (d,d), d >= 1: Laplace-type terms
  unsigned m = 0;
  for (int d = 1; d < E::NB; ++d) m |= 1u << (d * E::NB + d);
  return m;
}

template <class E, unsigned CMASK>
__device__ __forceinline__ double* emit_jac_y_K_impl(const double* __restrict__ Ks, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c,
                                                     double* p, unsigned nz) {
  using KT = KTab<E>;
  // every column keeps the same rows (the free ones): one pointer per row, advanced by the number of kept
  // rows per kept column — independent address chains instead of one pointer bumped after every store
  double* rp[E::NY];
  int fy = 0;
#pragma unroll
  for (int i = 0; i < E::NY; ++i) { rp[i] = p + fy; fy += c.gy[i] > 0; }
  int ncol = 0;
#pragma unroll
  for (int j = 0; j < E::NY; ++j) {
    if (c.gy[j] > 0) {
      double acc[E::NY];
#pragma unroll
      for (int i = 0; i < E::NY; ++i) acc[i] = 0.0;
#pragma unroll
      for (int t = 0; t < E::NB; ++t)
#pragma unroll
        for (int k = 0; k < E::NB; ++k)
          if (((CMASK ? CMASK : nz) >> (t * E::NB + k)) & 1u) {
            const double d = R[t].g[k];
#pragma unroll
            for (int i = 0; i < E::NY; ++i) acc[i] = fma(d, Ks[KT::yy(t, k, j, i)], acc[i]);
          }
#pragma unroll
      for (int i = 0; i < E::NY; ++i) {
        if (c.gy[i] > 0) *rp[i] = acc[i];
        rp[i] += fy;
      }
      ++ncol;
    }
  }
  return p + ncol * fy;
}
template <class E>
__device__ __forceinline__ double* emit_jac_y_K(const double* __restrict__ Ks, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, double* p) {
  unsigned nz = 0;
#pragma unroll
  for (int t = 0; t < E::NB; ++t)
#pragma unroll
    for (int k = 0; k < E::NB; ++k)
      if (__any_sync(0xffffffffu, R[t].g[k] != 0.0)) nz |= 1u << (t * E::NB + k);
  constexpr unsigned DG = kmask_diag_grad<E>();
  if (nz == DG) return emit_jac_y_K_impl<E, DG>(Ks, R, c, p, nz);
  return emit_jac_y_K_impl<E, 0u>(Ks, R, c, p, nz);
}
template <class E>
__device__ __forceinline__ double* emit_jac_u_K(const double* __restrict__ Ks, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, double* p) {
  using KT = KTab<E>;
  unsigned nz = 0;
#pragma unroll
  for (int t = 0; t < E::NB; ++t)
    if (__any_sync(0xffffffffu, R[t].g[E::SU] != 0.0)) nz |= 1u << t;
  double* rp[E::NY];
  int fy = 0;
#pragma unroll
  for (int i = 0; i < E::NY; ++i) { rp[i] = p + fy; fy += c.gy[i] > 0; }
  int ncol = 0;
#pragma unroll
  for (int b = 0; b < E::NU; ++b) {
    if (c.gu[b] > 0) {
      double acc[E::NY];
#pragma unroll
      for (int i = 0; i < E::NY; ++i) acc[i] = 0.0;
#pragma unroll
      for (int t = 0; t < E::NB; ++t)
        if ((nz >> t) & 1u) {
          const double d = R[t].g[E::SU];
#pragma unroll
          for (int i = 0; i < E::NY; ++i) acc[i] = fma(d, Ks[KT::yu(t, b, i)], acc[i]);
        }
#pragma unroll
      for (int i = 0; i < E::NY; ++i) {
        if (c.gy[i] > 0) *rp[i] = acc[i];
        rp[i] += fy;
      }
      ++ncol;
    }
  }
  return p + ncol * fy;
}
template <class E, unsigned CMASK>
__device__ __forceinline__ double* emit_hess_yy_K_impl(const double* __restrict__ Ks, const Jet2<E::M>& L, double sc, const Cell<E>& c,
                                                       double* p, unsigned nz) {
  using KT = KTab<E>;
  using J2 = Jet2<E::M>;
#pragma unroll
  for (int j = 0; j < E::NY; ++j) {
    if (c.gy[j] > 0) {
      double acc[E::NY];
#pragma unroll
      for (int i = 0; i < E::NY; ++i) acc[i] = 0.0;
#pragma unroll
      for (int k = 0; k < E::NB; ++k)
#pragma unroll
        for (int l = 0; l < E::NB; ++l)
          if (((CMASK ? CMASK : nz) >> (k * E::NB + l)) & 1u) {
            const double d = L.h[J2::idx(k, l)] * sc;
#pragma unroll
            for (int i = 0; i < E::NY; ++i) acc[i] = fma(d, Ks[KT::yy(k, l, j, i)], acc[i]);
          }
#pragma unroll
      for (int i = 0; i < E::NY; ++i)
        if (c.gy[j] <= c.gy[i]) { *p = acc[i]; ++p; }
    }
  }
  return p;
}
template <class E>
__device__ __forceinline__ double* emit_hess_yy_K(const double* __restrict__ Ks, const Jet2<E::M>& L, double sc, const Cell<E>& c, double* p) {
  using J2 = Jet2<E::M>;
  unsigned nz = 0;
#pragma unroll
  for (int k = 0; k < E::NB; ++k)
#pragma unroll
    for (int l = 0; l < E::NB; ++l)
      if (__any_sync(0xffffffffu, L.h[J2::idx(k, l)] != 0.0)) nz |= 1u << (k * E::NB + l);
  constexpr unsigned MS = kmask_mass<E>(), DG = kmask_diag_grad<E>();
  if (nz == MS) return emit_hess_yy_K_impl<E, MS>(Ks, L, sc, c, p, nz);
  if (nz == DG) return emit_hess_yy_K_impl<E, DG>(Ks, L, sc, c, p, nz);
  return emit_hess_yy_K_impl<E, 0u>(Ks, L, sc, c, p, nz);
}
template <class E>
__device__ __forceinline__ double* emit_hess_u_K(const double* __restrict__ Ks, const Jet2<E::M>& L, double sc, const Cell<E>& c, double* p) {
  using KT = KTab<E>;
  using J2 = Jet2<E::M>;
  if (E::NU > 0) {
    unsigned nz = 0;
#pragma unroll
    for (int l = 0; l < E::NB; ++l)
      if (__any_sync(0xffffffffu, L.h[J2::idx(E::SU, l)] != 0.0)) nz |= 1u << l;
#pragma unroll
    for (int j = 0; j < E::NY; ++j) {
      if (c.gy[j] > 0) {
        double acc[E::NUS];
#pragma unroll
        for (int i = 0; i < E::NUS; ++i) acc[i] = 0.0;
#pragma unroll
        for (int l = 0; l < E::NB; ++l)
          if ((nz >> l) & 1u) {
            const double d = L.h[J2::idx(E::SU, l)] * sc;
#pragma unroll
            for (int i = 0; i < E::NU; ++i) acc[i] = fma(d, Ks[KT::uy(l, j, i)], acc[i]);
          }
#pragma unroll
        for (int i = 0; i < E::NU; ++i)
          if (c.gu[i] > 0) { *p = acc[i]; ++p; }
      }
    }
    const double duu = L.h[J2::idx(E::SU, E::SU)] * sc;
#pragma unroll
    for (int j = 0; j < E::NU; ++j) {
      if (c.gu[j] > 0) {
#pragma unroll
        for (int i = 0; i < E::NU; ++i)
          if (c.gu[j] <= c.gu[i]) { *p = duu * Ks[KT::uu(j, i)]; ++p; }
      }
    }
  }
  return p;
}
// the warp's CTA stages the tensors in shared memory once (after the staging buffers)
template <class E>
__device__ __forceinline__ const double* load_ktab(const DevModel& D, double* dst) {
  if constexpr (KTab<E>::ENABLED) {
    if (D.ktab == nullptr) return nullptr;
    for (int i = threadIdx.x; i < KTab<E>::TOTAL; i += blockDim.x) dst[i] = __ldg(D.ktab + i);
    __syncthreads();
    return dst;
  } else {
    return nullptr;
  }
}

template <class E>
__host__ __device__ constexpr int jac_stage_doubles() {
  return StageLayout<E::RUN_JY>::WARP_DOUBLES > StageLayout<(E::NU > 0 ? E::RUN_JU : 1)>::WARP_DOUBLES
             ? StageLayout<E::RUN_JY>::WARP_DOUBLES
             : StageLayout<(E::NU > 0 ? E::RUN_JU : 1)>::WARP_DOUBLES;
}
template <class E>
__host__ __device__ constexpr int hess_stage_doubles() { return StageLayout<E::RUN_H>::WARP_DOUBLES; }
// Staging buffers per warp: elements whose tile image is small rotate over up to 4 buffers so that a warp does
// not wait for the bulk store of its previous tile to drain shared memory before staging the next one.
#ifndef PDENLP_NBUF_BYTES
#define PDENLP_NBUF_BYTES 8192
#endif
__host__ __device__ constexpr int coord_nbuf(int stage_doubles) {
  return stage_doubles * 8 * 2 > PDENLP_NBUF_BYTES ? 1 : (PDENLP_NBUF_BYTES / (stage_doubles * 8) > 4 ? 4 : PDENLP_NBUF_BYTES / (stage_doubles * 8));
}

// ------------------------------------------------------------------ template tiles
// For an integrand whose FE-field derivatives are the same in every cell (an affine residual with constant coefficients,
// a quadratic objective: trait FE_DERIVS_CELL_INDEPENDENT) on identical cells, the COO image of a tile depends only on
// which local dofs are free and on the relative order of the cell's global ids (the lower-triangle filter).  On a
// structured mesh almost every tile is "all 32 cells valid, all dofs free, same id order as the template cell"
// (config 2: 96.8 % of the tiles): the image of ONE such tile is computed by the ordinary kernel on a one-tile sub-model
// (pdenlp_cuda.cu: build_template), staged once per CTA in shared memory in both 16 B parities, and every template tile
// is a single TMA bulk store from that image — no dof ids, no x, no arithmetic: the kernel is a pure write stream
// (DRAM reads ~0.1 % of the writes).  The remaining tiles (gen_list) take the generic path below.  Tiles are
// classified once at create time (k_tile_class); the create-time self-check compares both paths on a sample that
// contains template tiles.
template <class E>
__device__ __forceinline__ void cell_signature(const Cell<E>& c, bool valid, bool& all_free, unsigned long long& sy, unsigned long long& su) {
  all_free = valid;
  sy = 0ull; su = 0ull;
#pragma unroll
  for (int a = 0; a < E::NY; ++a) {
    all_free = all_free && c.gy[a] > 0;
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < E::NY; ++b) r += c.gy[b] < c.gy[a];
    sy |= (unsigned long long)(r & 15u) << (4 * (a & 15));
  }
#pragma unroll
  for (int a = 0; a < E::NU; ++a) {
    all_free = all_free && c.gu[a] > 0;
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < E::NU; ++b) r += c.gu[b] < c.gu[a];
    su |= (unsigned long long)(r & 15u) << (4 * (a & 15));
  }
}
// cls[tile] = 1 iff all 32 cells of the tile are valid, all-free and carry the template's id-order signature
template <class E>
__global__ void __launch_bounds__(kThreads) k_tile_class(DevModel D, unsigned long long sig_y, unsigned long long sig_u, uint8_t* __restrict__ cls) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= D.ntiles) return;
  const int64_t cell = tile * 32 + lane;
  const bool valid = cell < D.ncells;
  Cell<E> c;
  load_ids<E>(D, valid ? cell : 0, valid, c);
  bool af;
  unsigned long long sy, su;
  cell_signature<E>(c, valid, af, sy, su);
  const bool ok = __all_sync(0xffffffffu, af && sy == sig_y && su == sig_u);
  if (lane == 0) cls[tile] = ok ? 1 : 0;
}
// mask[tile] bit l = 1 iff cell 32*tile + l is a template cell (valid, all dofs free, template id order)
template <class E>
__global__ void __launch_bounds__(kThreads) k_cell_class(DevModel D, unsigned long long sig_y, unsigned long long sig_u, uint32_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= D.ntiles) return;
  const int64_t cell = tile * 32 + lane;
  const bool valid = cell < D.ncells;
  Cell<E> c;
  load_ids<E>(D, valid ? cell : 0, valid, c);
  bool af;
  unsigned long long sy, su;
  cell_signature<E>(c, valid, af, sy, su);
  const unsigned b = __ballot_sync(0xffffffffu, af && sy == sig_y && su == sig_u);
  if (lane == 0) mask[tile] = b;
}
// one contiguous run of n doubles from a shared-memory image to global memory; img0 / img1 are the two copies of the
// image whose FIRST element sits at an even / odd double offset (bulk copies need 16 B alignment on both sides)
__device__ __forceinline__ void store_template(double* g, const double* img0, const double* img1, int n) {
  const bool odd = (reinterpret_cast<uintptr_t>(g) >> 3) & 1;
  const double* img = odd ? img1 : img0;
  int done = 0;
  if (odd) { g[0] = img[0]; done = 1; }
  const int nb = (n - done) & ~1;
  if (nb > 0) {
    bulk_store_s2g(g + done, img + done, (uint32_t)nb * 8u);
    bulk_commit();
  }
  if ((n - done) & 1) g[n - 1] = img[n - 1];
}
// shared-memory slot of one image copy (two spare doubles: parity shift + alignment)
__host__ __device__ constexpr int tpl_slot(int n) { return (n + 2 + 1) & ~1; }
__device__ __forceinline__ void load_template(const double* __restrict__ g, double* t0, double* t1, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = __ldg(g + i);
    t0[i] = v;
    t1[i + 1] = v;
  }
}

// ---- per-point emitters of the Jacobian blocks (single-field elements): quadrature point q adds its contribution to
// the cell's kept entries at p[0..) (q == 0 stores, q > 0 accumulates); column dof outer, row dof inner
//   y-block:  A_e[i][j] = sum_q w sum_t T_i[t] sum_k dR_t/ds_k By_j[k]
template <class E>
__device__ __forceinline__ void emit_jac_y_pt(const Tables<E>& tab, int q, double w, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, double* p) {
#pragma unroll
  for (int j = 0; j < E::NY; ++j) {
    double Mj[E::NR];
#pragma unroll
    for (int t = 0; t < E::NR; ++t) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < E::NB; ++k) a = fma(R[t].g[k], tab.By[q][j][k], a);
      Mj[t] = a * w;
    }
    if (c.gy[j] > 0) {
#pragma unroll
      for (int i = 0; i < E::NY; ++i) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < E::NR; ++t) v = fma(test_slot<E>(tab, q, i, t), Mj[t], v);
        if (c.gy[i] > 0) { if (q == 0) *p = v; else *p += v; ++p; }
      }
    }
  }
}
//   u-block:  B_e[i][b] = sum_q w sum_t T_i[t] dR_t/du psi_b
template <class E>
__device__ __forceinline__ void emit_jac_u_pt(const Tables<E>& tab, int q, double w, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, double* p) {
#pragma unroll
  for (int b = 0; b < E::NU; ++b) {
    double Mb[E::NR];
#pragma unroll
    for (int t = 0; t < E::NR; ++t) Mb[t] = R[t].g[E::SU] * tab.Bu[q][b] * w;
    if (c.gu[b] > 0) {
#pragma unroll
      for (int i = 0; i < E::NY; ++i) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < E::NR; ++t) v = fma(test_slot<E>(tab, q, i, t), Mb[t], v);
        if (c.gy[i] > 0) { if (q == 0) *p = v; else *p += v; ++p; }
      }
    }
  }
}

// ------------------------------------------------------------------ jac_coord!
// vals layout (this slab): [k-block (ncon*p, filled elsewhere) | y-block | u-block]
// Small elements (staging of a few KB per warp) leave most of the SM idle with one 8-warp CTA, so
// their register budget is capped to let two CTAs share an SM; the host sizes the persistent grid
// with the occupancy API.
template <class E>
constexpr int coord_min_blocks() { return (E::RUN_H <= 48 && E::NQ == 1) ? 2 : 1; }
// Several quadrature points without the reference-tensor path (a residual / objective that is not affine in the
// fields, e.g. the sinh term with degree 2): small elements sum their element blocks over the points in registers
// and stage them once, instead of a shared-memory read-modify-write per entry and point.
// (Integrands that declare FE_DERIVS_POINT_INDEPENDENT take the reference-tensor path instead.)
template <class E>
__host__ __device__ constexpr bool coord_reg_accumulate() {
  return E::NQ > 1 && E::NY <= 4 && E::NU <= 4 && !E::MULTI &&
         !(KTab<E>::ENABLED && E::Integrand::FE_DERIVS_POINT_INDEPENDENT);
}
// Small elements also prefetch the next tile's segment offsets (a dependent global load per segment that is a
// visible part of their short per-tile time); for the large elements the extra live registers cost more than the
// load latency (measured on config 2: jac 0.67 -> 0.73 ms), so they load the offsets at the point of use.
// Measured: config 3 (Q1/Q1 2-D, RUN_H = 36) jac 0.081 -> 0.064 ms but hess 0.124 -> 0.136 ms (its Hessian kernel sits at
// the 128-register cap of two CTAs per SM); config 4 (P1/P1 1-D, RUN_H = 10) jac 0.046 -> 0.040, hess 0.105 -> 0.101 ms.
#ifndef PDENLP_PREFETCH_BASE_RUN_J
#define PDENLP_PREFETCH_BASE_RUN_J 48
#endif
#ifndef PDENLP_PREFETCH_BASE_RUN_H
#define PDENLP_PREFETCH_BASE_RUN_H 16
#endif
template <class E, bool HESS>
__host__ __device__ constexpr bool coord_prefetch_base() {
  return E::RUN_H <= (HESS ? PDENLP_PREFETCH_BASE_RUN_H : PDENLP_PREFETCH_BASE_RUN_J);
}

// which coordinate kernels have a template-tile path (HESS: the objective FE block)
template <class E, bool HESS>
__host__ __device__ constexpr bool tpl_ok() {
  using I = typename E::Integrand;
  return !E::MULTI && !E::GEO && !E::HAS_CONST && I::FE_DERIVS_CELL_INDEPENDENT && I::FE_DERIVS_POINT_INDEPENDENT &&
         I::RES_FE_HESSIAN_ZERO && !(HESS && I::NO_FE_OBJ) &&
         coord_nbuf(HESS ? hess_stage_doubles<E>() : jac_stage_doubles<E>()) == 1;
}

template <class E>
__global__ void __launch_bounds__(kThreads, coord_min_blocks<E>())
k_jac_coord(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
            double* __restrict__ vals_y, double* __restrict__ vals_u) {
  using I = typename E::Integrand;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int STAGE = jac_stage_doubles<E>(), NBUF = coord_nbuf(STAGE);
  double* const wbase = smem + warp * STAGE * NBUF;
  int slot = 0;
  SegStage<E::RUN_JY> sty{wbase, nullptr};
  SegStage<(E::NU > 0 ? E::RUN_JU : 1)> stu{wbase, nullptr};
  const double* const Ks = load_ktab<E>(D, smem + nwarps * STAGE * NBUF);

  const int64_t stride = (int64_t)gridDim.x * nwarps;
  // ---- template tiles: one bulk store per tile and segment from the CTA's shared-memory image
  bool use_tpl = false;
  if constexpr (tpl_ok<E, false>()) {
    use_tpl = D.tpl_list != nullptr && !D.notrait;
    if (use_tpl) {
      constexpr int LY = 32 * E::RUN_JY, LU = 32 * E::RUN_JU;
      double* const ty0 = smem;
      double* const ty1 = ty0 + tpl_slot(LY);
      double* const tu0 = ty1 + tpl_slot(LY);
      double* const tu1 = tu0 + tpl_slot(LU);
      griddep_wait();  // (the image is written by the preceding one-tile launch on the stream)
      griddep_launch_dependents();
      load_template(D.tpl_jy, ty0, ty1, LY);
      if (E::NU > 0) load_template(D.tpl_ju, tu0, tu1, LU);
      fence_proxy_async_smem();
      __syncthreads();
      for (int64_t k = (int64_t)blockIdx.x * nwarps + warp; k < D.ntpl; k += stride) {
        const int64_t t = D.tpl_list[k];
        if (lane == 0) store_template(vals_y + D.base_jy[t], ty0, ty1 + 1, LY);
        if (E::NU > 0 && lane == 1) store_template(vals_u + D.base_ju[t], tu0, tu1 + 1, LU);
      }
      if (lane < 2) bulk_wait_read_all();  // the staging buffers below reuse this shared memory
      __syncthreads();
    }
  }
  // the generic path runs over gen_list (all tiles when there is no template)
  const int64_t ngen = use_tpl ? D.ngen : D.ntiles;
  auto tile_of = [&](int64_t k) -> int64_t { return use_tpl ? (int64_t)__ldg(D.gen_list + k) : k; };
  int64_t kt = (int64_t)blockIdx.x * nwarps + warp;
  if (kt >= ngen) { if (lane < 2) bulk_wait_all(); return; }
  int64_t tile = tile_of(kt);
  // Register software pipeline over tiles: while tile t is processed, the gathered values of
  // tile t+stride and the dof ids of tile t+2*stride are already in flight (the id -> value
  // gather is a two-level dependent load chain; without this it is ~30% of the warp's time).
  Cell<E> c, n;
  int64_t tnext = (kt + stride) < ngen ? tile_of(kt + stride) : 0;
  {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    load_ids<E>(D, valid ? cell : 0, valid, c);
    const int64_t ncell = tnext * 32 + lane;
    const bool nvalid = (kt + stride) < ngen && ncell < D.ncells;
    load_ids<E>(D, nvalid ? ncell : 0, nvalid, n);
    griddep_wait();               // from here on the caller's vectors are touched
    griddep_launch_dependents();
    gather_state<E>(D, x, valid, c);
  }
  constexpr bool PFB = coord_prefetch_base<E, false>();
  int64_t b_jy = PFB ? D.base_jy[tile] : 0, b_ju = (PFB && E::NU > 0) ? D.base_ju[tile] : 0;
  while (true) {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    const bool has_next = (kt + stride) < ngen;
    const bool has_next2 = (kt + 2 * stride) < ngen;
    const int64_t tnext2 = has_next2 ? tile_of(kt + 2 * stride) : 0;
    Cell<E> nn;
    int64_t nb_jy = 0, nb_ju = 0;
    {
      gather_state<E>(D, x, has_next && (tnext * 32 + lane) < D.ncells, n);
      const int64_t c2 = tnext2 * 32 + lane;
      const bool v2 = has_next2 && c2 < D.ncells;
      load_ids<E>(D, v2 ? c2 : 0, v2, nn);
      if (PFB && has_next) { nb_jy = D.base_jy[tnext]; if (E::NU > 0) nb_ju = D.base_ju[tnext]; }
    }

    int fy = 0, fu = 0;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) fy += c.gy[a] > 0;
#pragma unroll
    for (int b = 0; b < E::NU; ++b) fu += c.gu[b] > 0;

    // K-path test (nq > 1): is dR/d(y, grad y, u) the same at every quadrature point, on all lanes?
    bool kpath = false;
    Jet1<E::M> R0[E::NR];
    constexpr bool RA = coord_reg_accumulate<E>();
    double A[E::NY][E::NY], B[E::NUS][E::NY];  // element blocks summed over the points (RA elements only)
    if (RA) {  // one pass over the quadrature points for both blocks
#pragma unroll
      for (int j = 0; j < E::NY; ++j)
#pragma unroll
        for (int i = 0; i < E::NY; ++i) A[j][i] = 0.0;
#pragma unroll
      for (int b = 0; b < E::NUS; ++b)
#pragma unroll
        for (int i = 0; i < E::NY; ++i) B[b][i] = 0.0;
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet1<E::M> s[E::M], R[E::NR];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
#pragma unroll
        for (int j = 0; j < E::NY; ++j) {
          double Mj[E::NR];
#pragma unroll
          for (int t = 0; t < E::NR; ++t) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < E::NB; ++k) a = fma(R[t].g[k], tab.By[q][j][k], a);
            Mj[t] = a * pc.w;
          }
#pragma unroll
          for (int i = 0; i < E::NY; ++i)
#pragma unroll
            for (int t = 0; t < E::NR; ++t) A[j][i] = fma(test_slot<E>(tab, q, i, t), Mj[t], A[j][i]);
        }
#pragma unroll
        for (int b = 0; b < E::NU; ++b) {
          double Mb[E::NR];
#pragma unroll
          for (int t = 0; t < E::NR; ++t) Mb[t] = R[t].g[E::SU] * tab.Bu[q][b] * pc.w;
#pragma unroll
          for (int i = 0; i < E::NY; ++i)
#pragma unroll
            for (int t = 0; t < E::NR; ++t) B[b][i] = fma(test_slot<E>(tab, q, i, t), Mb[t], B[b][i]);
        }
      }
    }
    if (KTab<E>::ENABLED && Ks != nullptr && !RA) {
      bool uni = true;
      // integrands that declare FE_DERIVS_POINT_INDEPENDENT are evaluated at the first point only
      const int NQT = (I::FE_DERIVS_POINT_INDEPENDENT && !D.notrait) ? 1 : E::NQ;
#pragma unroll 1
      for (int q = 0; q < NQT; ++q) {
        Jet1<E::M> s[E::M], R[E::NR];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
        if (q == 0) {
#pragma unroll
          for (int t = 0; t < E::NR; ++t) R0[t] = R[t];
        } else {
#pragma unroll
          for (int t = 0; t < E::NR; ++t)
#pragma unroll
            for (int k = 0; k <= E::SU; ++k) uni = uni && (R[t].g[k] == R0[t].g[k]);
        }
      }
      kpath = __all_sync(0xffffffffu, uni || !valid);
    }

    // ---- y-block:  A_e[i][j] = sum_q w sum_t T_i[t] sum_k dR_t/ds_k By_j[k]
    {
      const int cnt = fy * fy;
      const int incl = warp_incl_scan(cnt, lane);
      double* gdst = vals_y + (PFB ? b_jy : D.base_jy[tile]);
      if (NBUF > 1) { sty.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      sty.template acquire<NBUF>(lane);
      double* const p0 = sty.begin(gdst, incl - cnt, lane);
      if (kpath) emit_jac_y_K<E>(Ks, R0, c, p0);
      else if (RA) {
        double* p = p0;
#pragma unroll
        for (int j = 0; j < E::NY; ++j)
          if (c.gy[j] > 0) {
#pragma unroll
            for (int i = 0; i < E::NY; ++i)
              if (c.gy[i] > 0) { *p = A[j][i]; ++p; }
          }
      } else
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet1<E::M> s[E::M], R[E::NR];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
        emit_jac_y_pt<E>(tab, q, pc.w, R, c, p0);
      }
      sty.release(gdst, incl - cnt, incl, lane);
    }
    // ---- u-block:  B_e[i][b] = sum_q w sum_t T_i[t] dR_t/du psi_b
    if (E::NU > 0) {
      const int cnt = fy * fu;
      const int incl = warp_incl_scan(cnt, lane);
      double* gdst = vals_u + (PFB ? b_ju : D.base_ju[tile]);
      if (NBUF > 1) { stu.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      stu.template acquire<NBUF>(lane);
      double* const p0 = stu.begin(gdst, incl - cnt, lane);
      if (kpath) emit_jac_u_K<E>(Ks, R0, c, p0);
      else if (RA) {
        double* p = p0;
#pragma unroll
        for (int b = 0; b < E::NU; ++b)
          if (c.gu[b] > 0) {
#pragma unroll
            for (int i = 0; i < E::NY; ++i)
              if (c.gy[i] > 0) { *p = B[b][i]; ++p; }
          }
      } else
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet1<E::M> s[E::M], R[E::NR];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
        emit_jac_u_pt<E>(tab, q, pc.w, R, c, p0);
      }
      stu.release(gdst, incl - cnt, incl, lane);
    }
    if (!has_next) break;
    c = n;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) n.gy[a] = nn.gy[a];
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) n.gu[b] = nn.gu[b];
    b_jy = nb_jy; b_ju = nb_ju;
    tile = tnext;
    tnext = tnext2;
    kt += stride;
  }
  sty.drain(lane);
  if (use_tpl && lane == 1) bulk_wait_all();  // (the u-block template stores were issued by lane 1)
}

// ------------------------------------------------------------------ hess_coord!
// Emit one FE Hessian block set of a cell from the pointwise second derivatives Dm (packed
// symmetric M x M, already scaled by the quadrature weight), in the reference order:
// block (1,1) [y,y], (2,1) [u rows, y cols], (1,2) [never kept], (2,2) [u,u]; within a block
// columns outer, rows inner; keep iff gidcol>0 and gidcol<=gidrow (global ids).
template <class E, bool ACC>
__device__ __forceinline__ double* emit_hess_yy(const Tables<E>& tab, int q, const Jet2<E::M>& L, double scale,
                                                const Cell<E>& c, double* p) {
  using J2 = Jet2<E::M>;
#pragma unroll
  for (int j = 0; j < E::NY; ++j) {
    double Nj[E::NB];
#pragma unroll
    for (int k = 0; k < E::NB; ++k) {
      double a = 0.0;
#pragma unroll
      for (int l = 0; l < E::NB; ++l) a = fma(L.h[J2::idx(k, l)], tab.By[q][j][l], a);
      Nj[k] = a * scale;
    }
    if (c.gy[j] > 0) {
#pragma unroll
      for (int i = 0; i < E::NY; ++i) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < E::NB; ++k) v = fma(tab.By[q][i][k], Nj[k], v);
        if (c.gy[j] <= c.gy[i]) { if (ACC) *p += v; else *p = v; ++p; }
      }
    }
  }
  return p;
}
// blocks (2,1) [rows = control dofs, cols = state dofs] and (2,2)
template <class E, bool ACC>
__device__ __forceinline__ double* emit_hess_u(const Tables<E>& tab, int q, const Jet2<E::M>& L, double scale,
                                               const Cell<E>& c, double* p) {
  using J2 = Jet2<E::M>;
  if (E::NU > 0) {
#pragma unroll
    for (int j = 0; j < E::NY; ++j) {
      double a = 0.0;
#pragma unroll
      for (int l = 0; l < E::NB; ++l) a = fma(L.h[J2::idx(E::SU, l)], tab.By[q][j][l], a);
      a *= scale;
      if (c.gy[j] > 0) {
#pragma unroll
        for (int i = 0; i < E::NU; ++i) {
          const double v = tab.Bu[q][i] * a;
          if (c.gu[i] > 0) { if (ACC) *p += v; else *p = v; ++p; }
        }
      }
    }
    const double duu = L.h[J2::idx(E::SU, E::SU)] * scale;
#pragma unroll
    for (int j = 0; j < E::NU; ++j) {
      const double a = duu * tab.Bu[q][j];
      if (c.gu[j] > 0) {
#pragma unroll
        for (int i = 0; i < E::NU; ++i) {
          const double v = tab.Bu[q][i] * a;
          if (c.gu[j] <= c.gu[i]) { if (ACC) *p += v; else *p = v; ++p; }
        }
      }
    }
  }
  return p;
}
template <class E, bool ACC>
__device__ __forceinline__ double* emit_hess(const Tables<E>& tab, int q, const Jet2<E::M>& L, double scale,
                                             const Cell<E>& c, double* p) {
  p = emit_hess_yy<E, ACC>(tab, q, L, scale, c, p);
  return emit_hess_u<E, ACC>(tab, q, L, scale, c, p);
}

// Run-time structural zero tests on the pointwise second derivatives w.r.t. the FE fields.
// UROW = false: the (y, grad y) x (y, grad y) part, which generates block (1,1);
// UROW = true : the u row (u x (y, grad y, u)), which generates blocks (2,1) and (2,2).
// These are data-dependent (not integrand-dependent) tests: both fire for the constraint block of
// every PDE that is linear in (y,u) — e.g. the membrane problem.
template <class E, bool UROW>
__device__ __forceinline__ bool fe_hess_block_is_zero(const Jet2<E::M>& L) {
  using J2 = Jet2<E::M>;
  bool z = true;
  if (!UROW) {
#pragma unroll
    for (int k = 0; k < E::NB; ++k)
#pragma unroll
      for (int l = 0; l <= k; ++l) z = z && (L.h[J2::idx(k, l)] == 0.0);
  } else if (E::NU > 0) {
#pragma unroll
    for (int l = 0; l <= E::SU; ++l) z = z && (L.h[J2::idx(E::SU, l)] == 0.0);
  }
  return z;
}

// lambda-weighted residual  l(s) = sum_t Lam_t R_t(s)  as a Jet2 (its Hessian is the pointwise
// Lagrangian Hessian; its gradient the pointwise J^T lambda)
template <class E, class T>
__device__ __forceinline__ T weighted_res(const PointGeo<E>& pc, const T* s, const double* Lam) {
  T R[E::NR];
  eval_res<E>(pc, s, R);
  T acc = Lam[0] * R[0];
#pragma unroll
  for (int t = 1; t < E::NR; ++t) acc = acc + Lam[t] * R[t];
  return acc;
}
template <class E>
__device__ __forceinline__ void lambda_slots(const Tables<E>& tab, int q, const double* lam_e, double* Lam) {
#pragma unroll
  for (int f = 0; f < E::NFY; ++f)
#pragma unroll
    for (int t = 0; t < E::NB; ++t) {
      double a = 0.0;
#pragma unroll
      for (int i = 0; i < E::NLY; ++i) a = fma(tab.By[q][i][t], lam_e[f * E::NLY + i], a);
      Lam[f * E::NB + t] = a;
    }
  if (E::HAS_CONST) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < E::NY; ++i) a += lam_e[i];
    Lam[E::NFY * E::NB] = a;
  }
}

// vals_obj / vals_con point at the FE blocks of this slab; either may be NULL (skipped).
template <class E>
__global__ void __launch_bounds__(kThreads, coord_min_blocks<E>())
k_hess_coord(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
             const double* __restrict__ lambda, double obj_weight,
             double* __restrict__ vals_obj, double* __restrict__ vals_con) {
  using I = typename E::Integrand;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int STAGE = hess_stage_doubles<E>(), NBUF = coord_nbuf(STAGE);
  double* const wbase = smem + warp * STAGE * NBUF;
  int slot = 0;
  SegStage<E::RUN_H> st{wbase, nullptr};
  const double* const Ks = load_ktab<E>(D, smem + nwarps * STAGE * NBUF);

  const int64_t stride = (int64_t)gridDim.x * nwarps;
  // ---- template tiles of the objective block (see k_jac_coord); only when it is the one block of this launch
  bool use_tpl = false;
  if constexpr (tpl_ok<E, true>()) {
    use_tpl = D.tpl_list != nullptr && !D.notrait && vals_obj != nullptr && vals_con == nullptr;
    if (use_tpl) {
      constexpr int LH = 32 * E::RUN_H;
      double* const th0 = smem;
      double* const th1 = th0 + tpl_slot(LH);
      griddep_wait();
      griddep_launch_dependents();
      load_template(D.tpl_h, th0, th1, LH);
      fence_proxy_async_smem();
      __syncthreads();
      for (int64_t k = (int64_t)blockIdx.x * nwarps + warp; k < D.ntpl; k += stride) {
        const int64_t t = D.tpl_list[k];
        if (lane == 0) store_template(vals_obj + D.base_h[t], th0, th1 + 1, LH);
      }
      if (lane == 0) bulk_wait_read_all();
      __syncthreads();
    }
  }
  const int64_t ngen = use_tpl ? D.ngen : D.ntiles;
  auto tile_of = [&](int64_t k) -> int64_t { return use_tpl ? (int64_t)__ldg(D.gen_list + k) : k; };
  int64_t kt = (int64_t)blockIdx.x * nwarps + warp;
  if (kt >= ngen) { if (lane == 0) bulk_wait_all(); return; }
  int64_t tile = tile_of(kt);
  // Register software pipeline over tiles: while tile t is processed, the gathered values of
  // tile t+stride and the dof ids of tile t+2*stride are already in flight (the id -> value
  // gather is a two-level dependent load chain; without this it is ~30% of the warp's time).
  Cell<E> c, n;
  int64_t tnext = (kt + stride) < ngen ? tile_of(kt + stride) : 0;
  {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    load_ids<E>(D, valid ? cell : 0, valid, c);
    const int64_t ncell = tnext * 32 + lane;
    const bool nvalid = (kt + stride) < ngen && ncell < D.ncells;
    load_ids<E>(D, nvalid ? ncell : 0, nvalid, n);
    griddep_wait();               // from here on the caller's vectors are touched
    griddep_launch_dependents();  // e.g. the zero fill of the other (disjoint) FE block starts right away
    gather_state<E>(D, x, valid, c);
  }
  double lam_e[E::NY], nlam[E::NY];
  if (vals_con != nullptr) gather_free<E::NY>(lambda, c.gy, lam_e);
  constexpr bool PFB = coord_prefetch_base<E, true>();
  int64_t b_h = PFB ? D.base_h[tile] : 0;
  while (true) {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    const bool has_next = (kt + stride) < ngen;
    const bool has_next2 = (kt + 2 * stride) < ngen;
    const int64_t tnext2 = has_next2 ? tile_of(kt + 2 * stride) : 0;
    Cell<E> nn;
    int64_t nb_h = 0;
    {
      gather_state<E>(D, x, has_next && (tnext * 32 + lane) < D.ncells, n);
      if (vals_con != nullptr) gather_free<E::NY>(lambda, n.gy, nlam);
      const int64_t c2 = tnext2 * 32 + lane;
      const bool v2 = has_next2 && c2 < D.ncells;
      load_ids<E>(D, v2 ? c2 : 0, v2, nn);
      if (PFB && has_next) nb_h = D.base_h[tnext];
    }
    int fy, fu, cnt;
    closed_counts<E>(c, fy, fu, cnt);
    const int incl = warp_incl_scan(cnt, lane);
    const int64_t tb = PFB ? b_h : D.base_h[tile];

    const int n11 = fy * (fy + 1) / 2;  // kept entries of block (1,1) (distinct ids)
    // One staged FE block of the tile.  jet_at(q, scale) returns the pointwise Jet2 whose Hessian
    // w.r.t. the FE fields generates the block.  Structural zeros are detected at run time, per
    // sub-block and warp-uniformly (all quadrature points, all lanes): a vanishing (y,y) part and/or
    // a vanishing u row skip their candidates; if both vanish the image is zero-filled.
    auto segment = [&](double* gdst, auto&& jet_at_w) {
      double wq = 0.0;  // quadrature weight * det(J) of the point last evaluated
      auto jet_at = [&](int q, double& scale) { return jet_at_w(q, scale, wq); };
      if (NBUF > 1) { st.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      st.template acquire<NBUF>(lane);
      double* const p0 = st.begin(gdst, incl - cnt, lane);
      // (elements that sum their blocks over the points in registers skip the test pass: no zero / uniformity tests)
      constexpr bool RA = coord_reg_accumulate<E>();
      bool zyy = !RA, zu = !RA, quni = true;
      Jet2<E::M> L, L0;
      double sc = 0.0;  // quadrature-weight-free scale (obj_weight or 1)
      const int NQT = RA ? 0 : ((KTab<E>::ENABLED && I::FE_DERIVS_POINT_INDEPENDENT && !D.notrait) ? 1 : E::NQ);
#pragma unroll 1
      for (int q = 0; q < NQT; ++q) {
        L = jet_at(q, sc);
        zyy = zyy && (sc == 0.0 || fe_hess_block_is_zero<E, false>(L));
        zu = zu && (sc == 0.0 || fe_hess_block_is_zero<E, true>(L));
        if (KTab<E>::ENABLED) {
          if (q == 0) L0 = L;
          else {
#pragma unroll
            for (int k = 0; k <= E::SU; ++k)
#pragma unroll
              for (int l = 0; l <= k; ++l) quni = quni && (L.h[Jet2<E::M>::idx(k, l)] == L0.h[Jet2<E::M>::idx(k, l)]);
          }
        }
      }
      zyy = __all_sync(0xffffffffu, zyy || !valid);
      zu = __all_sync(0xffffffffu, zu || !valid);
      const bool kpath = !RA && KTab<E>::ENABLED && Ks != nullptr && __all_sync(0xffffffffu, quni || !valid);
      if (zyy && zu) {
        st.zero_fill(__shfl_sync(0xffffffffu, incl, 31), lane);
      } else if (kpath) {  // q-independent second derivatives: pre-integrated tensors, one pass
        double* p = p0;
        if (!zyy) p = emit_hess_yy_K<E>(Ks, L0, sc, c, p);
        else { for (int k = 0; k < n11; ++k) p[k] = 0.0; p += n11; }
        if (!zu) emit_hess_u_K<E>(Ks, L0, sc, c, p);
        else for (int k = 0; k < cnt - n11; ++k) p[k] = 0.0;
      } else if (RA) {  // sum the element blocks over the points in registers, stage once
        using J2 = Jet2<E::M>;
        double Hyy[E::NY][E::NY], Huy[E::NY][E::NUS], Huu[E::NUS][E::NUS];
#pragma unroll
        for (int j = 0; j < E::NY; ++j) {
#pragma unroll
          for (int i = 0; i < E::NY; ++i) Hyy[j][i] = 0.0;
#pragma unroll
          for (int i = 0; i < E::NUS; ++i) Huy[j][i] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < E::NUS; ++j)
#pragma unroll
          for (int i = 0; i < E::NUS; ++i) Huu[j][i] = 0.0;
#pragma unroll 1
        for (int q = 0; q < E::NQ; ++q) {
          L = jet_at(q, sc);
          const double scale = sc * wq;
#pragma unroll
          for (int j = 0; j < E::NY; ++j) {
            double Nj[E::NB];
#pragma unroll
            for (int k = 0; k < E::NB; ++k) {
              double a = 0.0;
#pragma unroll
              for (int l = 0; l < E::NB; ++l) a = fma(L.h[J2::idx(k, l)], tab.By[q][j][l], a);
              Nj[k] = a * scale;
            }
#pragma unroll
            for (int i = 0; i < E::NY; ++i)
#pragma unroll
              for (int k = 0; k < E::NB; ++k) Hyy[j][i] = fma(tab.By[q][i][k], Nj[k], Hyy[j][i]);
            if (E::NU > 0) {
              double a = 0.0;
#pragma unroll
              for (int l = 0; l < E::NB; ++l) a = fma(L.h[J2::idx(E::SU, l)], tab.By[q][j][l], a);
              a *= scale;
#pragma unroll
              for (int i = 0; i < E::NU; ++i) Huy[j][i] = fma(tab.Bu[q][i], a, Huy[j][i]);
            }
          }
          if (E::NU > 0) {
            const double duu = L.h[J2::idx(E::SU, E::SU)] * scale;
#pragma unroll
            for (int j = 0; j < E::NU; ++j)
#pragma unroll
              for (int i = 0; i < E::NU; ++i) Huu[j][i] = fma(tab.Bu[q][i] * tab.Bu[q][j], duu, Huu[j][i]);
          }
        }
        double* p = p0;
#pragma unroll
        for (int j = 0; j < E::NY; ++j)
          if (c.gy[j] > 0) {
#pragma unroll
            for (int i = 0; i < E::NY; ++i)
              if (c.gy[j] <= c.gy[i]) { *p = Hyy[j][i]; ++p; }
          }
        if (E::NU > 0) {
#pragma unroll
          for (int j = 0; j < E::NY; ++j)
            if (c.gy[j] > 0) {
#pragma unroll
              for (int i = 0; i < E::NU; ++i)
                if (c.gu[i] > 0) { *p = Huy[j][i]; ++p; }
            }
#pragma unroll
          for (int j = 0; j < E::NU; ++j)
            if (c.gu[j] > 0) {
#pragma unroll
              for (int i = 0; i < E::NU; ++i)
                if (c.gu[j] <= c.gu[i]) { *p = Huu[j][i]; ++p; }
            }
        }
      } else {
#pragma unroll 1
        for (int q = 0; q < E::NQ; ++q) {
          if (E::NQ > 1) L = jet_at(q, sc);  // NQ == 1: the jet of the test pass is reused
          const double scale = sc * wq;
          double* p = p0;
          if (!zyy) p = (q == 0) ? emit_hess_yy<E, false>(tab, q, L, scale, c, p) : emit_hess_yy<E, true>(tab, q, L, scale, c, p);
          else { if (q == 0) for (int k = 0; k < n11; ++k) p[k] = 0.0; p += n11; }
          if (!zu) { if (q == 0) emit_hess_u<E, false>(tab, q, L, scale, c, p); else emit_hess_u<E, true>(tab, q, L, scale, c, p); }
          else if (q == 0) for (int k = 0; k < cnt - n11; ++k) p[k] = 0.0;
        }
      }
      st.release(gdst, incl - cnt, incl, lane);
    };
    if (vals_obj != nullptr) {
      segment(vals_obj + tb, [&](int q, double& scale, double& wq) {
        Jet2<E::M> s[E::M];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        scale = obj_weight;
        wq = pc.w;
        return I::f(pc, s);
      });
    }
    if (vals_con != nullptr) {
      segment(vals_con + tb, [&](int q, double& scale, double& wq) {
        Jet2<E::M> s[E::M];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        double Lam[E::NR];
        lambda_slots<E>(tab, q, lam_e, Lam);
        scale = 1.0;
        wq = pc.w;
        return weighted_res<E>(pc, s, Lam);
      });
    }
    if (!has_next) break;
    c = n;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) { n.gy[a] = nn.gy[a]; lam_e[a] = nlam[a]; }
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) n.gu[b] = nn.gu[b];
    b_h = nb_h;
    tile = tnext;
    tnext = tnext2;
    kt += stride;
  }
  st.drain(lane);
}


// ------------------------------------------------------------------ cell-granular templates
// The template-tile idea at the granularity of cells.  The image of a template TILE is 32 copies of the image of a
// template CELL (same values, same kept entries, cell after cell), so any run of n <= 32 consecutive template cells is
// the first n*RUN doubles of the tile image: pdenlp_create cuts the slab into CHUNKS (maximal runs of template cells,
// split every 32 cells, each with the offset of its first entry in every segment) and a compact list of the remaining
// GENERIC cells (boundary layers: 0.2 % of config 2, 4.6 % of config 5) with their offsets.
// Two launches, one CTA per unit of work, like the zero fill (the fastest write stream measured on this part is a
// non-persistent grid of plain 16 B stores — 7.4 TB/s against 6.4 TB/s for TMA bulk stores):
//   * k_copy_chunks: one chunk per CTA — it copies the first n*RUN doubles of the image (global scratch, L1 hits
//     after the first touch) to the chunk's run with 16 B stores; no dof ids, no x, no arithmetic, ~30 registers.
//   * k_jac_cells_cg / k_hess_cells_cg: 256 generic cells per CTA.  They are scattered, so a staged image would not
//     be one run.  Elements on the reference-tensor path evaluate lane = cell and then emit lane = ENTRY, cell after
//     cell (the cell's derivative data is broadcast by shuffles, the kept entries are ranked by popcounts /
//     ballots): the stores of a cell are contiguous across the lanes.  (Lane = cell stores cost one 32 B sector
//     request per 8 B entry: 55 us for the 97 k boundary cells of config 5.)  One-point elements emit lane = cell
//     straight to global memory (config 2: 0.2 % of the cells).  The host runs this launch, together with the
//     latency-bound kappa-block kernels, on the model's auxiliary stream, concurrently with the chunk copy.
// Same arithmetic as the tile kernels, entry by entry (same fma chains; terms the tile kernels skip are exact zeros).
// n <= 2*len doubles of the periodic image img[0..len) (len = 32 cells) to g[0..n)
__device__ __forceinline__ void copy_image(double* __restrict__ g, const double* __restrict__ img, int n, int len) {
  const int head = (int)((reinterpret_cast<uintptr_t>(g) >> 3) & 1);  // 8 B-aligned start
  const int n2 = (n - head) >> 1;
  double2* const q = reinterpret_cast<double2*>(g + head);
#pragma unroll 4
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    int j0 = 2 * i + head, j1 = j0 + 1;
    j0 -= j0 >= len ? len : 0;
    j1 -= j1 >= len ? len : 0;
    q[i] = make_double2(__ldg(img + j0), __ldg(img + j1));
  }
  if (threadIdx.x == 0) {
    if (head) g[0] = img[0];
    if ((n - head) & 1) { int j = n - 1; j -= j >= len ? len : 0; g[n - 1] = img[j]; }
  }
}
__device__ __forceinline__ double* shfl_ptr(double* p, int src) {
  return reinterpret_cast<double*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(p), src));
}
template <int N>
__device__ __forceinline__ unsigned free_mask(const int* g) {
  unsigned m = 0;
#pragma unroll
  for (int a = 0; a < N; ++a) m |= (g[a] > 0 ? 1u : 0u) << a;
  return m;
}
// lane = entry emission of the Jacobian blocks of the warp's 32 cells (lane = cell data in R, c, py, pu)
template <class E>
__device__ __forceinline__ void cg_emit_jac_K(const double* __restrict__ Ks, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, bool valid,
                                              double* py, double* pu, int lane) {
  using KT = KTab<E>;
  constexpr int NB = E::NB, NY = E::NY, NU = E::NU;
  unsigned nzy = 0, nzu = 0;
#pragma unroll
  for (int t = 0; t < NB; ++t) {
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (__any_sync(0xffffffffu, valid && R[t].g[k] != 0.0)) nzy |= 1u << (t * NB + k);
    if (NU > 0 && __any_sync(0xffffffffu, valid && R[t].g[E::SU] != 0.0)) nzu |= 1u << t;
  }
  const unsigned fmy = free_mask<NY>(c.gy), fmu = NU > 0 ? free_mask<E::NUS>(c.gu) : 0u;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll 1
  for (int r = 0; r < 32; ++r) {
    if (!((vmask >> r) & 1u)) continue;
    const unsigned my = __shfl_sync(0xffffffffu, fmy, r), mu = __shfl_sync(0xffffffffu, fmu, r);
    const int fy = __popc(my);
    double* const gy = shfl_ptr(py, r);
    double* const gu = shfl_ptr(pu, r);
    double dy[NB * NB], du[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
#pragma unroll
      for (int k = 0; k < NB; ++k) dy[t * NB + k] = ((nzy >> (t * NB + k)) & 1u) ? __shfl_sync(0xffffffffu, R[t].g[k], r) : 0.0;
      du[t] = (NU > 0 && ((nzu >> t) & 1u)) ? __shfl_sync(0xffffffffu, R[t].g[E::SU], r) : 0.0;
    }
#pragma unroll
    for (int e0 = 0; e0 < NY * NY; e0 += 32) {  // y-block: column j outer, row i inner
      const int e = e0 + lane;
      const bool in = e < NY * NY;
      const int j = in ? e / NY : 0, i = in ? e - j * NY : 0;
      double acc = 0.0;
#pragma unroll
      for (int t = 0; t < NB; ++t)
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if ((nzy >> (t * NB + k)) & 1u) acc = fma(dy[t * NB + k], Ks[KT::yy(t, k, 0, 0) + (in ? e : 0)], acc);
      if (in && ((my >> j) & 1u) && ((my >> i) & 1u))
        gy[__popc(my & ((1u << j) - 1u)) * fy + __popc(my & ((1u << i) - 1u))] = acc;
    }
    if (NU > 0) {
#pragma unroll
      for (int e0 = 0; e0 < NU * NY; e0 += 32) {  // u-block: column b outer, row i inner
        const int e = e0 + lane;
        const bool in = e < NU * NY;
        const int b = in ? e / NY : 0, i = in ? e - b * NY : 0;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < NB; ++t)
          if ((nzu >> t) & 1u) acc = fma(du[t], Ks[KT::yu(t, 0, 0) + (in ? e : 0)], acc);
        if (in && ((mu >> b) & 1u) && ((my >> i) & 1u))
          gu[__popc(mu & ((1u << b) - 1u)) * fy + __popc(my & ((1u << i) - 1u))] = acc;
      }
    }
  }
}
// lane = entry emission of the FE Hessian blocks (1,1), (2,1), (2,2) of the warp's 32 cells; ids = this warp's
// shared-memory copy of the cells' dof ids, [32][IDS] (the lower-triangle filter compares global ids)
template <class E>
__host__ __device__ constexpr int cg_ids_stride() { return (E::NY + E::NU) | 1; }
template <class E>
__device__ __forceinline__ void cg_emit_hess_K(const double* __restrict__ Ks, const Jet2<E::M>& L, double sc, bool zyy, bool zu,
                                               const int* ids, bool valid, double* p0, int lane) {
  using KT = KTab<E>;
  using J2 = Jet2<E::M>;
  constexpr int NB = E::NB, NY = E::NY, NU = E::NU, IDS = cg_ids_stride<E>();
  unsigned nzy = 0, nzu = 0;
#pragma unroll
  for (int k = 0; k < NB; ++k)
#pragma unroll
    for (int l = 0; l < NB; ++l)
      if (__any_sync(0xffffffffu, valid && L.h[J2::idx(k, l)] != 0.0)) nzy |= 1u << (k * NB + l);
  if (NU > 0) {
#pragma unroll
    for (int l = 0; l < NB; ++l)
      if (__any_sync(0xffffffffu, valid && L.h[J2::idx(E::SU, l)] != 0.0)) nzu |= 1u << l;
  }
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
  for (int r = 0; r < 32; ++r) {
    if (!((vmask >> r) & 1u)) continue;
    double* const gp = shfl_ptr(p0, r);
    const int* const id = ids + r * IDS;
    int base = 0;
    {  // block (1,1): column j outer, row i inner, keep gid_col > 0 and gid_col <= gid_row
      double d[NB * NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
#pragma unroll
        for (int l = 0; l < NB; ++l)
          d[k * NB + l] = ((nzy >> (k * NB + l)) & 1u) ? __shfl_sync(0xffffffffu, L.h[J2::idx(k, l)] * sc, r) : 0.0;
#pragma unroll
      for (int e0 = 0; e0 < NY * NY; e0 += 32) {
        const int e = e0 + lane;
        const bool in = e < NY * NY;
        const int j = in ? e / NY : 0, i = in ? e - j * NY : 0;
        const int gj = id[j], gi = id[i];
        const bool kept = in && gj > 0 && gj <= gi;
        double acc = 0.0;
        if (!zyy) {
#pragma unroll
          for (int k = 0; k < NB; ++k)
#pragma unroll
            for (int l = 0; l < NB; ++l)
              if ((nzy >> (k * NB + l)) & 1u) acc = fma(d[k * NB + l], Ks[KT::yy(k, l, 0, 0) + (in ? e : 0)], acc);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (kept) gp[base + __popc(bal & lt)] = acc;
        base += __popc(bal);
      }
    }
    if (NU > 0) {
      double dl[NB];
#pragma unroll
      for (int l = 0; l < NB; ++l) dl[l] = ((nzu >> l) & 1u) ? __shfl_sync(0xffffffffu, L.h[J2::idx(E::SU, l)] * sc, r) : 0.0;
      const double duu = __shfl_sync(0xffffffffu, L.h[J2::idx(E::SU, E::SU)] * sc, r);
#pragma unroll
      for (int e0 = 0; e0 < NY * NU; e0 += 32) {  // block (2,1): state column j outer, control row i inner
        const int e = e0 + lane;
        const bool in = e < NY * NU;
        const int j = in ? e / NU : 0, i = in ? e - j * NU : 0;
        const bool kept = in && id[j] > 0 && id[NY + i] > 0;
        double acc = 0.0;
        if (!zu) {
#pragma unroll
          for (int l = 0; l < NB; ++l)
            if ((nzu >> l) & 1u) acc = fma(dl[l], Ks[KT::uy(l, 0, 0) + (in ? e : 0)], acc);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (kept) gp[base + __popc(bal & lt)] = acc;
        base += __popc(bal);
      }
#pragma unroll
      for (int e0 = 0; e0 < NU * NU; e0 += 32) {  // block (2,2)
        const int e = e0 + lane;
        const bool in = e < NU * NU;
        const int j = in ? e / NU : 0, i = in ? e - j * NU : 0;
        const int gj = id[NY + j], gi = id[NY + i];
        const bool kept = in && gj > 0 && gj <= gi;
        const double v = zu ? 0.0 : duu * Ks[KT::uu(0, 0) + (in ? e : 0)];
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (kept) gp[base + __popc(bal & lt)] = v;
        base += __popc(bal);
      }
    }
  }
}

// one CTA per chunk: copies the first n*RUN doubles of the image(s) to the chunk's run(s) (element-independent).
// which = 0: Jacobian (segment a = y-block, b = u-block when b.vals != nullptr); 1: FE Hessian block (segment a)
struct ChunkSeg {
  double* vals;         // segment of the output (nullptr: unused)
  const double* img;    // image of a template tile
  int run;              // entries per template cell
};
// chained = 1: launched with programmatic stream serialization right behind the generic-cell kernel of the same callback
// (same stream, disjoint outputs).  That kernel started only after the image was complete, so this one need not wait
// for it before copying; ONE thread waits at the end, so that the completion of this grid implies the completion of
// its predecessor (the k_zero_fill protocol).  Both run concurrently without an event fork / join.
// CTAs [nchunk, gridDim.x) zero-fill fillp[0, filln) — the structurally zero constraint FE block of hess_coord! — one
// 16 KB page each, like k_zero_fill: one launch for both streams of the callback.
static __global__ void __launch_bounds__(kThreads) k_copy_chunks(ChunkSeg a, ChunkSeg b, const ChunkRec* __restrict__ ch, int which, int chained,
                                                                 int nchunk, double* __restrict__ fillp, int64_t filln) {
  if (!chained) griddep_wait();  // (the image is written by the preceding one-tile launch on the stream)
  griddep_launch_dependents();
  if ((int)blockIdx.x >= nchunk) {
    const int64_t head = (filln > 0 && ((reinterpret_cast<uintptr_t>(fillp) >> 3) & 1)) ? 1 : 0;
    const int64_t n2 = (filln - head) >> 1;
    double2* q = reinterpret_cast<double2*>(fillp + head);
    const double2 z = make_double2(0.0, 0.0);
    const int64_t c = (int64_t)blockIdx.x - nchunk;
    for (int64_t i = (c << 10) + threadIdx.x; i < min(n2, (c + 1) << 10); i += blockDim.x) q[i] = z;
    if (c == 0 && threadIdx.x == 0) {
      if (head) fillp[0] = 0.0;
      if ((filln - head) & 1) fillp[filln - 1] = 0.0;
    }
    return;
  }
  // the record is two 16 B loads of one sector: both segment offsets are known before the first store
  const longlong2 r0 = __ldg(reinterpret_cast<const longlong2*>(ch + blockIdx.x));
  const longlong2 r1 = __ldg(reinterpret_cast<const longlong2*>(ch + blockIdx.x) + 1);
  const int n = (int)r1.y;
  copy_image(a.vals + (which ? r1.x : r0.x), a.img, n * a.run, 32 * a.run);
  if (b.vals != nullptr) copy_image(b.vals + r0.y, b.img, n * b.run, 32 * b.run);
  if (chained && blockIdx.x == 0 && threadIdx.x == 0) griddep_wait();
}

// generic cells of the Jacobian blocks: one warp per 32 cells of the list
template <class E>
__global__ void __launch_bounds__(kThreads)
k_jac_cells_cg(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
               double* __restrict__ vals_y, double* __restrict__ vals_u) {
  using I = typename E::Integrand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_wait();
  griddep_launch_dependents();
  const int64_t i = ((int64_t)blockIdx.x * kWarpsPerCta + warp) * 32 + lane;
  const bool valid = i < D.ngc;
  const int64_t cell = valid ? (int64_t)__ldg(D.gc_cell + i) : 0;
  double* const py = vals_y + (valid ? __ldg(D.gc_jy + i) : 0);
  double* const pu = vals_u + ((valid && E::NU > 0) ? __ldg(D.gc_ju + i) : 0);
  Cell<E> c;
  load_ids<E>(D, cell, valid, c);
  gather_state<E>(D, x, valid, c);
  Jet1<E::M> s[E::M], R[E::NR];
  const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, 0, I::NEEDS_COEF_DERIV);
  point_state<E>(tab, pc, c, x, 0, s);
  eval_res<E>(pc, s, R);
  if constexpr (E::NQ > 1) {  // point-independent derivatives: pre-integrated reference tensors
    cg_emit_jac_K<E>(D.ktab, R, c, valid, py, pu, lane);
  } else {
    emit_jac_y_pt<E>(tab, 0, pc.w, R, c, py);
    if (E::NU > 0) emit_jac_u_pt<E>(tab, 0, pc.w, R, c, pu);
  }
}

// generic cells of the objective FE Hessian block (the one computed block: tpl_ok<E, true> integrands have a
// structurally zero constraint block)
template <class E>
__global__ void __launch_bounds__(kThreads)
k_hess_cells_cg(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x, double obj_weight,
                double* __restrict__ vals_obj) {
  using I = typename E::Integrand;
  __shared__ int sids[E::NQ > 1 ? kWarpsPerCta * 32 * cg_ids_stride<E>() : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_wait();
  griddep_launch_dependents();
  const int64_t i = ((int64_t)blockIdx.x * kWarpsPerCta + warp) * 32 + lane;
  const bool valid = i < D.ngc;
  const int64_t cell = valid ? (int64_t)__ldg(D.gc_cell + i) : 0;
  double* const p0 = vals_obj + (valid ? __ldg(D.gc_h + i) : 0);
  Cell<E> c;
  load_ids<E>(D, cell, valid, c);
  gather_state<E>(D, x, valid, c);
  int fy, fu, cnt;
  closed_counts<E>(c, fy, fu, cnt);
  const int n11 = fy * (fy + 1) / 2;
  Jet2<E::M> s[E::M];
  const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, 0, I::NEEDS_COEF_DERIV);
  point_state<E>(tab, pc, c, x, 0, s);
  const Jet2<E::M> L = I::f(pc, s);
  const double sc = obj_weight;
  // structural zeros, as in the tile kernel (warp-uniform; a vanishing part is written as zeros)
  const bool zyy = __all_sync(0xffffffffu, sc == 0.0 || fe_hess_block_is_zero<E, false>(L) || !valid);
  const bool zu = __all_sync(0xffffffffu, sc == 0.0 || fe_hess_block_is_zero<E, true>(L) || !valid);
  if constexpr (E::NQ > 1) {
    int* const ids = sids + warp * 32 * cg_ids_stride<E>();
#pragma unroll
    for (int a = 0; a < E::NY; ++a) ids[lane * cg_ids_stride<E>() + a] = c.gy[a];
#pragma unroll
    for (int b = 0; b < E::NU; ++b) ids[lane * cg_ids_stride<E>() + E::NY + b] = c.gu[b];
    __syncwarp();
    cg_emit_hess_K<E>(D.ktab, L, sc, zyy, zu, ids, valid, p0, lane);
  } else {
    double* p = p0;
    const double scale = sc * pc.w;
    if (!zyy) p = emit_hess_yy<E, false>(tab, 0, L, scale, c, p);
    else { for (int k = 0; k < n11; ++k) p[k] = 0.0; p += n11; }
    if (!zu) emit_hess_u<E, false>(tab, 0, L, scale, c, p);
    else for (int k = 0; k < cnt - n11; ++k) p[k] = 0.0;
  }
}


// ------------------------------------------------------------------ multi-field coordinate kernels
// Several state / control fields (MultiFieldFESpace): the same lane = cell, warp = tile streaming
// writer, with the entries of a cell emitted block by block (column field outer, row field inner —
// fill_matrix_hessian_functions.jl:50-94), and inside a block column dof outer / row dof inner.  The
// per-tile segment offsets are those of the single-field kernels (counts do not depend on the order).
// These are the ODE-type problems of the reference (1-D, one quadrature point); they take the plain
// per-tile loop without the register pipeline and reference-tensor paths of the kernels above.
template <class E, bool ACC>
__device__ __forceinline__ double* emit_jac_y_mf(const Tables<E>& tab, int q, double w, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, double* p) {
#pragma unroll
  for (int cf = 0; cf < E::NFY; ++cf)
#pragma unroll
    for (int rf = 0; rf < E::NFY; ++rf)
#pragma unroll
      for (int j = 0; j < E::NLY; ++j) {
        double Mj[E::NB];
#pragma unroll
        for (int t = 0; t < E::NB; ++t) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < E::NB; ++k) a = fma(R[rf * E::NB + t].g[cf * E::NB + k], tab.By[q][j][k], a);
          Mj[t] = a * w;
        }
        if (c.gy[cf * E::NLY + j] > 0) {
#pragma unroll
          for (int i = 0; i < E::NLY; ++i) {
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < E::NB; ++t) v = fma(tab.By[q][i][t], Mj[t], v);
            if (c.gy[rf * E::NLY + i] > 0) { if (ACC) *p += v; else *p = v; ++p; }
          }
        }
      }
  return p;
}
template <class E, bool ACC>
__device__ __forceinline__ double* emit_jac_u_mf(const Tables<E>& tab, int q, double w, const Jet1<E::M> (&R)[E::NR], const Cell<E>& c, double* p) {
#pragma unroll
  for (int cf = 0; cf < E::NFU; ++cf)
#pragma unroll
    for (int rf = 0; rf < E::NFY; ++rf)
#pragma unroll
      for (int b = 0; b < E::NLU; ++b) {
        double Mb[E::NB];
#pragma unroll
        for (int t = 0; t < E::NB; ++t) Mb[t] = R[rf * E::NB + t].g[E::SU + cf] * tab.Bu[q][b] * w;
        if (c.gu[cf * E::NLU + b] > 0) {
#pragma unroll
          for (int i = 0; i < E::NLY; ++i) {
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < E::NB; ++t) v = fma(tab.By[q][i][t], Mb[t], v);
            if (c.gy[rf * E::NLY + i] > 0) { if (ACC) *p += v; else *p = v; ++p; }
          }
        }
      }
  return p;
}

template <class E>
__global__ void __launch_bounds__(kThreads)
k_jac_coord_mf(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
               double* __restrict__ vals_y, double* __restrict__ vals_u) {
  using I = typename E::Integrand;
  static_assert(!E::HAS_CONST, "multi-field integrands have no constant residual slot");
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int STAGE = jac_stage_doubles<E>(), NBUF = coord_nbuf(STAGE);
  double* const wbase = smem + warp * STAGE * NBUF;
  int slot = 0;
  SegStage<E::RUN_JY> sty{wbase, nullptr};
  SegStage<(E::NU > 0 ? E::RUN_JU : 1)> stu{wbase, nullptr};
  const int64_t stride = (int64_t)gridDim.x * nwarps;
  int64_t tile = (int64_t)blockIdx.x * nwarps + warp;
  if (tile >= D.ntiles) return;
  // the same register pipeline as k_jac_coord: state of tile t+stride, ids of tile t+2*stride and the segment
  // offsets of tile t+stride are in flight while tile t is processed
  Cell<E> c, n;
  {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    load_ids<E>(D, valid ? cell : 0, valid, c);
    gather_state<E>(D, x, valid, c);
    const int64_t ncell = (tile + stride) * 32 + lane;
    const bool nvalid = (tile + stride) < D.ntiles && ncell < D.ncells;
    load_ids<E>(D, nvalid ? ncell : 0, nvalid, n);
  }
  int64_t b_jy = D.base_jy[tile], b_ju = E::NU > 0 ? D.base_ju[tile] : 0;
  while (true) {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    const int64_t tnext = tile + stride;
    const bool has_next = tnext < D.ntiles;
    Cell<E> nn;
    int64_t nb_jy = 0, nb_ju = 0;
    {
      gather_state<E>(D, x, has_next && (tnext * 32 + lane) < D.ncells, n);
      const int64_t c2 = (tnext + stride) * 32 + lane;
      const bool v2 = (tnext + stride) < D.ntiles && c2 < D.ncells;
      load_ids<E>(D, v2 ? c2 : 0, v2, nn);
      if (has_next) { nb_jy = D.base_jy[tnext]; if (E::NU > 0) nb_ju = D.base_ju[tnext]; }
    }
    int fy = 0, fu = 0;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) fy += c.gy[a] > 0;
#pragma unroll
    for (int b = 0; b < E::NU; ++b) fu += c.gu[b] > 0;
    {
      const int cnt = fy * fy;
      const int incl = warp_incl_scan(cnt, lane);
      double* gdst = vals_y + b_jy;
      if (NBUF > 1) { sty.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      sty.template acquire<NBUF>(lane);
      double* const p0 = sty.begin(gdst, incl - cnt, lane);
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet1<E::M> s[E::M], R[E::NR];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
        if (q == 0) emit_jac_y_mf<E, false>(tab, q, pc.w, R, c, p0); else emit_jac_y_mf<E, true>(tab, q, pc.w, R, c, p0);
      }
      sty.release(gdst, incl - cnt, incl, lane);
    }
    if (E::NU > 0) {
      const int cnt = fy * fu;
      const int incl = warp_incl_scan(cnt, lane);
      double* gdst = vals_u + b_ju;
      if (NBUF > 1) { stu.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      stu.template acquire<NBUF>(lane);
      double* const p0 = stu.begin(gdst, incl - cnt, lane);
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet1<E::M> s[E::M], R[E::NR];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
        if (q == 0) emit_jac_u_mf<E, false>(tab, q, pc.w, R, c, p0); else emit_jac_u_mf<E, true>(tab, q, pc.w, R, c, p0);
      }
      stu.release(gdst, incl - cnt, incl, lane);
    }
    if (!has_next) break;
    c = n;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) n.gy[a] = nn.gy[a];
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) n.gu[b] = nn.gu[b];
    b_jy = nb_jy; b_ju = nb_ju;
    tile = tnext;
  }
  sty.drain(lane);
}

// One FE Hessian block set of a cell over the fields F = (state fields, control fields): blocks with
// row field < column field hold no lower-triangle entry (multi-field ids grow with the field), the
// diagonal blocks keep gidcol <= gidrow, the blocks below keep every free pair.
template <class E, bool ACC>
__device__ __forceinline__ double* emit_hess_mf(const Tables<E>& tab, int q, const Jet2<E::M>& L, double scale,
                                                const Cell<E>& c, double* p) {
  using J2 = Jet2<E::M>;
  constexpr int NF = E::NFY + E::NFU;
#pragma unroll
  for (int cf = 0; cf < NF; ++cf)
#pragma unroll
    for (int rf = cf; rf < NF; ++rf) {
      if (cf < E::NFY) {
#pragma unroll
        for (int j = 0; j < E::NLY; ++j) {
          const int gc = c.gy[cf * E::NLY + j];
          if (rf < E::NFY) {  // state rows, state columns
            double Nj[E::NB];
#pragma unroll
            for (int k = 0; k < E::NB; ++k) {
              double a = 0.0;
#pragma unroll
              for (int l = 0; l < E::NB; ++l) a = fma(L.h[J2::idx(rf * E::NB + k, cf * E::NB + l)], tab.By[q][j][l], a);
              Nj[k] = a * scale;
            }
            if (gc > 0) {
#pragma unroll
              for (int i = 0; i < E::NLY; ++i) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < E::NB; ++k) v = fma(tab.By[q][i][k], Nj[k], v);
                const int gr = c.gy[rf * E::NLY + i];
                if (rf == cf ? gc <= gr : gr > 0) { if (ACC) *p += v; else *p = v; ++p; }
              }
            }
          } else {            // control rows, state columns
            const int ru = rf - E::NFY;
            double a = 0.0;
#pragma unroll
            for (int l = 0; l < E::NB; ++l) a = fma(L.h[J2::idx(E::SU + ru, cf * E::NB + l)], tab.By[q][j][l], a);
            a *= scale;
            if (gc > 0) {
#pragma unroll
              for (int i = 0; i < E::NLU; ++i) {
                const double v = tab.Bu[q][i] * a;
                if (c.gu[ru * E::NLU + i] > 0) { if (ACC) *p += v; else *p = v; ++p; }
              }
            }
          }
        }
      } else {                // control rows, control columns
        const int cu = cf - E::NFY, ru = rf - E::NFY;
        const double duu = L.h[J2::idx(E::SU + ru, E::SU + cu)] * scale;
#pragma unroll
        for (int j = 0; j < E::NLU; ++j) {
          const double a = duu * tab.Bu[q][j];
          const int gc = c.gu[cu * E::NLU + j];
          if (gc > 0) {
#pragma unroll
            for (int i = 0; i < E::NLU; ++i) {
              const double v = tab.Bu[q][i] * a;
              const int gr = c.gu[ru * E::NLU + i];
              if (ru == cu ? gc <= gr : gr > 0) { if (ACC) *p += v; else *p = v; ++p; }
            }
          }
        }
      }
    }
  return p;
}

template <class E>
__global__ void __launch_bounds__(kThreads)
k_hess_coord_mf(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
                const double* __restrict__ lambda, double obj_weight,
                double* __restrict__ vals_obj, double* __restrict__ vals_con) {
  using I = typename E::Integrand;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int STAGE = hess_stage_doubles<E>(), NBUF = coord_nbuf(STAGE);
  double* const wbase = smem + warp * STAGE * NBUF;
  int slot = 0;
  SegStage<E::RUN_H> st{wbase, nullptr};
  const int64_t stride = (int64_t)gridDim.x * nwarps;
  int64_t tile = (int64_t)blockIdx.x * nwarps + warp;
  if (tile >= D.ntiles) return;
  Cell<E> c, n;  // register pipeline over tiles, as in k_hess_coord
  {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    load_ids<E>(D, valid ? cell : 0, valid, c);
    gather_state<E>(D, x, valid, c);
    const int64_t ncell = (tile + stride) * 32 + lane;
    const bool nvalid = (tile + stride) < D.ntiles && ncell < D.ncells;
    load_ids<E>(D, nvalid ? ncell : 0, nvalid, n);
  }
  int64_t b_h = D.base_h[tile];
  while (true) {
    const int64_t cell = tile * 32 + lane;
    const bool valid = cell < D.ncells;
    const int64_t tnext = tile + stride;
    const bool has_next = tnext < D.ntiles;
    Cell<E> nn;
    int64_t nb_h = 0;
    {
      gather_state<E>(D, x, has_next && (tnext * 32 + lane) < D.ncells, n);
      const int64_t c2 = (tnext + stride) * 32 + lane;
      const bool v2 = (tnext + stride) < D.ntiles && c2 < D.ncells;
      load_ids<E>(D, v2 ? c2 : 0, v2, nn);
      if (has_next) nb_h = D.base_h[tnext];
    }
    int fy, fu, cnt;
    closed_counts<E>(c, fy, fu, cnt);
    const int incl = warp_incl_scan(cnt, lane);
    const int64_t tb = b_h;
    if (vals_obj != nullptr) {
      double* gdst = vals_obj + tb;
      if (NBUF > 1) { st.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      st.template acquire<NBUF>(lane);
      double* const p0 = st.begin(gdst, incl - cnt, lane);
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet2<E::M> s[E::M];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        const Jet2<E::M> L = I::f(pc, s);
        const double scale = obj_weight * pc.w;
        if (q == 0) emit_hess_mf<E, false>(tab, q, L, scale, c, p0); else emit_hess_mf<E, true>(tab, q, L, scale, c, p0);
      }
      st.release(gdst, incl - cnt, incl, lane);
    }
    if (vals_con != nullptr) {
      double lam_e[E::NY];
      gather_free<E::NY>(lambda, c.gy, lam_e);
      double* gdst = vals_con + tb;
      if (NBUF > 1) { st.wbuf = wbase + slot * STAGE; slot = slot + 1 == NBUF ? 0 : slot + 1; }
      st.template acquire<NBUF>(lane);
      double* const p0 = st.begin(gdst, incl - cnt, lane);
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet2<E::M> s[E::M];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, I::NEEDS_COEF_DERIV);
        point_state<E>(tab, pc, c, x, q, s);
        double Lam[E::NR];
        lambda_slots<E>(tab, q, lam_e, Lam);
        const Jet2<E::M> L = weighted_res<E>(pc, s, Lam);
        if (q == 0) emit_hess_mf<E, false>(tab, q, L, pc.w, c, p0); else emit_hess_mf<E, true>(tab, q, L, pc.w, c, p0);
      }
      st.release(gdst, incl - cnt, incl, lane);
    }
    if (!has_next) break;
    c = n;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) n.gy[a] = nn.gy[a];
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) n.gu[b] = nn.gu[b];
    b_h = nb_h;
    tile = tnext;
  }
  st.drain(lane);
}

// ------------------------------------------------------------------ structure (setup-time, not hot)
// rows/cols of the FE blocks; kind 0 = Jacobian y-block, 1 = Jacobian u-block, 2 = Hessian FE block.
// shift_r / shift_c are added to the 1-based ids (nparam and multi-field offsets).
template <class E>
__global__ void __launch_bounds__(kThreads)
k_structure(const DevModel D, int kind, int64_t* __restrict__ rows, int64_t* __restrict__ cols, int64_t shift) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= D.ntiles) return;
  const int64_t cell = tile * 32 + lane;
  const bool valid = cell < D.ncells;
  Cell<E> c;
  load_ids<E>(D, valid ? cell : 0, valid, c);
  int njy, nju, nh;
  count_entries<E>(c, njy, nju, nh);
  const int cnt = kind == 0 ? njy : (kind == 1 ? nju : nh);
  const int incl = warp_incl_scan(cnt, lane);
  const int64_t base = (kind == 0 ? D.base_jy[tile] : (kind == 1 ? D.base_ju[tile] : D.base_h[tile])) + (incl - cnt);
  int64_t n = base;
  const int64_t offu = D.nfree_y;
  // blocks in eachblockid order (column field outer, row field inner), inside a block column dof
  // outer / row dof inner; with one field per space this is the plain (j, i) order
  constexpr int NF = E::NFY + E::NFU;
  auto gid = [&](int F, int a) -> int64_t {  // multi-field global id of local dof a of field F (<= 0: not free)
    if (F < E::NFY) return c.gy[F * E::NLY + a];
    const int g = c.gu[(F - E::NFY) * E::NLU + a];
    return g > 0 ? g + offu : g;
  };
  auto nloc = [&](int F) { return F < E::NFY ? E::NLY : E::NLU; };
  const int c0 = kind == 1 ? E::NFY : 0, c1 = kind == 0 ? E::NFY : NF;
  const int r1 = kind == 2 ? NF : E::NFY;
  const int64_t shift_r = kind == 2 ? shift : 0;
  for (int cf = c0; cf < c1; ++cf)
    for (int rf = 0; rf < r1; ++rf)
      for (int j = 0; j < nloc(cf); ++j) {
        const int64_t gc = gid(cf, j);
        if (gc <= 0) continue;
        for (int i = 0; i < nloc(rf); ++i) {
          const int64_t gr = gid(rf, i);
          if (gr > 0 && (kind != 2 || gc <= gr)) { rows[n] = gr + shift_r; cols[n] = gc + shift; ++n; }
        }
      }
}

// The kappa entries of the dof-indexed outputs and the kappa-kappa corners of the Hessian blocks are sums over ALL
// cells of the slab.  Each CTA writes ONE partial per component (warp shuffles + a fixed-order sum over its warps) and a
// one-CTA second pass adds the partials in block order: no atomics, so these entries are bitwise reproducible run to
// run (the grid of a kernel is a function of the model only) and accurate like a pairwise sum.
struct KappaDst {
  int64_t pos[16];  // destination index of each component in the output array
};
static __global__ void k_add_partials(const double* __restrict__ partials, int nblocks, int ncomp, KappaDst dst, double* __restrict__ out) {
  __shared__ double sh[256];
  for (int c = 0; c < ncomp; ++c) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) a += partials[(int64_t)i * ncomp + c];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
      if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[dst.pos[c]] += sh[0];
    __syncthreads();
  }
}
// per-CTA sum of one value per thread -> partials[blockIdx.x * ncomp + comp]   (all threads of the CTA must call)
__device__ __forceinline__ void cta_partial(double v, double* __restrict__ partials, int ncomp, int comp, double* red /* [kWarpsPerCta] shared */) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
    partials[(int64_t)blockIdx.x * ncomp + comp] = a;
  }
  __syncthreads();
}

// Second pass without a second launch: every CTA publishes its partials, takes a ticket, and the CTA that arrives last
// adds all partials in BLOCK ORDER (the same fixed-order strided + tree sum whichever CTA it is), so the result stays
// bitwise reproducible run to run.  Saves one launch (~5 us + its ramp) per callback: visible for the 20-50 us callbacks
// of the small problems.  All threads of every CTA must call last_cta() exactly once, after their cta_partial() calls.
__device__ __forceinline__ bool last_cta(unsigned* ticket) {
  __shared__ unsigned s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // this CTA's partials are visible before its ticket
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (s_last) { *ticket = 0u; __threadfence(); }  // (every CTA has arrived: reset for the next launch on the stream)
  }
  __syncthreads();
  return s_last != 0u;
}
// sum of p[0], p[stride], ..., n terms, by the whole CTA (blockDim.x a power of two <= 256), result on every thread
__device__ __forceinline__ double cta_sum_strided(const double* p, int n, int stride) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += __ldcg(p + (int64_t)i * stride);
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int d = blockDim.x >> 1; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}
// position of component `comp` of the kappa-kappa corner (column j outer, row i >= j inner) in a trapezoid with
// `prows` rows; prows < 0: identity (the kappa entries of a dof-indexed vector)
__device__ __forceinline__ int64_t kk_pos(int comp, int np, int64_t prows) {
  if (prows < 0) return comp;
  int c = 0;
  int64_t col0 = 0;
  for (int j = 0; j < np; ++j) {
    for (int i = j; i < np; ++i, ++c)
      if (c == comp) return col0 + (i - j);
    col0 += prows - j;
  }
  return 0;
}
// called by the last CTA: out[pos(c)] += sum over the CTAs of partial component c
__device__ __forceinline__ void finish_partials(const double* kpart /* [gridDim.x][ncomp] */, int ncomp, int np, int64_t prows, double* out) {
  for (int c = 0; c < ncomp; ++c) {
    const double a = cta_sum_strided(kpart + c, (int)gridDim.x, ncomp);
    if (threadIdx.x == 0) out[kk_pos(c, np, prows)] += a;
  }
}
__device__ __forceinline__ void finish_objective(const DevModel& D, const double* fpart /* [gridDim.x] */) {
  const double a = cta_sum_strided(fpart, (int)gridDim.x, 1);
  if (threadIdx.x == 0) {
    *D.fsum = a;
    if (D.fsum_host != nullptr) *D.fsum_host = a;
  }
}

// ------------------------------------------------------------------ vector-valued callbacks
// (lane = cell; scatter with red.global.add.f64 — north star: "segmented atomics")
enum VecOp { OP_OBJ = 0, OP_GRAD = 1, OP_CONS = 2, OP_JPROD = 3, OP_JTPROD = 4, OP_HPROD = 5,
             OP_OBJGRAD = 6 /* objgrad!: objective and gradient from one sweep (one gather, one jet evaluation) */,
             OP_OBJCONS = 7 /* objcons!: objective and residual from one sweep */,
             OP_JAC_KAPPA = 8 /* dense kappa-block of jac_coord! */, OP_HESS_KAPPA = 9 /* kappa trapezoid of hess_coord! */,
             OP_HESS_KAPPA_KK = 10 /* objective trapezoid of an `inde` energy: the p x p corner only */ };
// number of per-CTA partial components (kappa entries) an op produces; they live at partials + gridDim.x * 1 (the first
// gridDim.x doubles hold the objective partials of OBJ / OBJGRAD / OBJCONS)
template <class E>
__host__ __device__ constexpr int kappa_partials(int op) {
  return E::NP == 0 ? 0
       : (op == OP_GRAD || op == OP_OBJGRAD || op == OP_JTPROD || op == OP_HPROD) ? E::NP
       : (op == OP_HESS_KAPPA || op == OP_HESS_KAPPA_KK) ? E::NP * (E::NP + 1) / 2 : 0;
}

// element vector d/dz_a += sum_k By_a[k] gs[k], d/du_b += Bu_b gs[SU]: accumulated over the quadrature points in
// registers, scattered once per cell (one red.global.add.f64 per local dof, not per dof and point)
template <class E>
__device__ __forceinline__ void accumulate_fields(const Tables<E>& tab, int q, const double* gs /*[M], weighted*/,
                                                  double* ey, double* eu) {
#pragma unroll
  for (int f = 0; f < E::NFY; ++f)
#pragma unroll
    for (int a = 0; a < E::NLY; ++a) {
      double v = ey[f * E::NLY + a];
#pragma unroll
      for (int k = 0; k < E::NB; ++k) v = fma(tab.By[q][a][k], gs[f * E::NB + k], v);
      ey[f * E::NLY + a] = v;
    }
#pragma unroll
  for (int f = 0; f < E::NFU; ++f)
#pragma unroll
    for (int b = 0; b < E::NLU; ++b) eu[f * E::NLU + b] = fma(tab.Bu[q][b], gs[E::SU + f], eu[f * E::NLU + b]);
}
template <class E>
__device__ __forceinline__ void scatter_element(const DevModel& D, const Cell<E>& c, const double* ey, const double* eu,
                                                double* __restrict__ out /* nvar-vector */) {
  double* oy = out + D.nparam;
  double* ou = oy + D.nfree_y;
#pragma unroll
  for (int a = 0; a < E::NY; ++a)
    if (c.gy[a] > 0) accumulate(D, oy + (c.gy[a] - 1), ey[a]);
  if (E::NU > 0) {
#pragma unroll
    for (int b = 0; b < E::NU; ++b)
      if (c.gu[b] > 0) accumulate(D, ou + (c.gu[b] - 1), eu[b]);
  }
}

// direction of the pointwise fields induced by a free-dof direction v (Dirichlet entries = 0)
template <class E>
__device__ __forceinline__ void field_direction(const Tables<E>& tab, int q, const DevModel& D, const Cell<E>& c,
                                                const double* __restrict__ v, double* ds) {
  double vy[E::NY], vu[E::NUS];
  gather_free<E::NY>(v + D.nparam, c.gy, vy);
  if (E::NU > 0) gather_free<E::NUS>(v + D.nparam + D.nfree_y, c.gu, vu);
#pragma unroll
  for (int k = 0; k < E::M; ++k) ds[k] = 0.0;
#pragma unroll
  for (int f = 0; f < E::NFY; ++f)
#pragma unroll
    for (int a = 0; a < E::NLY; ++a)
#pragma unroll
      for (int k = 0; k < E::NB; ++k) ds[f * E::NB + k] = fma(tab.By[q][a][k], vy[f * E::NLY + a], ds[f * E::NB + k]);
#pragma unroll
  for (int f = 0; f < E::NFU; ++f)
#pragma unroll
    for (int b = 0; b < E::NLU; ++b) ds[E::SU + f] = fma(tab.Bu[q][b], vu[f * E::NLU + b], ds[E::SU + f]);
#pragma unroll
  for (int k = 0; k < E::NP; ++k) ds[E::SK + k] = __ldg(v + k);
}

// These kernels are load-latency bound (ncu: long-scoreboard 8-14 stall cycles per issue at 16-24
// resident warps), so the register budget is capped to keep 3-4 CTAs (24-32 warps) per SM.
// hprod! over >= 6 pointwise fields (Jet2<7>: 28 second derivatives per jet) spills at 128 registers; with the
// whole register file (1 CTA/SM) config 5 runs in 0.50 ms instead of 0.93 ms.
#ifndef PDENLP_VEC_MINB_BIG_HPROD
#define PDENLP_VEC_MINB_BIG_HPROD 1
#endif
template <class E, int OP>
constexpr int vec_min_blocks() {
  if (E::M >= 6 && OP == OP_HPROD) return PDENLP_VEC_MINB_BIG_HPROD;
#ifndef PDENLP_VEC_MINB_SMALL
#define PDENLP_VEC_MINB_SMALL 4
#endif
  return (E::M >= 6 || OP == OP_HPROD) ? 2 : ((OP == OP_JPROD || OP == OP_JTPROD) ? 3 : (OP == OP_OBJ ? 4 : PDENLP_VEC_MINB_SMALL));
}
template <class E, int OP>
__global__ void __launch_bounds__(kThreads, vec_min_blocks<E, OP>())
k_vector(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
         const double* __restrict__ lambda, const double* __restrict__ v, double obj_weight,
         double* __restrict__ out, double* __restrict__ partials) {
  using I = typename E::Integrand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double fsum = 0.0;
  double ksum[E::NP > 0 ? E::NP : 1];
#pragma unroll
  for (int k = 0; k < (E::NP > 0 ? E::NP : 1); ++k) ksum[k] = 0.0;

  const int64_t vstride = (int64_t)gridDim.x * kWarpsPerCta;
  Cell<E> c, nx;  // the dof ids of the next tile are loaded one iteration ahead
  bool nvalid;
  int64_t ncell = lane_cell(D, (int64_t)blockIdx.x * kWarpsPerCta + warp, lane, nvalid);
  load_ids<E>(D, ncell, nvalid, nx);
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + warp; tile < D.ntiles; tile += vstride) {
    const int64_t cell = ncell;
    const bool valid = nvalid;
#pragma unroll
    for (int a = 0; a < E::NY; ++a) c.gy[a] = nx.gy[a];
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) c.gu[b] = nx.gu[b];
    ncell = lane_cell(D, tile + vstride, lane, nvalid);
    load_ids<E>(D, ncell, nvalid, nx);
    gather_state<E>(D, x, valid, c);
    if (!valid) continue;
    double ey[E::NY], eu[E::NUS];  // element vector, summed over the quadrature points
#pragma unroll
    for (int a = 0; a < E::NY; ++a) ey[a] = 0.0;
#pragma unroll
    for (int b = 0; b < E::NUS; ++b) eu[b] = 0.0;
#pragma unroll 1
    for (int q = 0; q < E::NQ; ++q) {
      // (the objective and the residual read the coefficient fields; their derivatives only for some integrands)
      constexpr bool VAL = OP == OP_OBJ || OP == OP_GRAD || OP == OP_OBJGRAD || OP == OP_CONS || OP == OP_OBJCONS;
      const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, VAL || I::NEEDS_COEF_DERIV);
      const double w = pc.w;
      if (OP == OP_OBJ) {
        double s[E::M];
        point_state<E>(tab, pc, c, x, q, s);
        fsum += w * I::f(pc, s);
      } else if (OP == OP_GRAD || OP == OP_OBJGRAD) {
        Jet1<E::M> s[E::M];
        point_state<E>(tab, pc, c, x, q, s);
        const Jet1<E::M> F = I::f(pc, s);
        if (OP == OP_OBJGRAD) fsum += w * F.v;
        double gs[E::M];
#pragma unroll
        for (int k = 0; k < E::M; ++k) gs[k] = w * F.g[k];
        accumulate_fields<E>(tab, q, gs, ey, eu);
#pragma unroll
        for (int k = 0; k < E::NP; ++k) ksum[k] += gs[E::SK + k];
      } else if (OP == OP_CONS || OP == OP_OBJCONS) {
        double s[E::M], R[E::NR];
        point_state<E>(tab, pc, c, x, q, s);
        if (OP == OP_OBJCONS) fsum += w * I::f(pc, s);
        eval_res<E>(pc, s, R);
#pragma unroll
        for (int f = 0; f < E::NFY; ++f)
#pragma unroll
          for (int i = 0; i < E::NLY; ++i) ey[f * E::NLY + i] = fma(w, test_contract<E>(tab, q, f, i, R), ey[f * E::NLY + i]);
      } else if (OP == OP_JPROD) {
        Jet1<E::M> s[E::M], R[E::NR];
        point_state<E>(tab, pc, c, x, q, s);
        eval_res<E>(pc, s, R);
        double ds[E::M];
        field_direction<E>(tab, q, D, c, v, ds);
        double dR[E::NR];
#pragma unroll
        for (int t = 0; t < E::NR; ++t) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < E::M; ++k) a = fma(R[t].g[k], ds[k], a);
          dR[t] = a * w;
        }
#pragma unroll
        for (int f = 0; f < E::NFY; ++f)
#pragma unroll
          for (int i = 0; i < E::NLY; ++i) ey[f * E::NLY + i] += test_contract<E>(tab, q, f, i, dR);
      } else if (OP == OP_JTPROD) {
        Jet1<E::M> s[E::M];
        point_state<E>(tab, pc, c, x, q, s);
        double w_e[E::NY], Lam[E::NR];
        gather_free<E::NY>(v, c.gy, w_e);
        lambda_slots<E>(tab, q, w_e, Lam);
        const Jet1<E::M> L = weighted_res<E>(pc, s, Lam);
        double gs[E::M];
#pragma unroll
        for (int k = 0; k < E::M; ++k) gs[k] = w * L.g[k];
        accumulate_fields<E>(tab, q, gs, ey, eu);
#pragma unroll
        for (int k = 0; k < E::NP; ++k) ksum[k] += gs[E::SK + k];
      } else if (OP == OP_HPROD) {
        Jet2<E::M> s[E::M];
        point_state<E>(tab, pc, c, x, q, s);
        double ds[E::M];
        field_direction<E>(tab, q, D, c, v, ds);
        double gs[E::M];
#pragma unroll
        for (int k = 0; k < E::M; ++k) gs[k] = 0.0;
        if (obj_weight != 0.0) {
          const Jet2<E::M> F = I::f(pc, s);
#pragma unroll
          for (int k = 0; k < E::M; ++k) {
            double a = 0.0;
#pragma unroll
            for (int l = 0; l < E::M; ++l) a = fma(F.h[Jet2<E::M>::idx(k, l)], ds[l], a);
            gs[k] = (obj_weight * w) * a;
          }
        }
        if (lambda != nullptr) {
          double lam_e[E::NY], Lam[E::NR];
          gather_free<E::NY>(lambda, c.gy, lam_e);
          lambda_slots<E>(tab, q, lam_e, Lam);
          const Jet2<E::M> L = weighted_res<E>(pc, s, Lam);
#pragma unroll
          for (int k = 0; k < E::M; ++k) {
            double a = 0.0;
#pragma unroll
            for (int l = 0; l < E::M; ++l) a = fma(L.h[Jet2<E::M>::idx(k, l)], ds[l], a);
            gs[k] = fma(w, a, gs[k]);
          }
        }
        accumulate_fields<E>(tab, q, gs, ey, eu);
#pragma unroll
        for (int k = 0; k < E::NP; ++k) ksum[k] += gs[E::SK + k];
      }
    }
    if (OP == OP_GRAD || OP == OP_OBJGRAD || OP == OP_JTPROD || OP == OP_HPROD) scatter_element<E>(D, c, ey, eu, out);
    if (OP == OP_CONS || OP == OP_OBJCONS || OP == OP_JPROD) {  // rows = free dofs of the test space
#pragma unroll
      for (int i = 0; i < E::NY; ++i)
        if (c.gy[i] > 0) accumulate(D, out + (c.gy[i] - 1), ey[i]);
    }
  }
  // block-level reductions: objective partial sums (deterministic two-pass) and kappa entries
  __shared__ double red[kWarpsPerCta];
  if (OP == OP_OBJ || OP == OP_OBJGRAD || OP == OP_OBJCONS) {
    fsum = warp_sum(fsum);
    if (lane == 0) red[warp] = fsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (int w = 0; w < kWarpsPerCta; ++w) a += red[w];
      partials[blockIdx.x] = a;
    }
    __syncthreads();
  }
  constexpr bool KP = E::NP > 0 && (OP == OP_GRAD || OP == OP_OBJGRAD || OP == OP_JTPROD || OP == OP_HPROD);
  if (KP) {
#pragma unroll
    for (int k = 0; k < E::NP; ++k) cta_partial(ksum[k], partials + gridDim.x, E::NP, k, red);
  }
  // OP_OBJ: the last CTA sums the partials (no second launch).  The ops that scatter with atomics keep the second
  // launch (k_sum_partials / k_add_partials): the __threadfence of the ticket protocol waits for the SM's outstanding
  // red.global.add traffic (measured on config 5: jtprod! 0.173 -> 0.195 ms with the in-kernel second pass).
  if (OP == OP_OBJ) {
    if (last_cta(D.ticket)) finish_objective(D, partials);
  }
}

// Zero fill of a COO segment that is structurally zero (constraint FE Hessian block of an affine
// residual, or hess_coord! without multipliers): pure streaming 16 B stores, 4 per thread per step.
static __global__ void __launch_bounds__(256) k_zero_fill(double* __restrict__ p, int64_t n) {
  // Launched behind the emitting kernel of the same callback (disjoint block): it runs concurrently with it, lets the
  // next kernel start its prologue, and ONE thread waits for the predecessor at the end so that the completion of this grid
  // implies theirs (all threads waiting would park the early CTAs on the SMs until the emitting kernel is done).
  griddep_launch_dependents();
  const int64_t head = (n > 0 && ((reinterpret_cast<uintptr_t>(p) >> 3) & 1)) ? 1 : 0;  // 8 B-aligned start
  const int64_t n2 = (n - head) >> 1;
  double2* q = reinterpret_cast<double2*>(p + head);
  const double2 z = make_double2(0.0, 0.0);
  // each CTA owns contiguous 16 KB chunks (4 x 256 x 16 B)
  const int64_t nchunks = (n2 + 1023) >> 10;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t i0 = (c << 10) + threadIdx.x;
    if (i0 + 768 < n2) {
      q[i0] = z; q[i0 + 256] = z; q[i0 + 512] = z; q[i0 + 768] = z;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (i0 + k * 256 < n2) q[i0 + k * 256] = z;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (head) p[0] = 0.0;
    if ((n - head) & 1) p[n - 1] = 0.0;
    griddep_wait();  // one thread is enough: the grid is not complete before it returns
  }
}

// fixed-order final sum of the per-CTA objective partials
static __global__ void k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += partials[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// ------------------------------------------------------------------ kappa blocks (dense, dof-indexed)
// Jacobian k-block: Jk[i,k] = d c_i / d kappa_k, column-major ncon x p, accumulated with atomics
// into a zero-filled block (/root/reference/src/jacobian_functions.jl:135-142).
template <class E>
__global__ void __launch_bounds__(kThreads)
k_jac_kappa(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x, double* __restrict__ vk) {
  using I = typename E::Integrand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + warp; tile < D.ntiles; tile += (int64_t)gridDim.x * kWarpsPerCta) {
    bool valid;
    const int64_t cell = lane_cell(D, tile, lane, valid);
    if (!valid) continue;
    Cell<E> c;
    load_ids<E>(D, cell, valid, c);
    gather_state<E>(D, x, valid, c);
    constexpr int NPS = E::NP > 0 ? E::NP : 1;
    double acc[NPS][E::NY];  // sum over quadrature points first: one atomic per (dof, kappa)
#pragma unroll
    for (int k = 0; k < NPS; ++k)
#pragma unroll
      for (int i = 0; i < E::NY; ++i) acc[k][i] = 0.0;
#pragma unroll 1
    for (int q = 0; q < E::NQ; ++q) {
      Jet1<E::M> s[E::M], R[E::NR];
      const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, true);
      point_state<E>(tab, pc, c, x, q, s);
      eval_res<E>(pc, s, R);
#pragma unroll
      for (int k = 0; k < E::NP; ++k)
#pragma unroll
        for (int i = 0; i < E::NY; ++i) {
          double a = 0.0;
#pragma unroll
          for (int t = 0; t < E::NR; ++t) a = fma(test_slot<E>(tab, q, i, t), R[t].g[E::SK + k], a);
          acc[k][i] = fma(pc.w, a, acc[k][i]);
        }
    }
#pragma unroll
    for (int k = 0; k < E::NP; ++k)
#pragma unroll
      for (int i = 0; i < E::NY; ++i)
        if (c.gy[i] > 0) accumulate(D, vk + ((int64_t)k * D.nfree_y + (c.gy[i] - 1)), acc[k][i]);
  }
}

// Hessian k-trapezoid: entries (row i, kappa col j), j <= i, "i fastest"; prows = p (inde) or nvar.
__device__ __forceinline__ int64_t trap_pos(int64_t i1, int j1, int64_t prows) {  // 1-based (i, j)
  int64_t pos = 0;
  for (int jj = 1; jj < j1; ++jj) pos += prows - jj + 1;
  return pos + (i1 - j1);
}

// NoFETerm objective fk(kappa): a p-variable scalar function, differentiated by ONE thread with a
// Jet2<p> (the reference runs ForwardDiff.gradient / hessian on it, additional_obj_terms.jl:88-106).
enum NoFeMode { NOFE_OBJ = 0, NOFE_GRAD = 1, NOFE_HESS = 2, NOFE_HPROD = 3 };
template <class E>
__global__ void k_nofe(const __grid_constant__ Tables<E> tab, const double* __restrict__ x, int mode, double obj_weight,
                       const double* __restrict__ v, double* __restrict__ out) {
  constexpr int NP = E::NP > 0 ? E::NP : 1;
  using J2 = Jet2<NP>;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  J2 k[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) jet_seed(k[i], x[i], i);
  const J2 F = E::Integrand::fk(tab.par, k);
  if (mode == NOFE_OBJ) {
    out[0] = F.v;
  } else if (mode == NOFE_GRAD) {        // out = g (zero-filled by the caller)
    for (int i = 0; i < E::NP; ++i) out[i] += F.g[i];
  } else if (mode == NOFE_HESS) {        // out = the p x p lower triangle, i fastest
    for (int j = 0; j < E::NP; ++j)
      for (int i = j; i < E::NP; ++i) out[trap_pos(i + 1, j + 1, E::NP)] = obj_weight * F.h[J2::idx(i, j)];
  } else {                               // out = Hv (kappa rows), v = direction
    for (int i = 0; i < E::NP; ++i) {
      double a = 0.0;
      for (int l = 0; l < E::NP; ++l) a = fma(F.h[J2::idx(i, l)], v[l], a);
      out[i] += obj_weight * a;
    }
  }
}
// Objective kappa-block of an `inde` energy (MixedEnergyFETerm(..., inde = true): only the p x p kappa-kappa
// triangle is stored): the pointwise fields enter as constants, only kappa is seeded, so the jets are Jet2<p>
// instead of Jet2<M> (config 5: 0.20 -> 0.03 ms).
template <class E>
__global__ void __launch_bounds__(kThreads)
k_hess_kappa_kk(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x, double obj_weight,
                int64_t prows, double* __restrict__ vk, double* __restrict__ partials) {
  using I = typename E::Integrand;
  constexpr int NPS = E::NP > 0 ? E::NP : 1;
  using JK = Jet2<NPS>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double kk[NPS][NPS];
#pragma unroll
  for (int a = 0; a < NPS; ++a)
#pragma unroll
    for (int b = 0; b < NPS; ++b) kk[a][b] = 0.0;
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + warp; tile < D.ntiles; tile += (int64_t)gridDim.x * kWarpsPerCta) {
    bool valid;
    const int64_t cell = lane_cell(D, tile, lane, valid);
    if (!valid) continue;
    Cell<E> c;
    load_ids<E>(D, cell, valid, c);
    gather_state<E>(D, x, valid, c);
#pragma unroll 1
    for (int q = 0; q < E::NQ; ++q) {
      const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, true);
      double val[E::M];
      point_state<E>(tab, pc, c, x, q, val);
      JK s[E::M];
#pragma unroll
      for (int k = 0; k < E::M; ++k) {
        if (k >= E::SK && k < E::SK + E::NP) jet_seed(s[k], val[k], k - E::SK);
        else s[k] = jet_const(val[k], s[k]);
      }
      const JK F = I::f(pc, s);
      const double scale = pc.w * obj_weight;
#pragma unroll
      for (int j = 0; j < E::NP; ++j)
#pragma unroll
        for (int i = j; i < E::NP; ++i) kk[i][j] += scale * F.h[JK::idx(i, j)];
    }
  }
  __shared__ double red[kWarpsPerCta];
  int comp = 0;
#pragma unroll
  for (int j = 0; j < E::NP; ++j)
#pragma unroll
    for (int i = j; i < E::NP; ++i) { cta_partial(kk[i][j], partials + gridDim.x, E::NP * (E::NP + 1) / 2, comp, red); ++comp; }
  if (last_cta(D.ticket)) finish_partials(partials + gridDim.x, E::NP * (E::NP + 1) / 2, E::NP, prows, vk);
}

// which: 0 = objective (scaled by obj_weight), 1 = constraints (lambda-weighted residual), 2 = both from one sweep
// (objective into vk / prows, constraints into vk2 / prows2; the cell's ids and state are gathered once and the
// callback saves two launches — hess_coord! of the 1-D problems is a chain of 20-30 us launches)
template <class E>
__global__ void __launch_bounds__(kThreads)
k_hess_kappa(const __grid_constant__ Tables<E> tab, const DevModel D, const double* __restrict__ x,
             const double* __restrict__ lambda, double obj_weight, int which, int64_t prows, double* __restrict__ vk,
             double* __restrict__ partials, int64_t prows2, double* __restrict__ vk2) {
  using I = typename E::Integrand;
  using J2 = Jet2<E::M>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NPS = E::NP > 0 ? E::NP : 1;
  const int w0 = which == 2 ? 0 : which, w1 = which == 2 ? 1 : which;
  double kk[2][NPS][NPS];
#pragma unroll
  for (int w = 0; w < 2; ++w)
#pragma unroll
    for (int a = 0; a < NPS; ++a)
#pragma unroll
      for (int b = 0; b < NPS; ++b) kk[w][a][b] = 0.0;
  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + warp; tile < D.ntiles; tile += (int64_t)gridDim.x * kWarpsPerCta) {
    bool valid;
    const int64_t cell = lane_cell(D, tile, lane, valid);
    if (!valid) continue;
    Cell<E> c;
    load_ids<E>(D, cell, valid, c);
    gather_state<E>(D, x, valid, c);
    double lam_e[E::NY];
    if (w1 != 0) gather_free<E::NY>(lambda, c.gy, lam_e);
#pragma unroll
    for (int w = 0; w < 2; ++w) {  // (unrolled: kk[w] stays in registers)
      if (w < w0 || w > w1) continue;
      const bool second = which == 2 && w == 1;
      const int64_t pr = second ? prows2 : prows;
      double* const out = second ? vk2 : vk;
      const bool cross = pr > E::NP;
      double accy[NPS][E::NY], accu[NPS][E::NUS];  // cross terms summed over quadrature points
#pragma unroll
      for (int j = 0; j < NPS; ++j) {
#pragma unroll
        for (int a = 0; a < E::NY; ++a) accy[j][a] = 0.0;
#pragma unroll
        for (int b = 0; b < E::NUS; ++b) accu[j][b] = 0.0;
      }
#pragma unroll 1
      for (int q = 0; q < E::NQ; ++q) {
        Jet2<E::M> s[E::M];
        const PointGeo<E> pc = point_ctx<E>(D, tab, cell, valid, q, true);
        point_state<E>(tab, pc, c, x, q, s);
        J2 L;
        double scale = pc.w;
        if (w == 0) { L = I::f(pc, s); scale *= obj_weight; }
        else {
          double Lam[E::NR];
          lambda_slots<E>(tab, q, lam_e, Lam);
          L = weighted_res<E>(pc, s, Lam);
        }
#pragma unroll
        for (int j = 0; j < E::NP; ++j) {
#pragma unroll
          for (int i = j; i < E::NP; ++i) kk[w][i][j] += scale * L.h[J2::idx(E::SK + i, E::SK + j)];
          if (cross) {
#pragma unroll
            for (int a = 0; a < E::NY; ++a) {
              double v = 0.0;
#pragma unroll
              for (int k = 0; k < E::NB; ++k) v = fma(tab.By[q][a][k], L.h[J2::idx(k, E::SK + j)], v);
              accy[j][a] = fma(scale, v, accy[j][a]);
            }
#pragma unroll
            for (int b = 0; b < E::NU; ++b) accu[j][b] = fma(scale, tab.Bu[q][b] * L.h[J2::idx(E::SU, E::SK + j)], accu[j][b]);
          }
        }
      }
      if (cross) {
#pragma unroll
        for (int j = 0; j < E::NP; ++j) {
#pragma unroll
          for (int a = 0; a < E::NY; ++a)
            if (c.gy[a] > 0) accumulate(D, out + trap_pos((int64_t)E::NP + c.gy[a], j + 1, pr), accy[j][a]);
#pragma unroll
          for (int b = 0; b < E::NU; ++b)
            if (c.gu[b] > 0) accumulate(D, out + trap_pos((int64_t)E::NP + D.nfree_y + c.gu[b], j + 1, pr), accu[j][b]);
        }
      }
    }
  }
  __shared__ double red[kWarpsPerCta];
  constexpr int NKK = E::NP * (E::NP + 1) / 2;
  const int ncomp = NKK * (w1 - w0 + 1);
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    if (w < w0 || w > w1) continue;  // (block-uniform)
    int comp = 0;
#pragma unroll
    for (int j = 0; j < E::NP; ++j)
#pragma unroll
      for (int i = j; i < E::NP; ++i) { cta_partial(kk[w][i][j], partials + gridDim.x, ncomp, (w - w0) * NKK + comp, red); ++comp; }
  }
}

// ------------------------------------------------------------------ create-time self-check helpers
// deterministic pseudo-random fill in [lo, hi) (splitmix64 of the index) and max |a - b|, max |b| of two arrays
static __global__ void k_fill_random(double* __restrict__ p, int64_t n, uint64_t seed, double lo, double hi) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = lo + (hi - lo) * ((double)(z >> 11) * (1.0 / 9007199254740992.0));
  }
}
static __global__ void k_max_diff(const double* __restrict__ a, const double* __restrict__ b, int64_t n, unsigned long long* __restrict__ res /* [2] */) {
  double md = 0.0, mb = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = fabs(a[i] - b[i]);
    md = (d > md || d != d) ? d : md;  // a NaN sticks
    mb = fmax(mb, fabs(b[i]));
  }
  // non-negative doubles order like their bit patterns (NaN patterns sort above every finite value)
  atomicMax(res + 0, (unsigned long long)__double_as_longlong(md));
  atomicMax(res + 1, (unsigned long long)__double_as_longlong(mb));
}

}  // namespace pdenlp
