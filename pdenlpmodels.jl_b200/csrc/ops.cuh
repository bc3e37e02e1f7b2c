// ops.cuh — type-erased per-element operations of libpdenlp_cuda (host side of the kernel launches).
// Included by pdenlp_cuda.cu and by the ops_*.cu translation units, which instantiate disjoint subsets of the
// (integrand, element) set so that they can be compiled in parallel.
#pragma once
#include "../../include/pdenlp_cuda.h"
#include "kernels.cuh"
#include "kvector.cuh"
#include "kmma.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

extern thread_local std::string pdenlp_g_err;  // defined in pdenlp_cuda.cu

namespace pdenlp {
namespace host {


inline int fail(int code, const std::string& msg) {
  pdenlp_g_err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(PDENLP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
  } while (0)

struct Launch {
  cudaStream_t stream;
  int sms;
  // true: ignore the integrand traits that select shortcuts (reference-tensor paths, structurally-zero blocks) and run
  // the generic per-quadrature-point kernels — the create-time self-check compares both (pdenlp_create)
  bool generic = false;
  int64_t* launches = nullptr;  // counts the kernel launches issued through this Launch (pdenlp_launch_count)
  // auxiliary stream (already forked from `stream` by the caller, who also joins it): the cell-granular template path
  // runs its generic-cell kernel there, concurrently with the chunk copy; nullptr: everything on `stream`
  cudaStream_t aux = nullptr;
  // true (and aux == nullptr): the template path chains its chunk copy behind its generic-cell kernel on `stream` with
  // programmatic stream serialization instead of forking the auxiliary stream (models without kappa-block kernels)
  bool chain = false;
  // hess_coord!: a structurally zero block the template path may fill in its copy launch (sets *fill_done)
  double* fill_ptr = nullptr;
  int64_t fill_n = 0;
  bool* fill_done = nullptr;
  int bg_ctas = 0;  // > 0: CTAs per SM of the kappa-block kernels when they run in the background of a streaming kernel
};
inline void count_launch(const Launch& L, int n = 1) { if (L.launches) *L.launches += n; }

// Kernel launch with the programmatic-stream-serialization attribute (see griddep_wait in kernels.cuh): the kernel may
// be scheduled while its predecessor on the stream drains; it synchronises itself with griddepcontrol.wait.
// PDENLP_PDL_CHAIN=1 (experiment) turns it on for the coord kernels; the fill behind k_hess_coord always uses it.
inline bool pdl_chain() {
  static const bool on = getenv("PDENLP_PDL_CHAIN") != nullptr;
  return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*k)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, std::forward<Args>(args)...);
}

// type-erased per-element operations
struct Ops {
  virtual ~Ops() {}
  virtual int count(const DevModel& D, int32_t* counts, int32_t* flag, cudaStream_t s) = 0;
  virtual int jac(const Launch& L, const DevModel& D, const double* x, double* vy, double* vu) = 0;
  virtual int hess(const Launch& L, const DevModel& D, const double* x, const double* lam, double ow, double* vo, double* vc) = 0;
  virtual int structure(const Launch& L, const DevModel& D, int kind, int64_t* r, int64_t* c, int64_t shift) = 0;
  virtual int vec(const Launch& L, const DevModel& D, int op, const double* x, const double* lam, const double* v,
                  double ow, double* out, double* partials, int* nblocks) = 0;
  virtual int jac_kappa(const Launch& L, const DevModel& D, const double* x, double* vk, double* partials) = 0;
  // which: 0 = objective trapezoid, 1 = constraint trapezoid, 2 = both (objective: prows / vk, constraints: prows2 / vk2)
  virtual int hess_kappa(const Launch& L, const DevModel& D, const double* x, const double* lam, double ow, int which,
                         int64_t prows, double* vk, double* partials, int64_t prows2 = 0, double* vk2 = nullptr) = 0;
  virtual bool has_shortcuts() const = 0;  // some trait-selected fast path exists (worth a create-time self-check)
  virtual int run_h() const = 0;
  virtual int run_jy() const = 0;
  virtual int run_ju() const = 0;
  // template tiles (kernels.cuh): which coordinate kernels have the path; per-tile classification against a signature
  virtual bool tpl_jac_ok() const = 0;
  virtual bool tpl_hess_ok() const = 0;
  virtual int tile_class(const DevModel& D, unsigned long long sig_y, unsigned long long sig_u, uint8_t* cls, cudaStream_t s) = 0;
  virtual int cell_class(const DevModel& D, unsigned long long sig_y, unsigned long long sig_u, uint32_t* mask, cudaStream_t s) = 0;
  virtual bool needs_coef_deriv() const = 0;  // derivative callbacks read the coefficient fields
  virtual bool res_fe_hessian_zero() const = 0;  // constraint FE Hessian block is structurally zero
  virtual bool no_fe_obj() const = 0;            // NoFETerm objective fk(kappa)
  virtual bool is_multi() const = 0;             // multi-field element (block-ordered kernels)
  virtual int nofe(const Launch& L, int mode, const double* x, double ow, const double* v, double* out) = 0;
  virtual const std::vector<double>& ktab() const = 0;  // pre-integrated reference tensors (may be empty)
};

constexpr int warps_for_stage(int stage_doubles, int reserved_doubles) {
  // per-warp staging buffer; at most 8 warps (launch bounds), at most ~225 KB per CTA
  return std::min(8, (225 * 1024 - reserved_doubles * 8) / (stage_doubles * 8));
}

template <class E>
struct OpsImpl : Ops {
  Tables<E> tab;
  bool attr_done = false;
  std::vector<double> ktab_;
  KConst<E> kc;  // the same tensors as a by-value kernel parameter (K-path vector kernels)
  const std::vector<double>& ktab() const override { return ktab_; }
  // per-warp staging = (buffer size) x (number of rotating buffers, > 1 only for small elements)
  static constexpr int STJ = jac_stage_doubles<E>() * coord_nbuf(jac_stage_doubles<E>());
  static constexpr int STH = hess_stage_doubles<E>() * coord_nbuf(hess_stage_doubles<E>());
  static constexpr int WJ = warps_for_stage(STJ, KTab<E>::TOTAL);
  static constexpr int WH = warps_for_stage(STH, KTab<E>::TOTAL);
  static constexpr size_t SMEM_J = ((size_t)WJ * STJ + KTab<E>::TOTAL) * 8;
  static constexpr size_t SMEM_H = ((size_t)WH * STH + KTab<E>::TOTAL) * 8;

  // multi-field elements run the block-ordered kernels
  using JacKernel = void (*)(const Tables<E>, const DevModel, const double*, double*, double*);
  using HessKernel = void (*)(const Tables<E>, const DevModel, const double*, const double*, double, double*, double*);
  static JacKernel jac_kernel() {
    if constexpr (E::MULTI) return k_jac_coord_mf<E>; else return k_jac_coord<E>;
  }
  static HessKernel hess_kernel() {
    if constexpr (E::MULTI) return k_hess_coord_mf<E>; else return k_hess_coord<E>;
  }
  int occ_j = 1, occ_h = 1;  // resident CTAs per SM of the persistent coord kernels
  int set_attrs() {
    if (attr_done) return PDENLP_OK;
    CU(cudaFuncSetAttribute(jac_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_J));
    CU(cudaFuncSetAttribute(hess_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_H));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_j, jac_kernel(), WJ * 32, SMEM_J));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_h, hess_kernel(), WH * 32, SMEM_H));
    occ_j = std::max(1, std::min(occ_j, 4));
    occ_h = std::max(1, std::min(occ_h, 4));
    attr_done = true;
    return PDENLP_OK;
  }
  int run_h() const override { return E::RUN_H; }
  int run_jy() const override { return E::RUN_JY; }
  int run_ju() const override { return E::RUN_JU; }
  bool tpl_jac_ok() const override { return tpl_ok<E, false>(); }
  bool tpl_hess_ok() const override { return tpl_ok<E, true>(); }
  int tile_class(const DevModel& D, unsigned long long sig_y, unsigned long long sig_u, uint8_t* cls, cudaStream_t s) override {
    k_tile_class<E><<<grid_for(D.ntiles, kWarpsPerCta, 1 << 30), kThreads, 0, s>>>(D, sig_y, sig_u, cls);
    CU(cudaGetLastError());
    return PDENLP_OK;
  }
  int cell_class(const DevModel& D, unsigned long long sig_y, unsigned long long sig_u, uint32_t* mask, cudaStream_t s) override {
    k_cell_class<E><<<grid_for(D.ntiles, kWarpsPerCta, 1 << 30), kThreads, 0, s>>>(D, sig_y, sig_u, mask);
    CU(cudaGetLastError());
    return PDENLP_OK;
  }
  bool needs_coef_deriv() const override { return E::Integrand::NEEDS_COEF_DERIV; }
  bool res_fe_hessian_zero() const override { return E::Integrand::RES_FE_HESSIAN_ZERO; }
  bool no_fe_obj() const override { return E::Integrand::NO_FE_OBJ; }
  bool is_multi() const override { return E::MULTI; }
  bool has_shortcuts() const override {
    using I = typename E::Integrand;
    return I::RES_FE_HESSIAN_ZERO || (KTab<E>::ENABLED && I::FE_DERIVS_POINT_INDEPENDENT) || tpl_ok<E, false>() || tpl_ok<E, true>();
  }
  int nofe(const Launch& L, int mode, const double* x, double ow, const double* v, double* out) override {
    if constexpr (E::Integrand::NO_FE_OBJ) {
      k_nofe<E><<<1, 32, 0, L.stream>>>(tab, x, mode, ow, v, out);
      CU(cudaGetLastError());
      count_launch(L);
    }
    return PDENLP_OK;
  }
  static int grid_for(int64_t ntiles, int warps, int cap) {
    int64_t g = (ntiles + warps - 1) / warps;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
  }
  int count(const DevModel& D, int32_t* counts, int32_t* flag, cudaStream_t s) override {
    k_count<E><<<grid_for(D.ntiles, kWarpsPerCta, 1 << 30), kThreads, 0, s>>>(D, counts, flag);
    CU(cudaGetLastError());
    return PDENLP_OK;
  }
  // tuning knob (experiments only): PDENLP_WARPS_J / PDENLP_WARPS_H lower the warps per CTA
  static int env_warps(const char* name, int maxw) {
    const char* e = getenv(name);
    if (!e) return maxw;
    int w = atoi(e);
    return w >= 1 && w <= maxw ? w : maxw;
  }
  // cell-granular templates (kernels.cuh): generic cells (256 per CTA) on the auxiliary stream, chunk copy (one per CTA)
  static int cg_cells_grid(const DevModel& D) { return (D.ngc + kThreads - 1) / kThreads; }
  // threads per CTA of the chunk copy (tuning knob PDENLP_COPY_THREADS = 64 | 128 | 256).  Measured (tools/ab_copy_threads.sh,
  // config 2 jac_coord! ms): 64: 0.628, 128: 0.570, 256: 0.546, 384: 0.618, 512: 0.738 — a sharp optimum at 256
  static int copy_threads() {
    static const int t = [] { const char* e = getenv("PDENLP_COPY_THREADS"); int v = e ? atoi(e) : kThreads; return (v == 64 || v == 128 || v == 256) ? v : kThreads; }();
    return t;
  }
  int jac(const Launch& L, const DevModel& D, const double* x, double* vy, double* vu) override {
    if constexpr (tpl_ok<E, false>()) {
      if (D.ch != nullptr && !D.notrait && !L.generic) {
        if (D.ngc > 0) {
          CU(launch_kernel(k_jac_cells_cg<E>, cg_cells_grid(D), kThreads, 0, L.aux ? L.aux : L.stream, false, tab, D, x, vy, vu));
          count_launch(L);
        }
        if (D.nchunk > 0) {
          const ChunkSeg a{vy, D.tpl_jy, E::RUN_JY};
          const ChunkSeg b{E::NU > 0 ? vu : nullptr, D.tpl_ju, E::RUN_JU};
          const bool chained = L.chain && !L.aux && D.ngc > 0;
          CU(launch_kernel(k_copy_chunks, D.nchunk, copy_threads(), 0, L.stream, chained || pdl_chain(), a, b, D.ch, 0, chained ? 1 : 0, D.nchunk, (double*)nullptr, (int64_t)0));
          count_launch(L);
        }
        return PDENLP_OK;
      }
    }
    if (int rc = set_attrs()) return rc;
    static const int wj = env_warps("PDENLP_WARPS_J", WJ);
    CU(launch_kernel(jac_kernel(), grid_for(D.ntiles, wj, L.sms * occ_j), wj * 32, SMEM_J, L.stream, pdl_chain() && !E::MULTI,
                     tab, D, x, vy, vu));
    count_launch(L);
    return PDENLP_OK;
  }
  int hess(const Launch& L, const DevModel& D, const double* x, const double* lam, double ow, double* vo, double* vc) override {
    if constexpr (tpl_ok<E, true>()) {
      if (D.ch != nullptr && !D.notrait && !L.generic && vo != nullptr && vc == nullptr) {
        if (D.ngc > 0) {
          CU(launch_kernel(k_hess_cells_cg<E>, cg_cells_grid(D), kThreads, 0, L.aux ? L.aux : L.stream, false, tab, D, x, ow, vo));
          count_launch(L);
        }
        if (D.nchunk > 0) {
          const ChunkSeg a{vo, D.tpl_h, E::RUN_H};
          const ChunkSeg b{nullptr, nullptr, 0};
          const bool chained = L.chain && !L.aux && D.ngc > 0;
          // the zero fill of the constraint block rides in the same launch (tools/ab_merge_fill.sh: 1/8-slab step
          // 0.1958 -> 0.1946 ms, full mesh 1.3834 -> 1.3806 ms); PDENLP_NO_MERGE_FILL=1: its own PDL launch
          static const bool merge_fill = getenv("PDENLP_NO_MERGE_FILL") == nullptr;
          const bool fill = merge_fill && L.fill_ptr != nullptr && L.fill_n >= 4096 && L.fill_done != nullptr && copy_threads() == kThreads;
          const int64_t nfill = fill ? ((L.fill_n / 2 + 1023) >> 10) : 0;
          CU(launch_kernel(k_copy_chunks, (int)(D.nchunk + nfill), copy_threads(), 0, L.stream, chained || pdl_chain(), a, b, D.ch, 1, chained ? 1 : 0,
                           D.nchunk, fill ? L.fill_ptr : (double*)nullptr, fill ? L.fill_n : (int64_t)0));
          if (fill) *L.fill_done = true;
          count_launch(L);
        }
        return PDENLP_OK;
      }
    }
    if (int rc = set_attrs()) return rc;
    static const int wh = env_warps("PDENLP_WARPS_H", WH);
    CU(launch_kernel(hess_kernel(), grid_for(D.ntiles, wh, L.sms * occ_h), wh * 32, SMEM_H, L.stream, pdl_chain() && !E::MULTI,
                     tab, D, x, lam, ow, vo, vc));
    count_launch(L);
    return PDENLP_OK;
  }
  int structure(const Launch& L, const DevModel& D, int kind, int64_t* r, int64_t* c, int64_t shift) override {
    k_structure<E><<<grid_for(D.ntiles, kWarpsPerCta, 1 << 30), kThreads, 0, L.stream>>>(D, kind, r, c, shift);
    CU(cudaGetLastError());
    count_launch(L);
    return PDENLP_OK;
  }
  // launches the (generic or reference-tensor) kernel of a dof-indexed callback, then the fixed-order second pass that
  // adds the per-CTA partials of its kappa entries to out[0..p)
  // threads per CTA of the reference-tensor kernels (tuning knob: PDENLP_KVEC_THREADS = 128 | 256)
  static int kvec_threads() {
    static const int t = [] { const char* e = getenv("PDENLP_KVEC_THREADS"); int v = e ? atoi(e) : kThreads; return (v == 64 || v == 128 || v == 256) ? v : kThreads; }();
    return t;
  }
  static constexpr size_t kvec_smem(int op) { return kvec_tensors_in_smem(op) ? KTab<E>::TOTAL * 8 : 0; }
  // which ops run on the FP64 tensor cores (kmma.cuh) rather than lane = cell (kvector.cuh): by measurement on config 5
  // (ms, tensor-core / lane = cell): cons! 0.12 / 0.15, hess_coord! kappa trapezoid 0.87 / 0.94 (whole callback),
  // jprod! 0.17 / 0.14, jtprod! 0.17 / 0.16, hprod! 0.20 / 0.12, jac_coord! kappa block 0.61 / 0.57 (whole callback)
  static constexpr bool use_mma(int op) {
#ifdef PDENLP_KMMA_ALL
    return kmma_enabled<E>();
#else
    return kmma_enabled<E>() && (op == OP_CONS || op == OP_HESS_KAPPA);
#endif
  }
  static int kvec_grid(const DevModel& D, const Launch& L, int op) {
    // tensor-core kernels: one wave of resident CTAs (their per-CTA prologue builds the B fragments)
    const int per_sm = use_mma(op) ? PDENLP_KMMA_MINB : 8;
    return grid_for(D.ntiles, kvec_threads() / 32, L.sms * (L.bg_ctas > 0 ? std::min(L.bg_ctas, per_sm) : per_sm));
  }
  template <int OP>
  int vec_op(const Launch& L, const DevModel& D, const double* x, const double* lam, const double* v, double ow,
             double* out, double* partials, int g) {
    if constexpr (kvec_enabled<E>() && kvec_op(OP)) {
      if (!L.generic && D.ktab != nullptr) {
        g = kvec_grid(D, L, OP);
        if constexpr (use_mma(OP))  // FP64 tensor-core version (kmma.cuh)
          k_vector_M<E, OP><<<g, kvec_threads(), 0, L.stream>>>(tab, kc, D, x, lam, v, ow, 0, out, partials);
        else
          k_vector_K<E, OP><<<g, kvec_threads(), kvec_smem(OP), L.stream>>>(tab, kc, D, x, lam, v, ow, 0, out, partials);
        CU(cudaGetLastError());
        count_launch(L);
        return add_kappa_partials<OP>(L, partials, g, out);
      }
    }
    k_vector<E, OP><<<g, kThreads, 0, L.stream>>>(tab, D, x, lam, v, ow, out, partials);
    CU(cudaGetLastError());
    count_launch(L);
    return add_kappa_partials<OP>(L, partials, g, out);
  }
  template <int OP>
  int add_kappa_partials(const Launch& L, double* partials, int g, double* out) {
    constexpr int NPART = kappa_partials<E>(OP);
    if constexpr (NPART > 0) {
      KappaDst dst;
      for (int k = 0; k < NPART; ++k) dst.pos[k] = k;
      k_add_partials<<<1, 256, 0, L.stream>>>(partials + g, g, NPART, dst, out);
      CU(cudaGetLastError());
      count_launch(L);
    }
    return PDENLP_OK;
  }
  int vec(const Launch& L, const DevModel& D, int op, const double* x, const double* lam, const double* v, double ow,
          double* out, double* partials, int* nblocks) override {
    const int g = grid_for(D.ntiles, kWarpsPerCta, L.sms * 8);
    *nblocks = g;
    switch (op) {
      case OP_OBJ: return vec_op<OP_OBJ>(L, D, x, lam, v, ow, out, partials, g);
      case OP_GRAD: return vec_op<OP_GRAD>(L, D, x, lam, v, ow, out, partials, g);
      case OP_CONS: return vec_op<OP_CONS>(L, D, x, lam, v, ow, out, partials, g);
      case OP_JPROD: return vec_op<OP_JPROD>(L, D, x, lam, v, ow, out, partials, g);
      case OP_JTPROD: return vec_op<OP_JTPROD>(L, D, x, lam, v, ow, out, partials, g);
      case OP_HPROD: return vec_op<OP_HPROD>(L, D, x, lam, v, ow, out, partials, g);
      case OP_OBJGRAD: return vec_op<OP_OBJGRAD>(L, D, x, lam, v, ow, out, partials, g);
      case OP_OBJCONS: return vec_op<OP_OBJCONS>(L, D, x, lam, v, ow, out, partials, g);
      default: return fail(PDENLP_EINVAL, "bad vector op");
    }
  }
  int jac_kappa(const Launch& L, const DevModel& D, const double* x, double* vk, double* partials) override {
    if constexpr (E::NP > 0) {
      const int g = grid_for(D.ntiles, kWarpsPerCta, L.sms * 8);
      if constexpr (kvec_enabled<E>()) {
        if (!L.generic && D.ktab != nullptr) {
          if constexpr (use_mma(OP_JAC_KAPPA))
            k_vector_M<E, OP_JAC_KAPPA><<<kvec_grid(D, L, OP_JAC_KAPPA), kvec_threads(), 0, L.stream>>>(tab, kc, D, x, nullptr, nullptr, 1.0, 0, vk, partials);
          else
            k_vector_K<E, OP_JAC_KAPPA><<<kvec_grid(D, L, OP_JAC_KAPPA), kvec_threads(), kvec_smem(OP_JAC_KAPPA), L.stream>>>(tab, kc, D, x, nullptr, nullptr, 1.0, 0, vk, partials);
          CU(cudaGetLastError());
          count_launch(L);
          return PDENLP_OK;
        }
      }
      k_jac_kappa<E><<<g, kThreads, 0, L.stream>>>(tab, D, x, vk);
      CU(cudaGetLastError());
      count_launch(L);
    }
    return PDENLP_OK;
  }
  int hess_kappa(const Launch& L, const DevModel& D, const double* x, const double* lam, double ow, int which,
                 int64_t prows, double* vk, double* partials, int64_t prows2 = 0, double* vk2 = nullptr) override {
    if constexpr (E::NP > 0) {
      int g = grid_for(D.ntiles, kWarpsPerCta, L.sms * 8);
      if (which == 2) {
        // one sweep for both trapezoids where both take the generic kernel; otherwise two calls
        bool fast_con = false;
        if constexpr (kvec_enabled<E>()) fast_con = !L.generic && D.ktab != nullptr;
        constexpr int NKK2 = E::NP * (E::NP + 1) / 2;
        if (fast_con || prows == E::NP || 2 * NKK2 > 16) {
          if (int rc = hess_kappa(L, D, x, nullptr, ow, 0, prows, vk, partials)) return rc;
          return hess_kappa(L, D, x, lam, 1.0, 1, prows2, vk2, partials);
        }
        KappaDst dst2;
        int c = 0;
        for (int w = 0; w < 2; ++w) {
          const int64_t pr = w ? prows2 : prows, base = w ? (vk2 - vk) : 0;
          for (int j = 0; j < E::NP; ++j) {
            int64_t col0 = 0;
            for (int jj = 0; jj < j; ++jj) col0 += pr - jj;
            for (int i = j; i < E::NP; ++i) dst2.pos[c++] = base + col0 + (i - j);
          }
        }
        k_hess_kappa<E><<<g, kThreads, 0, L.stream>>>(tab, D, x, lam, ow, 2, prows, vk, partials, prows2, vk2);
        CU(cudaGetLastError());
        k_add_partials<<<1, 256, 0, L.stream>>>(partials + g, g, 2 * NKK2, dst2, vk);
        CU(cudaGetLastError());
        count_launch(L, 2);
        return PDENLP_OK;
      }
      constexpr int NKK = E::NP * (E::NP + 1) / 2;
      KappaDst dst;  // the kappa-kappa corner: column j outer, row i >= j inner
      {
        int c = 0;
        for (int j = 0; j < E::NP; ++j) {
          int64_t col0 = 0;
          for (int jj = 0; jj < j; ++jj) col0 += prows - jj;
          for (int i = j; i < E::NP; ++i) dst.pos[c++] = col0 + (i - j);
        }
      }
      if (which == 0 && prows == E::NP) {  // objective of an `inde` energy: kappa-kappa triangle only
        // (no scatter in this kernel: its last CTA adds the partials itself, see last_cta)
        k_hess_kappa_kk<E><<<g, kThreads, 0, L.stream>>>(tab, D, x, ow, prows, vk, partials);
        CU(cudaGetLastError());
        count_launch(L);
        return PDENLP_OK;
      }
      bool done = false;
      if constexpr (kvec_enabled<E>()) {
        if (!done && which == 1 && !L.generic && D.ktab != nullptr) {
          g = kvec_grid(D, L, OP_HESS_KAPPA);
          if constexpr (use_mma(OP_HESS_KAPPA))
            k_vector_M<E, OP_HESS_KAPPA><<<g, kvec_threads(), 0, L.stream>>>(tab, kc, D, x, lam, nullptr, 1.0, prows, vk, partials);
          else
            k_vector_K<E, OP_HESS_KAPPA><<<g, kvec_threads(), kvec_smem(OP_HESS_KAPPA), L.stream>>>(tab, kc, D, x, lam, nullptr, 1.0, prows, vk, partials);
          done = true;
        }
      }
      if (!done) k_hess_kappa<E><<<g, kThreads, 0, L.stream>>>(tab, D, x, lam, ow, which, prows, vk, partials, (int64_t)0, (double*)nullptr);
      CU(cudaGetLastError());
      k_add_partials<<<1, 256, 0, L.stream>>>(partials + g, g, NKK, dst, vk);
      CU(cudaGetLastError());
      count_launch(L, 2);
    }
    return PDENLP_OK;
  }
};

template <class E>
Ops* make_ops(const pdenlp_desc& d) {
  auto* o = new OpsImpl<E>();
  for (int q = 0; q < E::NQ; ++q) {
    o->tab.wdet[q] = d.wdet[q];
    for (int a = 0; a < E::NLY; ++a) {
      o->tab.By[q][a][0] = d.phi_y[q * E::NLY + a];
      for (int k = 0; k < E::DIM; ++k) o->tab.By[q][a][1 + k] = d.grad_y[(q * E::NLY + a) * E::DIM + k];
    }
    for (int b = 0; b < E::NLUS; ++b) o->tab.Bu[q][b] = (E::NLU > 0) ? d.phi_u[q * E::NLU + b] : 0.0;
  }
  for (int i = 0; i < 8; ++i) o->tab.par[i] = d.params[i];
  if constexpr (KTab<E>::ENABLED) {  // pre-integrated reference tensors (see KTab in kernels.cuh)
    using KT = KTab<E>;
    o->ktab_.assign(KT::TOTAL, 0.0);
    const auto& T = o->tab;
    for (int q = 0; q < E::NQ; ++q) {
      const double w = T.wdet[q];
      for (int j = 0; j < E::NY; ++j)
        for (int i = 0; i < E::NY; ++i)
          for (int k = 0; k < E::NB; ++k)
            for (int l = 0; l < E::NB; ++l) o->ktab_[KT::yy(k, l, j, i)] += w * T.By[q][i][k] * T.By[q][j][l];
      for (int j = 0; j < E::NY; ++j)
        for (int i = 0; i < E::NU; ++i)
          for (int l = 0; l < E::NB; ++l) {
            o->ktab_[KT::uy(l, j, i)] += w * T.Bu[q][i] * T.By[q][j][l];
            o->ktab_[KT::yu(l, i, j)] += w * T.By[q][j][l] * T.Bu[q][i];
          }
      for (int j = 0; j < E::NU; ++j)
        for (int i = 0; i < E::NU; ++i) o->ktab_[KT::uu(j, i)] += w * T.Bu[q][i] * T.Bu[q][j];
      for (int j = 0; j < E::NY; ++j) {
        for (int i = 0; i < E::NY; ++i)
          for (int d = 1; d < E::NB; ++d) o->ktab_[KT::lap(j, i)] += w * T.By[q][i][d] * T.By[q][j][d];
        for (int k = 0; k < E::NB; ++k) o->ktab_[KT::y1(k, j)] += w * T.By[q][j][k];
      }
      for (int b = 0; b < E::NU; ++b) o->ktab_[KT::u1(b)] += w * T.Bu[q][b];
      o->ktab_[KT::OFF_VOL] += w;
      for (int t = 0; t < E::NB; ++t)
        for (int i = 0; i < E::NY; ++i) o->ktab_[KT::bt(t, q, i)] = w * T.By[q][i][t];
    }
    for (int i = 0; i < KT::TOTAL; ++i) o->kc.v[i] = o->ktab_[i];
  }
  return o;
}


// number of state / control fields of the description (0 = the single-field default)
inline int desc_nfy(const pdenlp_desc& d) { return d.nfields_y > 0 ? d.nfields_y : 1; }
inline int desc_nfu(const pdenlp_desc& d) { return d.nu == 0 ? 0 : (d.nfields_u > 0 ? d.nfields_u : 1); }

// one line per instantiated (integrand, dim, local dofs y / u, quadrature points, nparam) case
#define PDENLP_SELECT_PROLOGUE(d)                                                                       \
  const int I = (d).integrand, dim = (d).dim, ny = (d).ny, nu = (d).nu, nq = (d).nq, np = (d).nparam;  \
  const int nfy = desc_nfy(d), nfu = desc_nfu(d);                                                       \
  (void)I; (void)dim; (void)ny; (void)nu; (void)nq; (void)np; (void)nfy; (void)nfu
// (each case exists twice: the uniform fast path, and the per-cell-geometry variant chosen when desc.cell_jinvT is given)
#define CASE(INT, DIMV, NYV, NUV, NQV, NPV, TYPE)                                                       \
  if (I == INT && dim == DIMV && ny == NYV && nu == NUV && nq == NQV && np == NPV) {                   \
    using EL = Elem<TYPE, NYV, NUV, NQV, false>;                                                        \
    using ELG = Elem<TYPE, NYV, NUV, NQV, true>;                                                        \
    if (nfy == EL::NFY && nfu == EL::NFU) {                                                             \
      if ((d).cell_jinvT != nullptr) return make_ops<ELG>(d);                                           \
      return make_ops<EL>(d);                                                                           \
    }                                                                                                   \
  }

// the translation units of the closed set (each returns nullptr when the case is not one of its own)
Ops* select_ops_part0(const pdenlp_desc& d);   // ops_membrane.cu : membrane and Poisson-Boltzmann, 1 point (+ PB with 4)
Ops* select_ops_part1(const pdenlp_desc& d);   // ops_membrane4.cu: membrane with 4 points, Poisson-Boltzmann with a Q2 state
Ops* select_ops_part2(const pdenlp_desc& d);   // ops_kappa.cu    : kappa-Poisson 2-D / 3-D
Ops* select_ops_part3(const pdenlp_desc& d);   // ops_misc.cu     : the remaining 2-D problems, Burgers, the ODE systems
Ops* select_ops_part4(const pdenlp_desc& d);   // ops_3d.cu       : membrane / Poisson-Boltzmann on 3-D Q1 meshes

}  // namespace host
}  // namespace pdenlp
