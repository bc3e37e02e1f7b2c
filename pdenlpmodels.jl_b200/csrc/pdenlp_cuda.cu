// pdenlp_cuda.cu — C ABI of libpdenlp_cuda.so (see include/pdenlp_cuda.h).
//
// Host side of the library: owns the device copy of the flattened model, picks the
// template instantiation for the (integrand, element) pair, launches the kernels of
// kernels.cuh on the caller's stream.  No CPU compute path exists: every entry point
// needs an sm_100 device.
#include "ops.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace pdenlp;

thread_local std::string pdenlp_g_err;  // shared with compaction.cu and the ops_*.cu units

using namespace pdenlp::host;

namespace {

// the closed set of (integrand, dim, ny, nu, nq, nparam, fields) instantiations lives in the ops_*.cu units
Ops* select_ops(const pdenlp_desc& d) {
  if (Ops* o = select_ops_part0(d)) return o;
  if (Ops* o = select_ops_part1(d)) return o;
  if (Ops* o = select_ops_part2(d)) return o;
  if (Ops* o = select_ops_part3(d)) return o;
  return select_ops_part4(d);
}

}  // namespace

struct pdenlp_model {
  pdenlp_desc d;  // scalar fields only are meaningful after create
  pdenlp_info info;
  Ops* ops = nullptr;
  DevModel dm;
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  bool inde = false;
  bool nofe = false;  // NoFETerm objective
  // owned device memory
  std::vector<void*> owned;
  // scratch
  double* partials = nullptr;
  double* scalar = nullptr;
  double* scalar_host = nullptr;  // pinned + mapped mirror of `scalar`, written by the objective kernel itself
  unsigned* ticket = nullptr;     // arrival counter of the last-CTA reductions
  int32_t npartials = 0;
  // host-memspace staging
  // (each staging buffer remembers its capacity: sv serves directions of length nvar (jprod!, hprod!) and ncon (jtprod!))
  struct Stage { double* p = nullptr; int64_t len = 0; };
  Stage sx, slam, sv;
  double* sout = nullptr;
  int64_t sout_len = 0;
  // timing
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  double selfcheck_err = 0.0;  // worst relative disagreement fast path vs generic path seen by the create-time self-check
  // free-dof ranges [lo, hi) (0-based, concatenated numbering) the cells of this handle touch: a PDENLP_MEM_HOST model
  // stages only these parts of x / lambda / v (a slab of a sharded model needs its own rows of the mesh, not the whole vector)
  int64_t yr[2] = {0, 0}, ur[2] = {0, 0};
  // deterministic mode (PDENLP_FLAG_DETERMINISTIC): greedy cell colouring, the cells of every colour as one list
  bool deterministic = false;
  int32_t* col_cells = nullptr;      // device: cells grouped by colour
  std::vector<int64_t> col_off;      // [ncolours + 1]
  // template tiles (kernels.cuh): classification lists (host copies for the self-check's sample), image scratch
  bool tpl_active = false;
  int32_t tpl_tile = -1;
  std::vector<int32_t> tpl_list_h, gen_list_h;
  double *tpl_jy = nullptr, *tpl_ju = nullptr, *tpl_h = nullptr;
  int64_t* zero_base = nullptr;
  bool tpl_j_valid = false, tpl_h_valid = false;  // cached images (integrands without kappa: independent of x)
  double tpl_h_ow = 0.0;
  int64_t tpl_builds = 0;
  // cell-granular templates: host copies of the chunk table (the self-check cuts its sample at a chunk start)
  bool cg_active = false;
  int64_t ntpl_cells = 0;
  // auxiliary stream of the cell-granular template path: the generic-cell kernel and the latency-bound kappa-block
  // kernels of jac_coord! / hess_coord! run there, concurrently with the chunk copy on the caller's stream
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // replay cache of the multi-launch coordinate callbacks of SMALL models (see graph_call)
  struct GraphEntry {
    int cb = 0; const void *x = nullptr, *lam = nullptr, *vals = nullptr; double ow = 0.0; cudaStream_t stream = nullptr;
    cudaGraphExec_t exec = nullptr; int64_t launches = 0; bool failed = false; uint64_t used = 0;
  };
  std::vector<GraphEntry> graphs;
  uint64_t graph_clock = 0;
  cudaStream_t cap = nullptr;  // private capture stream
  int64_t graph_replays = 0;
  cudaEvent_t ev_switch = nullptr;  // orders a new caller stream behind the previous one (pdenlp_set_stream)
  std::vector<int64_t> ch_cell0_h, ch_jy_h, ch_ju_h, ch_h_h;
  std::vector<int32_t> gc_cell_h;
  std::vector<int64_t> hb_jy, hb_ju, hb_h;  // host copies of the per-tile segment offsets (create time only)
};

namespace {

template <class T>
int dev_upload(pdenlp_model* m, const T* host, size_t n, T** out) {
  *out = nullptr;
  if (n == 0) { CU(cudaMalloc((void**)out, 16)); m->owned.push_back(*out); CU(cudaMemset(*out, 0, 16)); return PDENLP_OK; }
  CU(cudaMalloc((void**)out, n * sizeof(T)));
  m->owned.push_back(*out);
  CU(cudaMemcpy(*out, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return PDENLP_OK;
}

int ensure_stage(pdenlp_model* m, pdenlp_model::Stage* st, int64_t n) {
  if (st->p == nullptr || st->len < n) {  // grow: the same buffer may be asked for a longer vector later
    if (st->p) {
      CU(cudaStreamSynchronize(m->stream));  // an earlier callback may still read it
      cudaFree(st->p);
      m->owned.erase(std::find(m->owned.begin(), m->owned.end(), (void*)st->p));
      st->p = nullptr; st->len = 0;
    }
    CU(cudaMalloc((void**)&st->p, std::max<int64_t>(n, 2) * 8));
    m->owned.push_back(st->p);
    st->len = n;
  }
  return PDENLP_OK;
}
int ensure_out(pdenlp_model* m, int64_t n) {
  if (m->sout_len < n) {
    if (m->sout) { cudaFree(m->sout); m->owned.erase(std::find(m->owned.begin(), m->owned.end(), (void*)m->sout)); }
    m->sout = nullptr;
    CU(cudaMalloc((void**)&m->sout, std::max<int64_t>(n, 2) * 8));
    m->owned.push_back(m->sout);
    m->sout_len = n;
  }
  return PDENLP_OK;
}

// resolve an input vector to a device pointer (staging it for PDENLP_MEM_HOST).  layout: 0 = x-layout [kappa | y | u]
// (nvar entries), 1 = constraint layout (ncon entries, the free dofs of the test space = state ids).  Only the
// entries the handle's cells can read are copied (see yr / ur).
enum { LAYOUT_X = 0, LAYOUT_C = 1 };
int in_ptr(pdenlp_model* m, const double* user, int64_t n, pdenlp_model::Stage* stage, const double** dev, int layout = LAYOUT_X) {
  if (user == nullptr) { *dev = nullptr; return PDENLP_OK; }
  if (m->d.memspace == PDENLP_MEM_DEVICE) { *dev = user; return PDENLP_OK; }
  if (int rc = ensure_stage(m, stage, n)) return rc;
  auto part = [&](int64_t lo, int64_t hi) -> int {
    lo = std::max<int64_t>(lo, 0); hi = std::min<int64_t>(hi, n);
    if (hi > lo) CU(cudaMemcpyAsync(stage->p + lo, user + lo, (size_t)(hi - lo) * 8, cudaMemcpyHostToDevice, m->stream));
    return PDENLP_OK;
  };
  static const bool whole = getenv("PDENLP_STAGE_WHOLE") != nullptr;  // A/B switch
  if (whole) {
    if (int rc = part(0, n)) return rc;
  } else if (layout == LAYOUT_C) {
    if (int rc = part(m->yr[0], m->yr[1])) return rc;
  } else {
    const int64_t p = m->d.nparam, ny = m->d.nfree_y;
    if (int rc = part(0, p)) return rc;
    if (int rc = part(p + m->yr[0], p + m->yr[1])) return rc;
    if (int rc = part(p + ny + m->ur[0], p + ny + m->ur[1])) return rc;
  }
  *dev = stage->p;
  return PDENLP_OK;
}
int out_ptr(pdenlp_model* m, double* user, int64_t n, double** dev) {
  if (m->d.memspace == PDENLP_MEM_DEVICE) { *dev = user; return PDENLP_OK; }
  if (int rc = ensure_out(m, n)) return rc;
  *dev = m->sout;
  return PDENLP_OK;
}
int out_finish(pdenlp_model* m, double* user, int64_t n) {
  if (m->d.memspace == PDENLP_MEM_DEVICE) return PDENLP_OK;
  CU(cudaMemcpyAsync(user, m->sout, n * 8, cudaMemcpyDeviceToHost, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  return PDENLP_OK;
}
// PDENLP_MEM_HOST: a structurally zero tail of the output is not sent over PCIe; host threads fill it
// while the DMA engine copies the computed part [0, ncopy)
void host_zero_parallel(double* p, int64_t n) {
  if (n <= 0) return;
  static const unsigned env_t = getenv("PDENLP_HOST_THREADS") ? (unsigned)atoi(getenv("PDENLP_HOST_THREADS")) : 0u;
  unsigned T = env_t ? env_t : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  if (n < (int64_t)1 << 20) T = 1;
  std::vector<std::thread> th;
  for (unsigned t = 1; t < T; ++t) {
    const int64_t a = n * t / T, b = n * (t + 1) / T;
    th.emplace_back([=] { memset(p + a, 0, (size_t)(b - a) * 8); });
  }
  memset(p, 0, (size_t)(n / T) * 8);
  for (auto& t : th) t.join();
}
int out_finish_zero_tail(pdenlp_model* m, double* user, int64_t ncopy, int64_t nzero) {
  CU(cudaMemcpyAsync(user, m->sout, ncopy * 8, cudaMemcpyDeviceToHost, m->stream));
  host_zero_parallel(user + ncopy, nzero);
  CU(cudaStreamSynchronize(m->stream));
  return PDENLP_OK;
}

// zero fill of a (large) COO segment on the model's stream
int zero_fill(pdenlp_model* m, double* p, int64_t n, bool overlap_previous = false) {
  static const bool use_memset = getenv("PDENLP_ZERO_MEMSET") != nullptr;  // experiments only
  if (n <= 0) return PDENLP_OK;
  if (use_memset || n < 4096) {
    CU(cudaMemsetAsync(p, 0, (size_t)n * 8, m->stream));
    return PDENLP_OK;
  }
  const int64_t want = (n / 2 + 256 * 4 - 1) / (256 * 4);
  // one 16 KB chunk per CTA (measured on B200 at 3.05 GB: 0.40 ms; persistent grids of 8/16/32 CTAs per SM
  // 0.56/0.50/0.47 ms; cudaMemsetAsync 0.41 ms); PDENLP_FILL_CTAS = CTAs per SM of a persistent grid (experiments)
  static const int per_sm = getenv("PDENLP_FILL_CTAS") ? atoi(getenv("PDENLP_FILL_CTAS")) : 0;
  const int grid = (int)std::max<int64_t>(1, per_sm > 0 ? std::min<int64_t>(want, (int64_t)m->sms * per_sm) : std::min<int64_t>(want, 0x7fffffff));
  if (overlap_previous) {
    // programmatic dependent launch: the previous kernel on the stream (k_hess_coord, which writes a disjoint
    // block and has signalled launch_dependents) need not have finished
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = m->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, k_zero_fill, p, n));
  } else {
    k_zero_fill<<<grid, 256, 0, m->stream>>>(p, n);
    CU(cudaGetLastError());
  }
  m->launches += 1;
  return PDENLP_OK;
}

// Image of one template tile (kernels.cuh, "template tiles"): the ordinary coordinate kernel on a one-tile sub-model
// (the model's own arrays, offset to the template tile) writes the tile's run into model-owned scratch.  Integrands
// without algebraic unknowns have x-independent images, which are cached (the Hessian image per obj_weight).
int build_template(pdenlp_model* m, bool hess, const double* dx, double ow) {
  if (!m->tpl_active) return PDENLP_OK;
  if (hess ? !m->ops->tpl_hess_ok() : !m->ops->tpl_jac_ok()) return PDENLP_OK;
  const bool cacheable = m->d.nparam == 0;
  if (cacheable && (hess ? (m->tpl_h_valid && m->tpl_h_ow == ow) : m->tpl_j_valid)) return PDENLP_OK;
  DevModel dt = m->dm;
  const int64_t off = (int64_t)m->tpl_tile * 32;
  dt.dofs_y += off;
  dt.dofs_u += off;
  if (dt.coef) dt.coef += off * m->d.nq;  // [field][cell][q]; ncells_coef keeps the field stride
  dt.ncells = 32;
  dt.ntiles = 1;
  dt.base_jy = dt.base_ju = dt.base_h = m->zero_base;
  dt.tpl_list = dt.gen_list = nullptr;
  dt.ntpl = 0; dt.ngen = 1;
  dt.ch = nullptr; dt.nchunk = 0; dt.ngc = 0;
  Launch L{m->stream, m->sms, false, &m->launches};
  if (hess) {
    if (int rc = m->ops->hess(L, dt, dx, nullptr, ow, m->tpl_h, nullptr)) return rc;
    m->tpl_h_valid = cacheable; m->tpl_h_ow = ow;
  } else {
    if (int rc = m->ops->jac(L, dt, dx, m->tpl_jy, m->tpl_ju)) return rc;
    m->tpl_j_valid = cacheable;
  }
  m->tpl_builds += 1;
  return PDENLP_OK;
}

// Cell-granular templates (kernels.cuh): the device marks the template cells (one ballot mask per tile); the host cuts
// the slab into chunks of <= 32 consecutive template cells and a list of the remaining (generic) cells, each with the
// offset of its first entry in every COO segment.  The offsets are carried along from the per-tile exclusive scans
// (a template cell has RUN entries per segment, a generic cell its counted ones) and re-checked at every tile end.
int classify_cells(pdenlp_model* m, const pdenlp_desc& desc, unsigned long long sy, unsigned long long su) {
  const int64_t nt = m->dm.ntiles, nc = desc.ncells;
  if (nt < 4096) return PDENLP_OK;  // (small slabs: launch-bound either way; the tile kernels need no image)
  if (nc >= ((int64_t)1 << 31)) return PDENLP_OK;  // (the generic-cell list holds int32 cell numbers)
  if (desc.nq > 1 && m->dm.ktab == nullptr) return PDENLP_OK;
  uint32_t* dmask = nullptr;
  CU(cudaMalloc((void**)&dmask, (size_t)nt * 4));
  int rc = m->ops->cell_class(m->dm, sy, su, dmask, nullptr);
  std::vector<uint32_t> mask((size_t)nt);
  cudaError_t e = cudaMemcpy(mask.data(), dmask, (size_t)nt * 4, cudaMemcpyDeviceToHost);
  cudaFree(dmask);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(PDENLP_ECUDA, std::string("cell classification: ") + cudaGetErrorString(e));
  int64_t ntplc = 0, first_full = -1;
  for (int64_t t = 0; t < nt; ++t) {
    ntplc += __builtin_popcount(mask[t]);
    if (first_full < 0 && mask[t] == 0xffffffffu) first_full = t;
  }
  if (first_full < 0 || ntplc * 4 < nc) return PDENLP_OK;
  m->ntpl_cells = ntplc;
  const int ny = desc.ny, nu = desc.nu;
  const int64_t rjy = m->ops->run_jy(), rju = m->ops->run_ju(), rh = m->ops->run_h();
  std::vector<int64_t> gjy, gju, gh;
  std::vector<uint8_t> chn;
  // cells per chunk: 32 = one image length; up to 64 through the periodicity of the image (tuning knob)
  static const int chunk_cells = [] { const char* e = getenv("PDENLP_CHUNK_CELLS"); int v = e ? atoi(e) : 32; return v >= 1 && v <= 64 ? v : 32; }();
  int64_t run0 = -1, runn = 0, r_jy = 0, r_ju = 0, r_h = 0;  // the open run of template cells
  auto flush = [&]() {
    if (runn > 0) {
      m->ch_cell0_h.push_back(run0); m->ch_jy_h.push_back(r_jy); m->ch_ju_h.push_back(r_ju); m->ch_h_h.push_back(r_h);
      chn.push_back((uint8_t)runn);
    }
    runn = 0;
  };
  for (int64_t t = 0; t < nt; ++t) {
    int64_t ojy = m->hb_jy[t], oju = m->hb_ju[t], oh = m->hb_h[t];
    const uint32_t mk = mask[t];
    if (mk == 0xffffffffu && runn == 0 && chunk_cells == 32) {  // (fast path: a whole tile of template cells starting a chunk)
      m->ch_cell0_h.push_back(t * 32); m->ch_jy_h.push_back(ojy); m->ch_ju_h.push_back(oju); m->ch_h_h.push_back(oh);
      chn.push_back(32);
      continue;
    }
    for (int l = 0; l < 32; ++l) {
      const int64_t c = t * 32 + l;
      if (c >= nc) break;
      if ((mk >> l) & 1u) {
        if (runn == 0) { run0 = c; r_jy = ojy; r_ju = oju; r_h = oh; }
        ++runn;
        ojy += rjy; oju += rju; oh += rh;
        if (runn == chunk_cells) flush();
      } else {
        flush();
        int fy = 0, fu = 0;
        for (int a = 0; a < ny; ++a) fy += desc.cell_dofs_y[c * ny + a] > 0;
        for (int b = 0; b < nu; ++b) fu += desc.cell_dofs_u[c * nu + b] > 0;
        m->gc_cell_h.push_back((int32_t)c); gjy.push_back(ojy); gju.push_back(oju); gh.push_back(oh);
        ojy += (int64_t)fy * fy; oju += (int64_t)fy * fu;
        oh += (int64_t)fy * (fy + 1) / 2 + (int64_t)fu * fy + (int64_t)fu * (fu + 1) / 2;  // (distinct ids: checked by k_count)
      }
    }
    if (ojy != m->hb_jy[t + 1] || oju != m->hb_ju[t + 1] || oh != m->hb_h[t + 1])
      return fail(PDENLP_EINVAL, "internal: per-cell segment offsets disagree with the per-tile scans");
  }
  flush();
  if (m->ch_cell0_h.size() >= ((size_t)1 << 31) || m->gc_cell_h.size() >= ((size_t)1 << 31)) return PDENLP_OK;
  int64_t* lp; int32_t* ip;
  {
    std::vector<ChunkRec> recs(m->ch_cell0_h.size());
    for (size_t k = 0; k < recs.size(); ++k) recs[k] = ChunkRec{m->ch_jy_h[k], m->ch_ju_h[k], m->ch_h_h[k], (long long)chn[k]};
    ChunkRec* rp = nullptr;
    if (int r = dev_upload(m, recs.data(), recs.size(), &rp)) return r;
    m->dm.ch = rp;
  }
  if (int r = dev_upload(m, m->gc_cell_h.data(), m->gc_cell_h.size(), &ip)) return r;
  m->dm.gc_cell = ip;
  if (int r = dev_upload(m, gjy.data(), gjy.size(), &lp)) return r;
  m->dm.gc_jy = lp;
  if (int r = dev_upload(m, gju.data(), gju.size(), &lp)) return r;
  m->dm.gc_ju = lp;
  if (int r = dev_upload(m, gh.data(), gh.size(), &lp)) return r;
  m->dm.gc_h = lp;
  m->dm.nchunk = (int32_t)m->ch_cell0_h.size();
  m->dm.ngc = (int32_t)m->gc_cell_h.size();
  m->tpl_tile = (int32_t)first_full;
  auto scratch = [&](double** p, int run) -> int {
    CU(cudaMalloc((void**)p, (size_t)std::max(32 * run, 2) * 8));
    m->owned.push_back(*p);
    return PDENLP_OK;
  };
  if (int r = scratch(&m->tpl_jy, (int)rjy)) return r;
  if (int r = scratch(&m->tpl_ju, (int)rju)) return r;
  if (int r = scratch(&m->tpl_h, (int)rh)) return r;
  m->dm.tpl_jy = m->tpl_jy; m->dm.tpl_ju = m->tpl_ju; m->dm.tpl_h = m->tpl_h;
  const int64_t z[2] = {0, 0};
  if (int r = dev_upload(m, z, 2, &m->zero_base)) return r;
  m->tpl_active = true;
  m->cg_active = true;
  return PDENLP_OK;
}

// Classify the tiles once at create time: the template pattern is the id-order signature of an all-free cell near the
// middle of the slab; a tile is a template tile iff all its 32 cells carry it.  Used only when it pays (>= 25 % of the tiles).
int classify_tiles(pdenlp_model* m, const pdenlp_desc& desc) {
  static const bool off = getenv("PDENLP_NO_TEMPLATE") != nullptr;  // experiments / A-B runs
  if (off || (desc.flags & PDENLP_FLAG_NO_TEMPLATES) || !(m->ops->tpl_jac_ok() || m->ops->tpl_hess_ok())) return PDENLP_OK;
  const int64_t nc = desc.ncells;
  const int ny = desc.ny, nu = desc.nu;
  auto signature = [&](int64_t c, unsigned long long& sy, unsigned long long& su) -> bool {
    sy = su = 0ull;
    for (int a = 0; a < ny; ++a) {
      const int32_t ga = desc.cell_dofs_y[c * ny + a];
      if (ga <= 0) return false;
      unsigned r = 0;
      for (int b = 0; b < ny; ++b) r += desc.cell_dofs_y[c * ny + b] < ga;
      sy |= (unsigned long long)(r & 15u) << (4 * (a & 15));
    }
    for (int a = 0; a < nu; ++a) {
      const int32_t ga = desc.cell_dofs_u[c * nu + a];
      if (ga <= 0) return false;
      unsigned r = 0;
      for (int b = 0; b < nu; ++b) r += desc.cell_dofs_u[c * nu + b] < ga;
      su |= (unsigned long long)(r & 15u) << (4 * (a & 15));
    }
    return true;
  };
  unsigned long long sy = 0, su = 0;
  bool found = false;
  for (int64_t c = nc / 2; c < nc && c < nc / 2 + 65536 && !found; ++c) found = signature(c, sy, su);
  if (!found) return PDENLP_OK;
  const int64_t nt = m->dm.ntiles;
  static const bool tile_mode = [] { const char* e = getenv("PDENLP_TPL_MODE"); return e && std::string(e) == "tile"; }();
  if (!tile_mode) return classify_cells(m, desc, sy, su);
  uint8_t* dcls = nullptr;
  CU(cudaMalloc((void**)&dcls, (size_t)nt));
  int rc = m->ops->tile_class(m->dm, sy, su, dcls, nullptr);
  std::vector<uint8_t> cls((size_t)nt);
  cudaError_t e = cudaMemcpy(cls.data(), dcls, (size_t)nt, cudaMemcpyDeviceToHost);
  cudaFree(dcls);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(PDENLP_ECUDA, std::string("tile classification: ") + cudaGetErrorString(e));
  for (int64_t t = 0; t < nt; ++t) (cls[t] ? m->tpl_list_h : m->gen_list_h).push_back((int32_t)t);
  // (small slabs: the per-CTA image load would cost more than it saves)
  if ((int64_t)m->tpl_list_h.size() * 4 < nt || nt < 4096) { m->tpl_list_h.clear(); m->gen_list_h.clear(); return PDENLP_OK; }
  m->tpl_tile = m->tpl_list_h[0];
  int32_t* ip;
  if (int r = dev_upload(m, m->tpl_list_h.data(), m->tpl_list_h.size(), &ip)) return r;
  m->dm.tpl_list = ip;
  if (int r = dev_upload(m, m->gen_list_h.data(), m->gen_list_h.size(), &ip)) return r;
  m->dm.gen_list = ip;
  m->dm.ntpl = (int32_t)m->tpl_list_h.size();
  m->dm.ngen = (int32_t)m->gen_list_h.size();
  auto scratch = [&](double** p, int run) -> int {
    CU(cudaMalloc((void**)p, (size_t)std::max(32 * run, 2) * 8));
    m->owned.push_back(*p);
    return PDENLP_OK;
  };
  if (int r = scratch(&m->tpl_jy, m->ops->run_jy())) return r;
  if (int r = scratch(&m->tpl_ju, m->ops->run_ju())) return r;
  if (int r = scratch(&m->tpl_h, m->ops->run_h())) return r;
  m->dm.tpl_jy = m->tpl_jy; m->dm.tpl_ju = m->tpl_ju; m->dm.tpl_h = m->tpl_h;
  const int64_t z[2] = {0, 0};
  if (int r = dev_upload(m, z, 2, &m->zero_base)) return r;
  m->tpl_active = true;
  return PDENLP_OK;
}

// Greedy colouring of the cells for the deterministic mode: a cell takes the lowest colour none of its free dofs has
// seen yet (one 64-bit mask per dof), so two cells of a colour never share a free dof and a launch over one colour can
// accumulate with plain stores.  Structured meshes need 2^dim colours.
int colour_cells(pdenlp_model* m, const pdenlp_desc& desc, const std::vector<int64_t>& offf, int nfy, int nfu) {
  const int64_t nc = desc.ncells;
  const int ny = nfy * desc.ny, nu = nfu * desc.nu;
  std::vector<uint64_t> masky((size_t)desc.nfree_y, 0), masku((size_t)std::max<int64_t>(desc.nfree_u, 1), 0);
  std::vector<uint8_t> colour((size_t)nc);
  std::vector<int64_t> count(64, 0);
  int ncol = 0;
  for (int64_t c = 0; c < nc; ++c) {
    uint64_t used = 0;
    for (int a = 0; a < ny; ++a) {
      const int64_t g = desc.cell_dofs_y[c * ny + a];
      if (g > 0) used |= masky[(size_t)(g + offf[a / desc.ny] - 1)];
    }
    for (int b = 0; b < nu; ++b) {
      const int64_t g = desc.cell_dofs_u[c * nu + b];
      if (g > 0) used |= masku[(size_t)(g + offf[nfy + b / desc.nu] - 1)];
    }
    if (used == ~0ull) return fail(PDENLP_EUNSUPPORTED, "deterministic mode: more than 64 colours needed");
    const int k = __builtin_ctzll(~used);
    colour[(size_t)c] = (uint8_t)k;
    ncol = std::max(ncol, k + 1);
    ++count[k];
    const uint64_t bit = 1ull << k;
    for (int a = 0; a < ny; ++a) {
      const int64_t g = desc.cell_dofs_y[c * ny + a];
      if (g > 0) masky[(size_t)(g + offf[a / desc.ny] - 1)] |= bit;
    }
    for (int b = 0; b < nu; ++b) {
      const int64_t g = desc.cell_dofs_u[c * nu + b];
      if (g > 0) masku[(size_t)(g + offf[nfy + b / desc.nu] - 1)] |= bit;
    }
  }
  m->col_off.assign(ncol + 1, 0);
  for (int k = 0; k < ncol; ++k) m->col_off[k + 1] = m->col_off[k] + count[k];
  std::vector<int64_t> pos(m->col_off.begin(), m->col_off.end() - 1);
  std::vector<int32_t> cells((size_t)nc);
  for (int64_t c = 0; c < nc; ++c) cells[(size_t)pos[colour[(size_t)c]]++] = (int32_t)c;  // ascending inside a colour
  if (int rc = dev_upload(m, cells.data(), cells.size(), &m->col_cells)) return rc;
  m->deterministic = true;
  return PDENLP_OK;
}
// run `launch(DevModel of one colour)` for every colour (deterministic mode) or once on the whole model
template <class F>
int for_each_colour(pdenlp_model* m, F&& launch) {
  if (!m->deterministic) return launch(m->dm, false);
  for (size_t k = 0; k + 1 < m->col_off.size(); ++k) {
    DevModel d = m->dm;
    d.cell_list = m->col_cells + m->col_off[k];
    d.nlist = m->col_off[k + 1] - m->col_off[k];
    d.ntiles = (int32_t)((d.nlist + 31) / 32);
    if (d.nlist == 0) continue;
    if (int rc = launch(d, true)) return rc;
  }
  return PDENLP_OK;
}

// Create-time self-check of the trait-selected shortcuts (RES_FE_HESSIAN_ZERO: the constraint FE Hessian block is a
// zero fill; FE_DERIVS_POINT_INDEPENDENT: the coordinate kernels evaluate one quadrature point and use the
// pre-integrated reference tensors; both: the reference-tensor vector / kappa-block kernels of kvector.cuh).  The traits
// are hand-set per integrand, so every model that uses one runs BOTH code paths once on a sample of its own cells with
// pseudo-random x, lambda, v and refuses to be created if they differ: a wrong flag cannot silently change results.
int self_check(pdenlp_model* m, double* worst_out) {
  static const bool skip = getenv("PDENLP_NO_SELFCHECK") != nullptr;
  if (worst_out) *worst_out = 0.0;
  if (skip || !m->ops->has_shortcuts()) return PDENLP_OK;
  const pdenlp_info& I = m->info;
  // (the sample is a prefix of the slab; it is extended to contain the first template tiles when there are any)
  int64_t ns = std::min<int64_t>(m->dm.ncells, 2048);
  if (m->tpl_active) ns = std::min<int64_t>(m->dm.ncells, std::max<int64_t>(ns, ((int64_t)m->tpl_tile + 24) * 32));
  DevModel ds = m->dm;
  int64_t njy = 0, nju = 0, nh = 0;  // sizes of the sample's COO runs
  bool sizes_known = false;
  if (m->cg_active) {  // the sample ends where a chunk starts (chunks are not aligned to tiles)
    const size_t k = std::lower_bound(m->ch_cell0_h.begin(), m->ch_cell0_h.end(), ns) - m->ch_cell0_h.begin();
    if (k < m->ch_cell0_h.size()) {
      ns = m->ch_cell0_h[k];
      njy = m->ch_jy_h[k]; nju = m->ch_ju_h[k]; nh = m->ch_h_h[k];
      sizes_known = true;
    } else {
      ns = m->dm.ncells;
    }
    ds.nchunk = (int32_t)k;
    ds.ngc = (int32_t)(std::lower_bound(m->gc_cell_h.begin(), m->gc_cell_h.end(), (int32_t)std::min<int64_t>(ns, INT32_MAX)) - m->gc_cell_h.begin());
  }
  ds.ncells = ns;
  ds.ntiles = (int32_t)((ns + 31) / 32);
  if (m->tpl_active && !m->cg_active) {
    ds.ntpl = (int32_t)(std::lower_bound(m->tpl_list_h.begin(), m->tpl_list_h.end(), ds.ntiles) - m->tpl_list_h.begin());
    ds.ngen = (int32_t)(std::lower_bound(m->gen_list_h.begin(), m->gen_list_h.end(), ds.ntiles) - m->gen_list_h.begin());
  }
  if (!sizes_known) {  // = segment offsets of the first tile after the sample
    CU(cudaMemcpy(&njy, m->dm.base_jy + ds.ntiles, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&nju, m->dm.base_ju + ds.ntiles, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&nh, m->dm.base_h + ds.ntiles, 8, cudaMemcpyDeviceToHost));
  }
  const int64_t nk = std::max<int64_t>(I.nnzj_k, I.nnzh_k_con);
  const int64_t nout = std::max<int64_t>({njy + nju, 2 * nh, I.nvar, nk, (int64_t)2});
  double *x = nullptr, *lam = nullptr, *v = nullptr, *a = nullptr, *b = nullptr;
  unsigned long long* res = nullptr;
  std::vector<void*> tmp;
  auto cleanup = [&]() { for (void* p : tmp) cudaFree(p); };
  auto alloc = [&](double** p, int64_t n) -> int {
    CU(cudaMalloc((void**)p, (size_t)std::max<int64_t>(n, 2) * 8));
    tmp.push_back(*p);
    return PDENLP_OK;
  };
  int rc = PDENLP_OK;
  if ((rc = alloc(&x, I.nvar)) || (rc = alloc(&lam, I.ncon)) || (rc = alloc(&v, I.nvar)) || (rc = alloc(&a, nout)) ||
      (rc = alloc(&b, nout)) || (rc = alloc((double**)&res, 2))) { cleanup(); return rc; }
  cudaStream_t st = m->stream;
  k_fill_random<<<1024, 256, 0, st>>>(x, I.nvar, 1998, -1.0, 1.0);
  if (m->d.nparam > 0) k_fill_random<<<1, 32, 0, st>>>(x, m->d.nparam, 7, 0.9, 1.1);
  k_fill_random<<<1024, 256, 0, st>>>(lam, I.ncon, 1999, -1.0, 1.0);
  k_fill_random<<<1024, 256, 0, st>>>(v, I.nvar, 2000, -1.0, 1.0);
  double worst = 0.0;
  std::string worst_op;
  // run(generic, out) fills `out` (n entries compared) through one of the two paths
  auto compare = [&](const char* what, int64_t n, auto&& run) -> int {
    for (int pass = 0; pass < 2; ++pass) {
      double* o = pass ? b : a;
      CU(cudaMemsetAsync(o, 0, (size_t)n * 8, st));
      Launch L{st, m->sms, pass == 1, nullptr};
      DevModel d = ds;
      d.notrait = pass | ds.notrait;
      if (int r = run(L, d, o)) return r;
    }
    CU(cudaMemsetAsync(res, 0, 16, st));
    k_max_diff<<<256, 256, 0, st>>>(a, b, n, res);
    unsigned long long h[2];
    CU(cudaMemcpyAsync(h, res, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    double md, mb;
    memcpy(&md, &h[0], 8); memcpy(&mb, &h[1], 8);
    const double err = md != md ? 1.0 : (mb > 0.0 ? md / mb : (md > 0.0 ? 1.0 : 0.0));
    if (err > worst || worst_op.empty()) { worst = std::max(worst, err); worst_op = what; }
    return PDENLP_OK;
  };
  Ops* ops = m->ops;
  const bool coef = m->dm.coef != nullptr, need = ops->needs_coef_deriv();
  if (!rc && (coef || !need) && I.ncon > 0) {
    rc = build_template(m, false, x, 1.0);
    if (!rc) rc = compare("jac_coord", njy + nju, [&](const Launch& L, const DevModel& d, double* o) { return ops->jac(L, d, x, o, o + njy); });
  }
  if (!rc && (coef || !need) && !m->nofe) rc = build_template(m, true, x, 0.37);
  if (!rc && (coef || !need))
    rc = compare("hess_coord", (m->nofe ? 1 : 2) * nh, [&](const Launch& L, const DevModel& d, double* o) {
      // generic pass: both FE blocks are computed; shortcut pass: what pdenlp_hess_coord does (zero fill of a
      // structurally zero constraint block — `o` is already zeroed)
      double* vo = m->nofe ? nullptr : o;
      double* vc = o + (m->nofe ? 0 : nh);
      const bool cons_zero = ops->res_fe_hessian_zero() && !L.generic;
      if (I.ncon == 0 || cons_zero) vc = nullptr;
      if (!vo && !vc) return (int)PDENLP_OK;
      return ops->hess(L, d, x, lam, 0.37, vo, vc);
    });
  if (!rc && coef && I.ncon > 0) {
    int nb = 0;
    rc = compare("cons", I.ncon, [&](const Launch& L, const DevModel& d, double* o) { return ops->vec(L, d, OP_CONS, x, nullptr, nullptr, 1.0, o, m->partials, &nb); });
    if (!rc) rc = compare("jprod", I.ncon, [&](const Launch& L, const DevModel& d, double* o) { return ops->vec(L, d, OP_JPROD, x, nullptr, v, 1.0, o, m->partials, &nb); });
    if (!rc) rc = compare("jtprod", I.nvar, [&](const Launch& L, const DevModel& d, double* o) { return ops->vec(L, d, OP_JTPROD, x, nullptr, lam, 1.0, o, m->partials, &nb); });
    if (!rc) rc = compare("hprod", I.nvar, [&](const Launch& L, const DevModel& d, double* o) { return ops->vec(L, d, OP_HPROD, x, lam, v, 0.37, o, m->partials, &nb); });
    if (!rc && I.nnzj_k > 0)
      rc = compare("jac kappa-block", I.nnzj_k, [&](const Launch& L, const DevModel& d, double* o) { return ops->jac_kappa(L, d, x, o, m->partials); });
    if (!rc && I.nnzh_k_con > 0)
      rc = compare("hess kappa-block", I.nnzh_k_con, [&](const Launch& L, const DevModel& d, double* o) { return ops->hess_kappa(L, d, x, lam, 1.0, 1, I.nvar, o, m->partials); });
  } else if (!rc && !m->nofe) {  // no coefficient fields: the objective part of hprod! is still checkable
    int nb = 0;
    if (!need) rc = compare("hprod(obj)", I.nvar, [&](const Launch& L, const DevModel& d, double* o) { return ops->vec(L, d, OP_HPROD, x, nullptr, v, 0.37, o, m->partials, &nb); });
  }
  cleanup();
  if (rc) return rc;
  if (worst_out) *worst_out = worst;
  if (worst > 1e-12)
    return fail(PDENLP_EINVAL, "self-check failed: a trait-selected fast path (" + worst_op + ") disagrees with the generic kernels, relative error " + std::to_string(worst));
  return PDENLP_OK;
}

// ---- create-time device helpers: SoA transpose of the cell -> dof maps and the segment-offset scans
struct FieldTab { long long nf[16], nd[16], of[16], od[16]; };  // per field: free / Dirichlet sizes and their offsets in the space
// soa[a][cell] = concatenated id of local dof a (validated against the field sizes); res[0] = min free id - 1,
// res[1] = max free id (0-based range [lo, hi) of the free dofs the cells touch), res[2] = 1 on a bad id
__global__ void __launch_bounds__(256) k_transpose_ids(const int32_t* __restrict__ aos, int64_t nc, int64_t ldc, int nloc, int nloc_field,
                                                        int f0, FieldTab ft, int32_t* __restrict__ soa, unsigned long long* __restrict__ res) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ldc) return;
  unsigned long long lo = ~0ull, hi = 0ull;
  bool bad = false;
  for (int a = 0; a < nloc; ++a) {
    int32_t o = -1;
    if (c < nc) {
      const int f = f0 + a / nloc_field;
      const long long g = aos[c * nloc + a];
      if (g == 0 || g > ft.nf[f] || -g > ft.nd[f]) bad = true;
      o = (int32_t)(g > 0 ? g + ft.of[f] : g - ft.od[f]);
      if (g > 0) { lo = min(lo, (unsigned long long)(g + ft.of[f] - 1)); hi = max(hi, (unsigned long long)(g + ft.of[f])); }
    }
    soa[(int64_t)a * ldc + c] = o;
  }
  for (int d = 16; d > 0; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    const int ob = __shfl_xor_sync(0xffffffffu, (int)bad, d);  // (unconditional: no short-circuit around a shuffle)
    bad = bad || ob != 0;
  }
  if ((threadIdx.x & 31) == 0) {
    if (hi > 0) { atomicMin(res + 0, lo); atomicMax(res + 1, hi); }
    if (bad) atomicExch(res + 2, 1ull);
  }
}
// exclusive scans of the three per-tile counts (AoS [ntiles][3], int32) -> base[k][0..ntiles] (int64); one CTA per
// segment kind, decoupled in chunks of 1024 tiles carried serially (ntiles <= 2^31 / 32: a few thousand chunks)
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* __restrict__ counts, int64_t nt, int64_t* __restrict__ b0,
                                                       int64_t* __restrict__ b1, int64_t* __restrict__ b2) {
  __shared__ long long wsum[32];
  __shared__ long long carry_s;
  int64_t* const base = blockIdx.x == 0 ? b0 : (blockIdx.x == 1 ? b1 : b2);
  const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t t0 = 0; t0 < nt; t0 += 1024) {
    const int64_t t = t0 + threadIdx.x;
    const long long v = t < nt ? (long long)counts[t * 3 + k] : 0ll;
    long long incl = v;
    for (int d = 1; d < 32; d <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long w = wsum[lane];
      for (int d = 1; d < 32; d <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += u; }
      wsum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long excl = carry + (warp > 0 ? wsum[warp - 1] : 0ll) + incl - v;
    if (t < nt) base[t] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) base[nt] = carry_s;
}

struct Fork {  // part of a callback on the model's auxiliary stream: forked from / joined back into the caller's stream
  pdenlp_model* m;
  bool on = false;
  Fork(pdenlp_model* mm, bool want) : m(mm) {
    static const bool off = getenv("PDENLP_NO_AUX_STREAM") != nullptr;  // experiments / A-B runs
    if (want && !off && m->aux &&
        cudaEventRecord(m->ev_fork, m->stream) == cudaSuccess && cudaStreamWaitEvent(m->aux, m->ev_fork, 0) == cudaSuccess)
      on = true;
  }
  cudaStream_t side() const { return on ? m->aux : m->stream; }
  cudaStream_t aux() const { return on ? m->aux : nullptr; }
  // CTAs per SM of the Jacobian kappa-block kernel while it shares the SMs with the chunk copy (0: its default grid).
  // Measured on config 5 (jac_coord! ms, auxiliary stream at high priority): default grid 0.454, 2 per SM 0.438, 1 per SM
  // 0.429 — a small resident footprint leaves the registers to the copy CTAs; the Hessian kappa kernels keep theirs
  // (hess_coord! 0.759 against 0.767 / 0.778).
  int bg_ctas() const {
    static const int v = [] { const char* e = getenv("PDENLP_KAPPA_BG_CTAS"); return e ? atoi(e) : 1; }();
    return (on && m->cg_active) ? v : 0;
  }
  void join() {
    if (on) { cudaEventRecord(m->ev_join, m->aux); cudaStreamWaitEvent(m->stream, m->ev_join, 0); }
    on = false;
  }
  ~Fork() { join(); }
};

struct Scope {  // device + timing bracket of one callback
  pdenlp_model* m;
  int prev = -1;
  explicit Scope(pdenlp_model* mm) : m(mm) {
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    if (m->timing) { cudaEventRecord(m->ev0, m->stream); }
  }
  ~Scope() {
    if (m->timing) { cudaEventRecord(m->ev1, m->stream); m->ev_valid = true; }
    if (prev != m->device && prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

extern "C" {

const char* pdenlp_last_error(void) { return pdenlp_g_err.c_str(); }
int pdenlp_abi_version(void) { return PDENLP_ABI_VERSION; }
#ifndef PDENLP_BUILD_ID
#define PDENLP_BUILD_ID "unknown"
#endif
// (the marker prefix lets build.py read the id from the file without loading it)
__attribute__((used, visibility("default"))) extern const char pdenlp_build_marker[] = "PDENLP_BUILD_ID=" PDENLP_BUILD_ID ";";
const char* pdenlp_build_id(void) {
  static const std::string id(pdenlp_build_marker + 16, sizeof(pdenlp_build_marker) - 16 - 2);
  return id.c_str();
}

int pdenlp_create(const pdenlp_desc* desc, pdenlp_model** out) {
  if (!desc || !out) return fail(PDENLP_EINVAL, "null argument");
  *out = nullptr;
  if (desc->abi_version != PDENLP_ABI_VERSION) return fail(PDENLP_EINVAL, "abi_version mismatch");
  if (desc->ncells <= 0 || desc->nfree_y <= 0) return fail(PDENLP_EINVAL, "empty model");
  if (!desc->cell_dofs_y || !desc->phi_y || !desc->grad_y || !desc->wdet) return fail(PDENLP_EINVAL, "missing arrays");
  if (desc->nu > 0 && (!desc->cell_dofs_u || !desc->phi_u)) return fail(PDENLP_EINVAL, "missing control arrays");
  if (desc->ncells + 31 >= ((int64_t)1 << 36)) return fail(PDENLP_EINVAL, "too many cells");
  if ((desc->cell_jinvT != nullptr) != (desc->cell_detj != nullptr)) return fail(PDENLP_EINVAL, "cell_jinvT and cell_detj must be given together");
  if (desc->cell_jinvT && desc->geom_nq != 1 && desc->geom_nq != desc->nq) return fail(PDENLP_EINVAL, "geom_nq must be 1 (affine cells) or nq");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(PDENLP_ECUDA, "no CUDA device (libpdenlp_cuda has no CPU fallback)");
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(PDENLP_EINVAL, "bad device ordinal");
  int prev = 0;
  CU(cudaGetDevice(&prev));
  CU(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) { cudaSetDevice(prev); return fail(PDENLP_ECUDA, "libpdenlp_cuda is built for sm_100a only"); }

  Ops* ops = select_ops(*desc);
  if (!ops) { cudaSetDevice(prev); return fail(PDENLP_EUNSUPPORTED, "(integrand, dim, ny, nu, nq, nparam) is outside the closed set of instantiated kernels"); }

  auto* m = new pdenlp_model();
  m->d = *desc;
  m->ops = ops;
  m->device = dev;
  m->sms = prop.multiProcessorCount;
  m->nofe = ops->no_fe_obj();
  // MixedEnergyFETerm(..., inde = true) and NoFETerm: objective kappa-block = p x p lower triangle
  m->inde = desc->integrand == PDENLP_KAPPA_POISSON || m->nofe;
  int rc = PDENLP_OK;
  auto bail = [&](int code) { pdenlp_destroy(m); cudaSetDevice(prev); return code; };

  const int64_t nc = desc->ncells, ldc = (nc + 31) / 32 * 32;
  const int nfy = desc_nfy(*desc), nfu = desc_nfu(*desc);
  const int ny = nfy * desc->ny, nu = nfu * desc->nu;  // local dofs of the whole state / control space
  // per-field sizes -> offsets of the concatenated (ConsecutiveMultiFieldStyle) numbering of each space
  std::vector<int64_t> fnf(nfy + nfu), fnd(nfy + nfu), offf(nfy + nfu), offd(nfy + nfu);
  if (nfy > 1 || nfu > 1) {
    if (!desc->field_nfree || !desc->field_ndir) return bail(fail(PDENLP_EINVAL, "multi-field description without field_nfree/field_ndir"));
    for (int f = 0; f < nfy + nfu; ++f) { fnf[f] = desc->field_nfree[f]; fnd[f] = desc->field_ndir[f]; }
  } else {
    fnf[0] = desc->nfree_y; fnd[0] = desc->ndir_y;
    if (nfu) { fnf[1] = desc->nfree_u; fnd[1] = desc->ndir_u; }
  }
  {
    int64_t sf = 0, sd = 0;
    for (int f = 0; f < nfy; ++f) { offf[f] = sf; offd[f] = sd; sf += fnf[f]; sd += fnd[f]; }
    if (sf != desc->nfree_y || sd != desc->ndir_y) return bail(fail(PDENLP_EINVAL, "state field sizes do not add up to nfree_y/ndir_y"));
    sf = sd = 0;
    for (int f = nfy; f < nfy + nfu; ++f) { offf[f] = sf; offd[f] = sd; sf += fnf[f]; sd += fnd[f]; }
    if (nfu && (sf != desc->nfree_u || sd != desc->ndir_u)) return bail(fail(PDENLP_EINVAL, "control field sizes do not add up to nfree_u/ndir_u"));
  }
  // SoA transpose of the cell->dof maps on the device (the raw maps are uploaded as they are), validating the per-field
  // ids against the declared sizes and renumbering them to the concatenated ids of the space
  if (nfy + nfu > 16) return bail(fail(PDENLP_EUNSUPPORTED, "more than 16 fields"));
  FieldTab ft;
  for (int f = 0; f < 16; ++f) { ft.nf[f] = ft.nd[f] = ft.of[f] = ft.od[f] = 0; }
  for (int f = 0; f < nfy + nfu; ++f) { ft.nf[f] = fnf[f]; ft.nd[f] = fnd[f]; ft.of[f] = offf[f]; ft.od[f] = offd[f]; }
  auto upload_ids = [&](const int32_t* ids, int nloc_total, int nloc_field, int f0, const char* what, const int32_t** out_ptr, int64_t* range) -> int {
    int32_t* soa = nullptr;
    CU(cudaMalloc((void**)&soa, (size_t)ldc * std::max(nloc_total, 1) * 4));
    m->owned.push_back(soa);
    *out_ptr = soa;
    range[0] = range[1] = 0;
    if (nloc_total == 0) { CU(cudaMemset(soa, 0xff, (size_t)ldc * 4)); return PDENLP_OK; }
    int32_t* raw = nullptr;
    unsigned long long* res = nullptr;
    if (cudaMalloc((void**)&raw, (size_t)nc * nloc_total * 4) != cudaSuccess || cudaMalloc((void**)&res, 24) != cudaSuccess) {
      cudaFree(raw);
      return fail(PDENLP_ENOMEM, "cudaMalloc failed");
    }
    const unsigned long long init[3] = {~0ull, 0ull, 0ull};
    unsigned long long h[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpy(raw, ids, (size_t)nc * nloc_total * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(res, init, 24, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      k_transpose_ids<<<(unsigned)((ldc + 255) / 256), 256>>>(raw, nc, ldc, nloc_total, nloc_field, f0, ft, soa, res);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost);
    cudaFree(raw);
    cudaFree(res);
    if (e != cudaSuccess) return fail(PDENLP_ECUDA, std::string("dof map transpose: ") + cudaGetErrorString(e));
    if (h[2]) return fail(PDENLP_EINVAL, std::string(what) + " dof id out of range");
    if (h[1] > 0) { range[0] = (int64_t)h[0]; range[1] = (int64_t)h[1]; }
    return PDENLP_OK;
  };
  if (desc->nfree_y >= ((int64_t)1 << 31) || desc->nfree_u >= ((int64_t)1 << 31)) return bail(fail(PDENLP_EINVAL, "too many dofs for int32 ids"));
  if ((rc = upload_ids(desc->cell_dofs_y, ny, desc->ny, 0, "state", &m->dm.dofs_y, m->yr))) return bail(rc);
  if ((rc = upload_ids(desc->cell_dofs_u, nu, std::max(desc->nu, 1), nfy, "control", &m->dm.dofs_u, m->ur))) return bail(rc);
  double* dp;
  if ((rc = dev_upload(m, desc->dir_y, (size_t)desc->ndir_y, &dp))) return bail(rc);
  m->dm.dir_y = dp;
  if ((rc = dev_upload(m, desc->dir_u, (size_t)desc->ndir_u, &dp))) return bail(rc);
  m->dm.dir_u = dp;
  if (desc->ncoef > 0 && desc->coef) {
    if (desc->ncoef != 2) return bail(fail(PDENLP_EINVAL, "ncoef must be 0 or 2"));
    if ((rc = dev_upload(m, desc->coef, (size_t)2 * nc * desc->nq, &dp))) return bail(rc);
    m->dm.coef = dp;
  } else {
    m->dm.coef = nullptr;
  }
  m->dm.ncells = nc;
  m->dm.ncells_coef = nc;
  m->dm.notrait = 0;
  m->dm.ldc = ldc;
  m->dm.nfree_y = desc->nfree_y;
  m->dm.nfree_u = desc->nfree_u;
  m->dm.ntiles = (int32_t)(ldc / 32);
  m->dm.nparam = desc->nparam;
  m->dm.base_jy = m->dm.base_ju = m->dm.base_h = nullptr;
  m->dm.ktab = nullptr;
  m->dm.cell_list = nullptr; m->dm.nlist = 0;
  m->dm.tpl_list = m->dm.gen_list = nullptr;
  m->dm.ntpl = 0; m->dm.ngen = 0;
  m->dm.tpl_jy = m->dm.tpl_ju = m->dm.tpl_h = nullptr;
  m->dm.ch = nullptr;
  m->dm.gc_cell = nullptr;
  m->dm.gc_jy = m->dm.gc_ju = m->dm.gc_h = nullptr;
  m->dm.nchunk = 0; m->dm.ngc = 0;
  m->dm.jinvT = m->dm.detj = nullptr;
  m->dm.geom_nq = 0;
  if (desc->cell_jinvT) {
    // per-cell geometry, transposed to SoA [g][d][e][ldc] so that a warp (lane = cell) reads whole lines
    const int gq = desc->geom_nq, dd = desc->dim * desc->dim;
    std::vector<double> soa((size_t)gq * dd * ldc, 0.0), sdet((size_t)gq * ldc, 0.0);
    for (int64_t c = 0; c < nc; ++c)
      for (int g = 0; g < gq; ++g) {
        const double dj = desc->cell_detj[c * gq + g];
        if (!(dj > 0.0)) return bail(fail(PDENLP_EINVAL, "cell_detj must be positive (|det J| of the cell map)"));
        sdet[(size_t)g * ldc + c] = dj;
        for (int k = 0; k < dd; ++k) soa[((size_t)g * dd + k) * ldc + c] = desc->cell_jinvT[(c * gq + g) * dd + k];
      }
    if ((rc = dev_upload(m, soa.data(), soa.size(), &dp))) return bail(rc);
    m->dm.jinvT = dp;
    if ((rc = dev_upload(m, sdet.data(), sdet.size(), &dp))) return bail(rc);
    m->dm.detj = dp;
    m->dm.geom_nq = gq;
    // Jacobians that vary inside a cell make every derivative point-dependent: no single-point shortcuts
    if (gq > 1) m->dm.notrait = 1;
  }
  if (!ops->ktab().empty()) {
    if ((rc = dev_upload(m, ops->ktab().data(), ops->ktab().size(), &dp))) return bail(rc);
    m->dm.ktab = dp;
  }

  // per-tile entry counts -> exclusive scans (segment offsets)
  {
    const int64_t nt = m->dm.ntiles;
    int32_t *dcounts = nullptr, *dflag = nullptr;
    if (cudaMalloc((void**)&dcounts, nt * 3 * 4) != cudaSuccess || cudaMalloc((void**)&dflag, 4) != cudaSuccess)
      return bail(fail(PDENLP_ENOMEM, "cudaMalloc failed"));
    cudaMemset(dflag, 0, 4);
    rc = ops->count(m->dm, dcounts, dflag, nullptr);
    int32_t hflag = 0;
    int64_t* base3 = nullptr;  // [3][nt + 1]
    if (!rc && cudaMalloc((void**)&base3, (size_t)(nt + 1) * 3 * 8) != cudaSuccess) { cudaFree(dcounts); cudaFree(dflag); return bail(fail(PDENLP_ENOMEM, "cudaMalloc failed")); }
    if (base3) m->owned.push_back(base3);
    if (!rc) k_scan_counts<<<3, 1024>>>(dcounts, nt, base3, base3 + (nt + 1), base3 + 2 * (nt + 1));
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaMemcpy(&hflag, dflag, 4, cudaMemcpyDeviceToHost);
    // (host copies of the offsets: totals for the info block, per-tile values for the template classification)
    std::vector<int64_t> b0(nt + 1), b1(nt + 1), b2(nt + 1);
    if (!rc && e1 == cudaSuccess && e2 == cudaSuccess) {
      e1 = cudaMemcpy(b0.data(), base3, (size_t)(nt + 1) * 8, cudaMemcpyDeviceToHost);
      if (e1 == cudaSuccess) e1 = cudaMemcpy(b1.data(), base3 + (nt + 1), (size_t)(nt + 1) * 8, cudaMemcpyDeviceToHost);
      if (e1 == cudaSuccess) e1 = cudaMemcpy(b2.data(), base3 + 2 * (nt + 1), (size_t)(nt + 1) * 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(dcounts);
    cudaFree(dflag);
    if (rc) return bail(rc);
    if (e1 != cudaSuccess || e2 != cudaSuccess)
      return bail(fail(PDENLP_ECUDA, std::string("count kernel: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)));
    if (hflag) return bail(fail(PDENLP_EINVAL, "a cell repeats a dof id; not supported"));
    m->hb_jy = b0; m->hb_ju = b1; m->hb_h = b2;
    m->dm.base_jy = base3;
    m->dm.base_ju = base3 + (nt + 1);
    m->dm.base_h = base3 + 2 * (nt + 1);
    const int64_t p = desc->nparam, n = p + desc->nfree_y + desc->nfree_u;
    pdenlp_info& I = m->info;
    I.nvar = n;
    I.ncon = desc->nfree_y;
    I.nnzj_k = desc->nfree_y * p;
    I.nnzj_y = b0[nt];
    I.nnzj_u = b1[nt];
    I.nnzj = I.nnzj_k + I.nnzj_y + I.nnzj_u;
    I.nnzh_fe = b2[nt];
    I.nnzh_k_obj = p == 0 ? 0 : (m->inde ? p * (p + 1) / 2 : p * (p + 1) / 2 + (n - p) * p);
    I.nnzh_k_con = p == 0 ? 0 : p * (p + 1) / 2 + (n - p) * p;
    I.nnzh_fe_obj = m->nofe ? 0 : I.nnzh_fe;
    I.nnzh_obj = I.nnzh_k_obj + I.nnzh_fe_obj;
    I.nnzh = I.nnzh_obj + I.nnzh_k_con + I.nnzh_fe;
  }
  m->npartials = m->sms * 8 * 17;  // per-CTA partials: 1 objective + up to 16 kappa components, grids of <= 8 CTAs per SM
  if (cudaMalloc((void**)&m->partials, (size_t)m->npartials * 8) != cudaSuccess || cudaMalloc((void**)&m->scalar, 16) != cudaSuccess)
    return bail(fail(PDENLP_ENOMEM, "cudaMalloc failed"));
  m->owned.push_back(m->partials);
  m->owned.push_back(m->scalar);
  if (cudaMalloc((void**)&m->ticket, 16) != cudaSuccess || cudaMemset(m->ticket, 0, 16) != cudaSuccess)
    return bail(fail(PDENLP_ENOMEM, "cudaMalloc failed"));
  m->owned.push_back(m->ticket);
  if (cudaHostAlloc((void**)&m->scalar_host, 16, cudaHostAllocMapped) != cudaSuccess)
    return bail(fail(PDENLP_ENOMEM, "cudaHostAlloc failed"));
  m->dm.ticket = m->ticket;
  m->dm.fsum = m->scalar;
  m->dm.fsum_host = nullptr;
  { double* hd = nullptr; if (cudaHostGetDevicePointer((void**)&hd, m->scalar_host, 0) == cudaSuccess) m->dm.fsum_host = hd; }
  if (cudaEventCreateWithFlags(&m->ev_switch, cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(PDENLP_ECUDA, "cudaEventCreate failed"));
  if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess)
    return bail(fail(PDENLP_ECUDA, "cudaEventCreate failed"));
  if ((rc = classify_tiles(m, *desc))) return bail(rc);
  if (m->cg_active || (desc->nparam > 0 && !m->nofe)) {  // (models with kappa-block kernels, see Fork)
    // high priority: the CTAs of the latency-bound side kernels are placed ahead of the pending chunk-copy CTAs
    // (config 5: jac_coord! 0.475 -> 0.454 ms, hess_coord! 0.794 -> 0.759); PDENLP_AUX_PRIO=0|lo: experiments
    int plo = 0, phi = 0;
    cudaDeviceGetStreamPriorityRange(&plo, &phi);  // (lowest, highest)
    const char* pe = getenv("PDENLP_AUX_PRIO");
    const int prio = !pe ? phi : (std::string(pe) == "lo" ? plo : 0);
    if (cudaStreamCreateWithPriority(&m->aux, cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess)
      return bail(fail(PDENLP_ECUDA, "auxiliary stream / events: creation failed"));
  }
  if (desc->flags & PDENLP_FLAG_DETERMINISTIC) { if ((rc = colour_cells(m, *desc, offf, nfy, nfu))) return bail(rc); }
  if ((rc = self_check(m, &m->selfcheck_err))) return bail(rc);
  for (auto* v : {&m->hb_jy, &m->hb_ju, &m->hb_h, &m->ch_cell0_h, &m->ch_jy_h, &m->ch_ju_h, &m->ch_h_h}) std::vector<int64_t>().swap(*v);
  std::vector<int32_t>().swap(m->gc_cell_h);
  // the host arrays of the description are not referenced after create
  m->d.cell_dofs_y = m->d.cell_dofs_u = nullptr;
  m->d.dir_y = m->d.dir_u = m->d.phi_y = m->d.grad_y = m->d.phi_u = m->d.wdet = m->d.coef = nullptr;
  m->d.cell_jinvT = m->d.cell_detj = nullptr;
  m->d.field_nfree = m->d.field_ndir = nullptr;
  cudaSetDevice(prev);
  *out = m;
  return PDENLP_OK;
}

int pdenlp_destroy(pdenlp_model* m) {
  if (!m) return PDENLP_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(m->device);
  for (void* p : m->owned) cudaFree(p);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->scalar_host) cudaFreeHost(m->scalar_host);
  for (auto& g : m->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (m->cap) cudaStreamDestroy(m->cap);
  if (m->ev_switch) cudaEventDestroy(m->ev_switch);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->aux) { cudaStreamSynchronize(m->aux); cudaStreamDestroy(m->aux); }
  delete m->ops;
  delete m;
  cudaSetDevice(prev);
  return PDENLP_OK;
}

int pdenlp_get_info(const pdenlp_model* m, pdenlp_info* info) {
  if (!m || !info) return fail(PDENLP_EINVAL, "null argument");
  *info = m->info;
  return PDENLP_OK;
}
int pdenlp_set_stream(pdenlp_model* m, void* s) {
  if (!m) return fail(PDENLP_EINVAL, "null model");
  cudaStream_t ns = (cudaStream_t)s;
  if (ns != m->stream) {
    // Model-owned scratch (cached template images, partials, staging buffers) may still be in use by work enqueued on
    // the previous stream: the new stream waits for it (an event, no host synchronisation).  Skipped while either
    // stream is being captured: a capture starts from quiescent streams, and an event wait across a capture boundary
    // is not legal.
    cudaStreamCaptureStatus so = cudaStreamCaptureStatusNone, sn = cudaStreamCaptureStatusNone;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    if (cudaStreamIsCapturing(m->stream, &so) != cudaSuccess) { cudaGetLastError(); so = cudaStreamCaptureStatusNone; }
    if (cudaStreamIsCapturing(ns, &sn) != cudaSuccess) { cudaGetLastError(); sn = cudaStreamCaptureStatusNone; }
    if (so == cudaStreamCaptureStatusNone && sn == cudaStreamCaptureStatusNone && m->ev_switch &&
        cudaEventRecord(m->ev_switch, m->stream) == cudaSuccess)
      cudaStreamWaitEvent(ns, m->ev_switch, 0);
    if (prev != m->device && prev >= 0) cudaSetDevice(prev);
  }
  m->stream = ns;
  return PDENLP_OK;
}
int pdenlp_sync(pdenlp_model* m) {
  if (!m) return fail(PDENLP_EINVAL, "null model");
  CU(cudaStreamSynchronize(m->stream));
  return PDENLP_OK;
}
int pdenlp_malloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return fail(PDENLP_EINVAL, "bad argument");
  CU(cudaMalloc(ptr, (size_t)std::max<int64_t>(bytes, 16)));
  return PDENLP_OK;
}
int pdenlp_free(void* ptr) { CU(cudaFree(ptr)); return PDENLP_OK; }
int pdenlp_memcpy_h2d(void* dst, const void* src, int64_t bytes) { CU(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice)); return PDENLP_OK; }
int pdenlp_memcpy_d2h(void* dst, const void* src, int64_t bytes) { CU(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost)); return PDENLP_OK; }
int pdenlp_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return fail(PDENLP_EINVAL, "bad argument");
  CU(cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 16), cudaHostAllocDefault));
  return PDENLP_OK;
}
int pdenlp_host_free(void* ptr) { CU(cudaFreeHost(ptr)); return PDENLP_OK; }

int pdenlp_obj(pdenlp_model* m, const double* x, double* f) {
  if (!m || !x || !f) return fail(PDENLP_EINVAL, "null argument");
  if (!m->nofe && !m->dm.coef) return fail(PDENLP_EINVAL, "obj needs the coefficient fields (desc.coef)");
  Scope sc(m);
  const double* dx;
  if (int rc = in_ptr(m, x, m->info.nvar, &m->sx, &dx)) return rc;
  Launch L{m->stream, m->sms, false, &m->launches};
  if (m->nofe) {  // _obj_integral(::NoFETerm, kappa, x) = f(kappa)
    if (int rc = m->ops->nofe(L, NOFE_OBJ, dx, 1.0, nullptr, m->scalar)) return rc;
  } else {
    int nb = 0;
    // (the last CTA of the kernel sums the per-CTA partials in block order and writes the pinned host mirror)
    if (int rc = m->ops->vec(L, m->dm, OP_OBJ, dx, nullptr, nullptr, 1.0, nullptr, m->partials, &nb)) return rc;
    if (m->dm.fsum_host != nullptr) {
      CU(cudaStreamSynchronize(m->stream));
      *f = *(volatile double*)m->scalar_host;
      return PDENLP_OK;
    }
  }
  CU(cudaMemcpyAsync(f, m->scalar, 8, cudaMemcpyDeviceToHost, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  return PDENLP_OK;
}

static int vec_call(pdenlp_model* m, int op, const double* x, const double* lam, const double* v, int64_t nv, double ow,
                    double* out, int64_t nout, bool need_coef) {
  if (!m || !x || !out) return fail(PDENLP_EINVAL, "null argument");
  if (need_coef && !m->dm.coef) return fail(PDENLP_EINVAL, "this callback needs the coefficient fields (desc.coef)");
  Scope sc(m);
  const double *dx, *dl, *dv;
  double* dout;
  if (int rc = in_ptr(m, x, m->info.nvar, &m->sx, &dx)) return rc;
  if (int rc = in_ptr(m, lam, m->info.ncon, &m->slam, &dl, LAYOUT_C)) return rc;
  if (int rc = in_ptr(m, v, nv, &m->sv, &dv, op == OP_JTPROD ? LAYOUT_C : LAYOUT_X)) return rc;
  if (int rc = out_ptr(m, out, nout, &dout)) return rc;
  CU(cudaMemsetAsync(dout, 0, (size_t)nout * 8, m->stream));  // assemble_vector!/coo_prod! zero-fill first
  Launch L{m->stream, m->sms, false, &m->launches};
  int nb = 0;
  // NoFETerm objective: the FE part of grad! is zero (additional_obj_terms.jl:82), hprod! keeps its constraint part
  const bool skip_fe = m->nofe && (op == OP_GRAD || (op == OP_HPROD && dl == nullptr));
  if (!skip_fe) {
    // (deterministic mode: one launch of the generic kernel per colour, plain read-modify-write accumulation)
    if (int rc = for_each_colour(m, [&](const DevModel& d, bool coloured) {
          Launch Lc = L;
          Lc.generic = coloured;
          return m->ops->vec(Lc, d, op, dx, dl, dv, ow, dout, m->partials, &nb);
        })) return rc;
  }
  if (m->nofe && m->d.nparam > 0 && (op == OP_GRAD || op == OP_HPROD)) {
    if (int rc = m->ops->nofe(L, op == OP_GRAD ? NOFE_GRAD : NOFE_HPROD, dx, ow, dv, dout)) return rc;
  }
  return out_finish(m, out, nout);
}

// objgrad!(nlp, x, g) / objcons!(nlp, x, c): the objective and the gradient (resp. the residual) from ONE sweep
static int obj_and_vector(pdenlp_model* m, int op, const double* x, double* f, double* out, int64_t n,
                          int (*vector_only)(pdenlp_model*, const double*, double*)) {
  if (!m || !x || !f || !out) return fail(PDENLP_EINVAL, "null argument");
  if (m->nofe || m->deterministic) {  // NoFETerm: the objective comes from the one-thread kernel; deterministic: per colour
    if (int rc = pdenlp_obj(m, x, f)) return rc;
    return vector_only(m, x, out);
  }
  if (!m->dm.coef) return fail(PDENLP_EINVAL, "this callback needs the coefficient fields (desc.coef)");
  Scope sc(m);
  const double* dx;
  double* dout;
  if (int rc = in_ptr(m, x, m->info.nvar, &m->sx, &dx)) return rc;
  if (int rc = out_ptr(m, out, n, &dout)) return rc;
  CU(cudaMemsetAsync(dout, 0, (size_t)n * 8, m->stream));
  Launch L{m->stream, m->sms, false, &m->launches};
  int nb = 0;
  if (int rc = m->ops->vec(L, m->dm, op, dx, nullptr, nullptr, 1.0, dout, m->partials, &nb)) return rc;
  k_sum_partials<<<1, 256, 0, m->stream>>>(m->partials, nb, m->scalar);
  CU(cudaGetLastError());
  m->launches += 1;
  CU(cudaMemcpyAsync(f, m->scalar, 8, cudaMemcpyDeviceToHost, m->stream));
  if (m->d.memspace == PDENLP_MEM_HOST) return out_finish(m, out, n);
  CU(cudaStreamSynchronize(m->stream));
  return PDENLP_OK;
}
int pdenlp_objgrad(pdenlp_model* m, const double* x, double* f, double* g) {
  return obj_and_vector(m, OP_OBJGRAD, x, f, g, m ? m->info.nvar : 0, pdenlp_grad);
}
int pdenlp_objcons(pdenlp_model* m, const double* x, double* f, double* c) {
  return obj_and_vector(m, OP_OBJCONS, x, f, c, m ? m->info.ncon : 0, pdenlp_cons);
}

int pdenlp_grad(pdenlp_model* m, const double* x, double* g) {
  return vec_call(m, OP_GRAD, x, nullptr, nullptr, 0, 1.0, g, m ? m->info.nvar : 0, m && !m->nofe);
}
int pdenlp_cons(pdenlp_model* m, const double* x, double* c) {
  return vec_call(m, OP_CONS, x, nullptr, nullptr, 0, 1.0, c, m ? m->info.ncon : 0, true);
}
int pdenlp_jprod(pdenlp_model* m, const double* x, const double* v, double* Jv) {
  if (!v) return fail(PDENLP_EINVAL, "null argument");
  return vec_call(m, OP_JPROD, x, nullptr, v, m ? m->info.nvar : 0, 1.0, Jv, m ? m->info.ncon : 0, m && m->ops->needs_coef_deriv());
}
int pdenlp_jtprod(pdenlp_model* m, const double* x, const double* v, double* Jtv) {
  if (!v) return fail(PDENLP_EINVAL, "null argument");
  return vec_call(m, OP_JTPROD, x, nullptr, v, m ? m->info.ncon : 0, 1.0, Jtv, m ? m->info.nvar : 0, m && m->ops->needs_coef_deriv());
}
int pdenlp_hprod(pdenlp_model* m, const double* x, const double* lambda, double obj_weight, const double* v, double* Hv) {
  if (!v) return fail(PDENLP_EINVAL, "null argument");
  return vec_call(m, OP_HPROD, x, lambda, v, m ? m->info.nvar : 0, obj_weight, Hv, m ? m->info.nvar : 0, m && m->ops->needs_coef_deriv());
}

static int jac_coord_body(pdenlp_model* m, const double* x, double* vals) {
  Scope sc(m);
  const double* dx;
  double* dv;
  if (int rc = in_ptr(m, x, m->info.nvar, &m->sx, &dx)) return rc;
  if (int rc = out_ptr(m, vals, m->info.nnzj, &dv)) return rc;
  if (m->ops->needs_coef_deriv() && !m->dm.coef) return fail(PDENLP_EINVAL, "this integrand's derivatives need the coefficient fields (desc.coef)");
  Launch L{m->stream, m->sms, false, &m->launches};
  // kappa block (+ generic cells) beside the FE stream on the auxiliary stream; template models WITHOUT a kappa block
  // need no second stream: their chunk copy is chained behind the generic-cell kernel (Launch::chain)
  static const bool no_chain = getenv("PDENLP_NO_CHAIN") != nullptr;  // experiments / A-B runs
  Fork fk(m, m->info.nnzj_k > 0 || (m->cg_active && no_chain));
  if (m->info.nnzj_k > 0) {
    if (!m->dm.coef) return fail(PDENLP_EINVAL, "the kappa block needs the coefficient fields (desc.coef)");
    CU(cudaMemsetAsync(dv, 0, (size_t)m->info.nnzj_k * 8, fk.side()));
    if (int rc = for_each_colour(m, [&](const DevModel& d, bool coloured) {
          Launch Lc = L;
          Lc.stream = fk.side();
          Lc.bg_ctas = fk.bg_ctas();
          Lc.generic = coloured;
          return m->ops->jac_kappa(Lc, d, dx, dv, m->partials);
        })) return rc;
  }
  double* vy = dv + m->info.nnzj_k;
  double* vu = vy + m->info.nnzj_y;
  if (int rc = build_template(m, false, dx, 1.0)) return rc;
  L.aux = fk.aux();
  L.chain = !fk.on && !no_chain;
  if (int rc = m->ops->jac(L, m->dm, dx, vy, vu)) return rc;
  fk.join();
  return out_finish(m, vals, m->info.nnzj);
}

static int hess_coord_body(pdenlp_model* m, const double* x, const double* lambda, double obj_weight, double* vals) {
  Scope sc(m);
  const double *dx, *dl;
  double* dv;
  if (int rc = in_ptr(m, x, m->info.nvar, &m->sx, &dx)) return rc;
  if (int rc = in_ptr(m, lambda, m->info.ncon, &m->slam, &dl, LAYOUT_C)) return rc;
  if (int rc = out_ptr(m, vals, m->info.nnzh, &dv)) return rc;
  if (m->ops->needs_coef_deriv() && !m->dm.coef) return fail(PDENLP_EINVAL, "this integrand's derivatives need the coefficient fields (desc.coef)");
  Launch L{m->stream, m->sms, false, &m->launches};
  const pdenlp_info& I = m->info;
  double* vko = dv;
  double* vfo = vko + I.nnzh_k_obj;
  double* vkc = vfo + I.nnzh_fe_obj;
  double* vfc = vkc + I.nnzh_k_con;
  // kappa blocks (+ generic cells) beside the FE stream and the zero fill (see pdenlp_jac_coord)
  static const bool no_chain = getenv("PDENLP_NO_CHAIN") != nullptr;
  Fork fk(m, (I.nnzh_k_obj + I.nnzh_k_con > 0 && !m->nofe) || (m->cg_active && no_chain));
  if (I.nnzh_k_obj + I.nnzh_k_con > 0) {
    if (!m->dm.coef) return fail(PDENLP_EINVAL, "the kappa blocks need the coefficient fields (desc.coef)");
    CU(cudaMemsetAsync(vko, 0, (size_t)I.nnzh_k_obj * 8, fk.side()));
    CU(cudaMemsetAsync(vkc, 0, (size_t)I.nnzh_k_con * 8, fk.side()));
    Launch Lk = L;
    Lk.stream = fk.side();
    bool both = false;
    if (m->nofe) {
      if (int rc = m->ops->nofe(Lk, NOFE_HESS, dx, obj_weight, nullptr, vko)) return rc;
    } else if (dl) {  // objective and constraint trapezoids from one sweep (which = 2)
      both = true;
      if (int rc = for_each_colour(m, [&](const DevModel& d, bool coloured) {
            Launch Lc = Lk;
            Lc.generic = coloured;
            return m->ops->hess_kappa(Lc, d, dx, dl, obj_weight, 2, m->inde ? m->d.nparam : I.nvar, vko, m->partials, I.nvar, vkc);
          })) return rc;
    } else {
      if (int rc = for_each_colour(m, [&](const DevModel& d, bool coloured) {
            Launch Lc = Lk;
            Lc.generic = coloured;
            return m->ops->hess_kappa(Lc, d, dx, nullptr, obj_weight, 0, m->inde ? m->d.nparam : I.nvar, vko, m->partials);
          })) return rc;
    }
    if (dl && !both) {
      if (int rc = for_each_colour(m, [&](const DevModel& d, bool coloured) {
            Launch Lc = Lk;
            Lc.generic = coloured;
            return m->ops->hess_kappa(Lc, d, dx, dl, 1.0, 1, I.nvar, vkc, m->partials);
          })) return rc;
    }
  }
  // The constraint FE block is a pure zero fill when no multipliers are given (GridapPDENLPModel.jl:295-297)
  // or when the residual is affine in the FE fields (its second derivatives vanish identically): a plain
  // memset streams at the pure-write rate, without the read/write mix of the emitting kernel.
  static const bool no_zero_trait = getenv("PDENLP_NO_ZERO_TRAIT") != nullptr;  // experiments only
  const bool cons_zero = !dl || (m->ops->res_fe_hessian_zero() && !no_zero_trait);
  // (host memory space: that block is the tail of vals; it is zero-filled on the host, not transferred)
  const bool host_tail = cons_zero && m->d.memspace == PDENLP_MEM_HOST;
  // The fill is launched BEHIND the emitting kernel with programmatic stream serialization, so it starts while
  // that kernel is still running (disjoint outputs): hides one launch ramp/tail per callback (config 2: -7 us at
  // N = 1, ~5 % of hess_coord! on a 1/8 slab).  PDENLP_NO_PDL=1 restores the plain order fill -> kernel.
  static const bool pdl = getenv("PDENLP_NO_PDL") == nullptr;
  if (cons_zero && !host_tail && !pdl) {
    if (int rc = zero_fill(m, vfc, I.nnzh_fe)) return rc;
  }
  bool fill_done = false;  // the template path's copy launch also filled the zero block
  double* vo = m->nofe ? nullptr : vfo;  // NoFETerm: the objective has no FE block
  double* vc = cons_zero ? nullptr : vfc;
  if (vo && obj_weight == 0.0) {  // vals .*= obj_weight (GridapPDENLPModel.jl:299): the objective block is a zero fill too
    if (int rc = zero_fill(m, vfo, I.nnzh_fe_obj)) return rc;
    vo = nullptr;
  }
  if (vo || vc) {
    if (vo && !vc) { if (int rc = build_template(m, true, dx, obj_weight)) return rc; }
    L.aux = fk.aux();
    L.chain = !fk.on && !no_chain;
    if (cons_zero && !host_tail && pdl && vo && !vc) { L.fill_ptr = vfc; L.fill_n = I.nnzh_fe; L.fill_done = &fill_done; }
    if (int rc = m->ops->hess(L, m->dm, dx, dl, obj_weight, vo, vc)) return rc;
  }
  if (cons_zero && !host_tail && pdl && !fill_done) {
    if (int rc = zero_fill(m, vfc, I.nnzh_fe, (vo || vc) && !m->ops->is_multi())) return rc;
  }
  fk.join();
  if (host_tail) return out_finish_zero_tail(m, vals, I.nnzh - I.nnzh_fe, I.nnzh_fe);
  return out_finish(m, vals, I.nnzh);
}

// Replay cache.  jac_coord! / hess_coord! of a SMALL model with kappa blocks are chains of 4-9 stream operations (memsets,
// kappa kernels and their second passes on the auxiliary stream, the FE kernel, event fork / join) of 5-35 us each: the
// callback takes as long as the CPU needs to submit them (config 4 hess_coord!: 68 us for 45 us of kernels).  The second
// call with the same (callback, x, lambda, vals, obj_weight, stream) captures the chain into a CUDA graph; later calls
// replay it with one launch.  Pointers and obj_weight are baked into the nodes, the CONTENTS of x / lambda are read at
// replay time.  Only device-memspace models without templates (whose image cache is host state) and without timing
// brackets; never while the caller's stream is itself being captured.  PDENLP_NO_GRAPHS=1 turns it off.
}  // extern "C"  (graph_call is a template: no C linkage)

static bool graph_eligible(pdenlp_model* m) {
  static const bool off = getenv("PDENLP_NO_GRAPHS") != nullptr;
  if (off || m->d.memspace != PDENLP_MEM_DEVICE || m->timing || m->tpl_active || m->aux == nullptr || m->nofe) return false;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(m->stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st == cudaStreamCaptureStatusNone;
}
template <class Body>
static int graph_call(pdenlp_model* m, int cb, const void* x, const void* lam, const void* vals, double ow, Body&& body) {
  using GE = pdenlp_model::GraphEntry;
  GE* e = nullptr;
  for (auto& g : m->graphs)
    if (g.cb == cb && g.x == x && g.lam == lam && g.vals == vals && g.ow == ow && g.stream == m->stream) { e = &g; break; }
  m->graph_clock += 1;
  if (e == nullptr) {  // first sight of this argument set: run it plainly (one-time attribute / occupancy queries happen here)
    if (m->graphs.size() >= 8) {
      size_t old = 0;
      for (size_t k = 1; k < m->graphs.size(); ++k) if (m->graphs[k].used < m->graphs[old].used) old = k;
      if (m->graphs[old].exec) cudaGraphExecDestroy(m->graphs[old].exec);
      m->graphs.erase(m->graphs.begin() + old);
    }
    GE g;
    g.cb = cb; g.x = x; g.lam = lam; g.vals = vals; g.ow = ow; g.stream = m->stream; g.used = m->graph_clock;
    m->graphs.push_back(g);
    return body();
  }
  e->used = m->graph_clock;
  if (e->failed) return body();
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != m->device) cudaSetDevice(m->device);
  int rc = PDENLP_OK;
  if (e->exec == nullptr) {  // second call: capture
    const int64_t l0 = m->launches;
    // captured on a private stream (the caller's may be the legacy default stream, which cannot be captured); the
    // graph is then launched into the caller's stream
    cudaStream_t user = m->stream;
    if (m->cap == nullptr && cudaStreamCreateWithFlags(&m->cap, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); m->cap = nullptr; }
    if (m->cap == nullptr || cudaStreamBeginCapture(m->cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      e->failed = true;
      rc = body();
    } else {
      m->stream = m->cap;
      rc = body();
      m->stream = user;
      cudaGraph_t g = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(m->cap, &g);
      if (rc == PDENLP_OK && ce == cudaSuccess && g != nullptr && cudaGraphInstantiate(&e->exec, g, 0) == cudaSuccess) {
        e->launches = m->launches - l0;
      } else {  // not capturable here: remember, and do the work the plain way
        cudaGetLastError();
        e->exec = nullptr;
        e->failed = true;
        m->launches = l0;
        if (rc == PDENLP_OK) rc = body();
      }
      if (g) cudaGraphDestroy(g);
      if (e->exec) m->launches = l0;  // (counted below, per replay)
    }
  }
  if (rc == PDENLP_OK && e->exec != nullptr) {
    if (cudaGraphLaunch(e->exec, m->stream) != cudaSuccess) rc = fail(PDENLP_ECUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(cudaGetLastError()));
    else { m->launches += e->launches; m->graph_replays += 1; }
  }
  if (prev != m->device && prev >= 0) cudaSetDevice(prev);
  return rc;
}

extern "C" {

int pdenlp_jac_coord(pdenlp_model* m, const double* x, double* vals) {
  if (!m || !x || !vals) return fail(PDENLP_EINVAL, "null argument");
  if (m->info.nnzj_k > 0 && graph_eligible(m))
    return graph_call(m, 1, x, nullptr, vals, 0.0, [&] { return jac_coord_body(m, x, vals); });
  return jac_coord_body(m, x, vals);
}
int pdenlp_hess_coord(pdenlp_model* m, const double* x, const double* lambda, double obj_weight, double* vals) {
  if (!m || !x || !vals) return fail(PDENLP_EINVAL, "null argument");
  if (m->info.nnzh_k_obj + m->info.nnzh_k_con > 0 && graph_eligible(m))
    return graph_call(m, 2, x, lambda, vals, obj_weight, [&] { return hess_coord_body(m, x, lambda, obj_weight, vals); });
  return hess_coord_body(m, x, lambda, obj_weight, vals);
}

// objective part only, vals[nnzh_obj]: the whole Hessian of a model whose operator is affine
// (its constraint Hessian structure is empty, src/hessian_struct_nnzh_functions.jl:70-72)
int pdenlp_hess_coord_obj(pdenlp_model* m, const double* x, double obj_weight, double* vals) {
  if (!m || !x || !vals) return fail(PDENLP_EINVAL, "null argument");
  Scope sc(m);
  const double* dx;
  double* dv;
  const pdenlp_info& I = m->info;
  if (int rc = in_ptr(m, x, I.nvar, &m->sx, &dx)) return rc;
  if (int rc = out_ptr(m, vals, I.nnzh_obj, &dv)) return rc;
  if (m->ops->needs_coef_deriv() && !m->dm.coef) return fail(PDENLP_EINVAL, "this integrand's derivatives need the coefficient fields (desc.coef)");
  Launch L{m->stream, m->sms, false, &m->launches};
  static const bool no_chain = getenv("PDENLP_NO_CHAIN") != nullptr;
  Fork fk(m, (I.nnzh_k_obj > 0 && !m->nofe) || (m->cg_active && no_chain));
  if (I.nnzh_k_obj > 0) {
    if (!m->dm.coef) return fail(PDENLP_EINVAL, "the kappa blocks need the coefficient fields (desc.coef)");
    CU(cudaMemsetAsync(dv, 0, (size_t)I.nnzh_k_obj * 8, fk.side()));
    Launch Lk = L;
    Lk.stream = fk.side();
    if (m->nofe) {
      if (int rc = m->ops->nofe(Lk, NOFE_HESS, dx, obj_weight, nullptr, dv)) return rc;
    } else {
      if (int rc = for_each_colour(m, [&](const DevModel& d, bool coloured) {
            Launch Lc = Lk;
            Lc.generic = coloured;
            return m->ops->hess_kappa(Lc, d, dx, nullptr, obj_weight, 0, m->inde ? m->d.nparam : I.nvar, dv, m->partials);
          })) return rc;
    }
  }
  if (!m->nofe) {
    if (int rc = build_template(m, true, dx, obj_weight)) return rc;
    L.aux = fk.aux();
    L.chain = !fk.on && !no_chain;
    if (int rc = m->ops->hess(L, m->dm, dx, nullptr, obj_weight, dv + I.nnzh_k_obj, nullptr)) return rc;
  }
  fk.join();
  return out_finish(m, vals, I.nnzh_obj);
}

static int structure_call(pdenlp_model* m, bool hess, int64_t* rows, int64_t* cols) {
  if (!m || !rows || !cols) return fail(PDENLP_EINVAL, "null argument");
  Scope sc(m);
  const pdenlp_info& I = m->info;
  const int64_t nnz = hess ? I.nnzh : I.nnzj;
  int64_t *dr = rows, *dc = cols;
  const bool host = m->d.memspace == PDENLP_MEM_HOST;
  if (host) {
    CU(cudaMalloc((void**)&dr, std::max<int64_t>(nnz, 2) * 8));
    CU(cudaMalloc((void**)&dc, std::max<int64_t>(nnz, 2) * 8));
  }
  Launch L{m->stream, m->sms, false, &m->launches};
  const int64_t p = m->d.nparam, n = I.nvar;
  int rc = PDENLP_OK;
  // dense kappa blocks are tiny index patterns: build on the host and copy
  auto put = [&](int64_t off, const std::vector<int64_t>& r, const std::vector<int64_t>& c) -> int {
    if (r.empty()) return PDENLP_OK;
    CU(cudaMemcpyAsync(dr + off, r.data(), r.size() * 8, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(dc + off, c.data(), c.size() * 8, cudaMemcpyHostToDevice, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return PDENLP_OK;
  };
  auto trapezoid = [&](int64_t prows, std::vector<int64_t>& r, std::vector<int64_t>& c) {
    for (int64_t j = 1; j <= p; ++j)
      for (int64_t i = j; i <= prows; ++i) { r.push_back(i); c.push_back(j); }
  };
  if (!hess) {
    std::vector<int64_t> r, c;
    for (int64_t j = 1; j <= p; ++j)
      for (int64_t i = 1; i <= I.ncon; ++i) { r.push_back(i); c.push_back(j); }
    if (!rc) rc = put(0, r, c);
    if (!rc) rc = m->ops->structure(L, m->dm, 0, dr + I.nnzj_k, dc + I.nnzj_k, p);
    if (!rc && m->d.nu > 0) rc = m->ops->structure(L, m->dm, 1, dr + I.nnzj_k + I.nnzj_y, dc + I.nnzj_k + I.nnzj_y, p);
  } else {
    std::vector<int64_t> r, c;
    trapezoid(m->inde ? p : n, r, c);
    if (!rc) rc = put(0, r, c);
    r.clear(); c.clear();
    trapezoid(n, r, c);
    if (!rc) rc = put(I.nnzh_obj, r, c);
    if (!rc && !m->nofe) rc = m->ops->structure(L, m->dm, 2, dr + I.nnzh_k_obj, dc + I.nnzh_k_obj, p);
    if (!rc) rc = m->ops->structure(L, m->dm, 2, dr + I.nnzh_obj + I.nnzh_k_con, dc + I.nnzh_obj + I.nnzh_k_con, p);
  }
  if (host) {
    if (!rc) {
      cudaError_t e = cudaMemcpyAsync(rows, dr, nnz * 8, cudaMemcpyDeviceToHost, m->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(cols, dc, nnz * 8, cudaMemcpyDeviceToHost, m->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
      if (e != cudaSuccess) rc = fail(PDENLP_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(dr);
    cudaFree(dc);
  }
  return rc;
}
int pdenlp_jac_structure(pdenlp_model* m, int64_t* rows, int64_t* cols) { return structure_call(m, false, rows, cols); }
int pdenlp_hess_structure(pdenlp_model* m, int64_t* rows, int64_t* cols) { return structure_call(m, true, rows, cols); }

int pdenlp_selfcheck_error(const pdenlp_model* m, double* max_rel_err) {
  if (!m || !max_rel_err) return fail(PDENLP_EINVAL, "null argument");
  *max_rel_err = m->selfcheck_err;
  return PDENLP_OK;
}
int pdenlp_input_ranges(const pdenlp_model* m, int64_t ranges[4]) {
  if (!m || !ranges) return fail(PDENLP_EINVAL, "null argument");
  ranges[0] = m->yr[0]; ranges[1] = m->yr[1]; ranges[2] = m->ur[0]; ranges[3] = m->ur[1];
  return PDENLP_OK;
}
int pdenlp_launch_count(const pdenlp_model* m, int64_t* launches) {
  if (!m || !launches) return fail(PDENLP_EINVAL, "null argument");
  *launches = m->launches;
  return PDENLP_OK;
}
int pdenlp_template_stats(const pdenlp_model* m, int64_t stats[4]) {
  if (!m || !stats) return fail(PDENLP_EINVAL, "null argument");
  stats[0] = m->cg_active ? m->dm.nchunk : 0;
  stats[1] = m->cg_active ? m->ntpl_cells : 0;
  stats[2] = m->cg_active ? m->dm.ngc : 0;
  stats[3] = m->tpl_builds;
  return PDENLP_OK;
}
int pdenlp_graph_replays(const pdenlp_model* m, int64_t* replays) {
  if (!m || !replays) return fail(PDENLP_EINVAL, "null argument");
  *replays = m->graph_replays;
  return PDENLP_OK;
}
int pdenlp_set_timing(pdenlp_model* m, int32_t enabled) {
  if (!m) return fail(PDENLP_EINVAL, "null model");
  m->timing = enabled != 0;
  m->ev_valid = false;
  return PDENLP_OK;
}
int pdenlp_last_kernel_ms(pdenlp_model* m, float* ms) {
  if (!m || !ms) return fail(PDENLP_EINVAL, "null argument");
  if (!m->ev_valid) return fail(PDENLP_EINVAL, "no timed callback yet (pdenlp_set_timing)");
  CU(cudaEventSynchronize(m->ev1));
  CU(cudaEventElapsedTime(ms, m->ev0, m->ev1));
  return PDENLP_OK;
}


// =====================================================================================================================
// Several GPUs behind one handle (include/pdenlp_cuda.h, "pdenlp_sharded").  One host thread drives every device:
// stage the dof ranges the slab reads, launch the slab's callback on the device's own stream, copy the slab's slices
// to their place in the caller's host arrays; rows shared by two slabs and the kappa entries are summed on the host in
// device order.
// =====================================================================================================================
}  // extern "C"

struct pdenlp_sharded {
  int ndev = 0;
  std::vector<int> dev;
  std::vector<pdenlp_model*> m;
  std::vector<cudaStream_t> st;
  std::vector<int64_t> c0, c1;
  std::vector<int64_t> off_jy, off_ju, off_h, off_ho;  // [ndev+1] prefix sums of the local FE segment sizes
  std::vector<double*> dx, dlam, dv, dout, dcoo;       // device buffers of every device
  std::vector<double*> hstage;                         // page-locked staging (shared rows, kappa entries / blocks)
  std::vector<int64_t> hstage_len;
  pdenlp_info gi;
  int64_t p = 0, ny = 0, nu = 0;
  bool ordered = true;  // the slabs' dof ranges start in ascending order (contiguous slabs of a mesh numbered row by row)
};

namespace {

struct Piece { int64_t dst, len; };  // a run of staged values and where it is added in the caller's array

struct DevGuard {
  int prev = 0;
  explicit DevGuard(int d) { cudaGetDevice(&prev); cudaSetDevice(d); }
  ~DevGuard() { cudaSetDevice(prev); }
};

int sh_stage_buf(pdenlp_sharded* s, int k, int64_t n) {
  if (s->hstage_len[k] >= n) return PDENLP_OK;
  if (s->hstage[k]) cudaFreeHost(s->hstage[k]);
  s->hstage[k] = nullptr; s->hstage_len[k] = 0;
  CU(cudaHostAlloc((void**)&s->hstage[k], (size_t)std::max<int64_t>(n, 16) * 8, cudaHostAllocDefault));
  s->hstage_len[k] = n;
  return PDENLP_OK;
}
// host -> device copy of the parts of an input vector slab k reads (see in_ptr)
int sh_stage_in(pdenlp_sharded* s, int k, const double* user, double* dbuf, int layout) {
  const pdenlp_model* m = s->m[k];
  auto part = [&](int64_t lo, int64_t hi) -> int {
    if (hi > lo) CU(cudaMemcpyAsync(dbuf + lo, user + lo, (size_t)(hi - lo) * 8, cudaMemcpyHostToDevice, s->st[k]));
    return PDENLP_OK;
  };
  if (layout == LAYOUT_C) return part(m->yr[0], m->yr[1]);
  if (int rc = part(0, s->p)) return rc;
  if (int rc = part(s->p + m->yr[0], s->p + m->yr[1])) return rc;
  return part(s->p + s->ny + m->ur[0], s->p + s->ny + m->ur[1]);
}
int sh_sync_all(pdenlp_sharded* s) {
  for (int k = 0; k < s->ndev; ++k) {
    DevGuard g(s->dev[k]);
    CU(cudaStreamSynchronize(s->st[k]));
  }
  return PDENLP_OK;
}

// Assemble a dof-indexed output (x-layout: kappa | y | u, or constraint layout: y only) from the slabs' partial vectors
// dout[k].  Rows only one slab touches are copied straight to the caller's array; rows an earlier slab already covers
// are staged and added on the host after the copies (device order => deterministic); gaps are zero-filled.
int sh_gather_vector(pdenlp_sharded* s, double* out, bool xlayout) {
  const int nd = s->ndev;
  const int64_t p = xlayout ? s->p : 0;
  const int64_t n = xlayout ? s->gi.nvar : s->gi.ncon;
  std::vector<std::vector<Piece>> adds(nd);
  std::vector<int64_t> used(nd, 0);
  if (!s->ordered) memset(out, 0, (size_t)n * 8);
  for (int k = 0; k < nd; ++k) {
    const int64_t need = p + (s->m[k]->yr[1] - s->m[k]->yr[0]) + (xlayout ? s->m[k]->ur[1] - s->m[k]->ur[0] : 0);
    if (int rc = sh_stage_buf(s, k, need)) return rc;
  }
  auto staged = [&](int k, int64_t lo, int64_t hi) -> int {  // [lo, hi) of dout[k] -> staging, to be added at out[lo..hi)
    if (hi <= lo) return PDENLP_OK;
    DevGuard g(s->dev[k]);
    CU(cudaMemcpyAsync(s->hstage[k] + used[k], s->dout[k] + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s->st[k]));
    adds[k].push_back({lo, hi - lo});
    used[k] += hi - lo;
    return PDENLP_OK;
  };
  auto direct = [&](int k, int64_t lo, int64_t hi) -> int {
    if (hi <= lo) return PDENLP_OK;
    DevGuard g(s->dev[k]);
    CU(cudaMemcpyAsync(out + lo, s->dout[k] + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s->st[k]));
    return PDENLP_OK;
  };
  // kappa entries: every slab contributes
  if (p > 0) {
    memset(out, 0, (size_t)p * 8);
    for (int k = 0; k < nd; ++k)
      if (int rc = staged(k, 0, p)) return rc;
  }
  auto space = [&](int64_t base, int64_t nfree, bool is_u) -> int {
    int64_t cover = 0;  // rows [0, cover) of this space are written (or zero-filled) already
    for (int k = 0; k < nd; ++k) {
      const int64_t lo = is_u ? s->m[k]->ur[0] : s->m[k]->yr[0], hi = is_u ? s->m[k]->ur[1] : s->m[k]->yr[1];
      if (hi <= lo) continue;
      if (!s->ordered) { if (int rc = staged(k, base + lo, base + hi)) return rc; continue; }
      if (lo > cover) { memset(out + base + cover, 0, (size_t)(lo - cover) * 8); cover = lo; }
      if (int rc = staged(k, base + lo, base + std::min(hi, cover))) return rc;
      if (hi > cover) { if (int rc = direct(k, base + cover, base + hi)) return rc; cover = hi; }
    }
    if (s->ordered && cover < nfree) memset(out + base + cover, 0, (size_t)(nfree - cover) * 8);
    return PDENLP_OK;
  };
  if (int rc = space(p, s->ny, false)) return rc;
  if (xlayout && s->nu > 0) { if (int rc = space(p + s->ny, s->nu, true)) return rc; }
  if (int rc = sh_sync_all(s)) return rc;
  for (int k = 0; k < nd; ++k) {
    const double* src = s->hstage[k];
    for (const Piece& pc : adds[k]) {
      for (int64_t i = 0; i < pc.len; ++i) out[pc.dst + i] += src[i];
      src += pc.len;
    }
  }
  return PDENLP_OK;
}

// Sum a dense kappa block over the slabs: block = p columns; column j holds the x-layout rows [row0(j), nrows) at
// col_off(j) + (row - row0(j)).  Slab k contributes the rows it touches (kappa rows, its y rows, its u rows — given in
// block-row coordinates by `ranges`).  dblock[k] points at slab k's partial block on its device.
template <class ColOff, class Row0>
int sh_reduce_block(pdenlp_sharded* s, double* block, int64_t block_len, int ncols, ColOff col_off, Row0 row0,
                    const std::vector<std::vector<std::pair<int64_t, int64_t>>>& ranges, const std::vector<const double*>& dblock) {
  const int nd = s->ndev;
  std::vector<std::vector<Piece>> adds(nd);
  memset(block, 0, (size_t)block_len * 8);
  for (int k = 0; k < nd; ++k) {
    int64_t need = 0;
    for (int j = 0; j < ncols; ++j)
      for (auto& r : ranges[k]) need += std::max<int64_t>(0, r.second - std::max(r.first, row0(j)));
    if (int rc = sh_stage_buf(s, k, need)) return rc;
    DevGuard g(s->dev[k]);
    int64_t used = 0;
    for (int j = 0; j < ncols; ++j)
      for (auto& r : ranges[k]) {
        const int64_t a = std::max(r.first, row0(j)), b = r.second;
        if (b <= a) continue;
        const int64_t pos = col_off(j) + (a - row0(j));
        CU(cudaMemcpyAsync(s->hstage[k] + used, dblock[k] + pos, (size_t)(b - a) * 8, cudaMemcpyDeviceToHost, s->st[k]));
        adds[k].push_back({pos, b - a});
        used += b - a;
      }
  }
  if (int rc = sh_sync_all(s)) return rc;
  for (int k = 0; k < nd; ++k) {
    const double* src = s->hstage[k];
    for (const Piece& pc : adds[k]) {
      for (int64_t i = 0; i < pc.len; ++i) block[pc.dst + i] += src[i];
      src += pc.len;
    }
  }
  return PDENLP_OK;
}
// x-layout rows slab k touches: kappa rows, its state rows, its control rows
std::vector<std::vector<std::pair<int64_t, int64_t>>> sh_x_rows(const pdenlp_sharded* s) {
  std::vector<std::vector<std::pair<int64_t, int64_t>>> r(s->ndev);
  for (int k = 0; k < s->ndev; ++k) {
    r[k].push_back({0, s->p});
    r[k].push_back({s->p + s->m[k]->yr[0], s->p + s->m[k]->yr[1]});
    if (s->nu > 0) r[k].push_back({s->p + s->ny + s->m[k]->ur[0], s->p + s->ny + s->m[k]->ur[1]});
  }
  return r;
}

int sh_vector_call(pdenlp_sharded* s, int op, const double* x, const double* lam, const double* v, double ow, double* out) {
  if (!s || !x || !out) return fail(PDENLP_EINVAL, "null argument");
  const bool out_x = op == OP_GRAD || op == OP_JTPROD || op == OP_HPROD;
  for (int k = 0; k < s->ndev; ++k) {
    DevGuard g(s->dev[k]);
    if (int rc = sh_stage_in(s, k, x, s->dx[k], LAYOUT_X)) return rc;
    if (lam) { if (int rc = sh_stage_in(s, k, lam, s->dlam[k], LAYOUT_C)) return rc; }
    if (v) { if (int rc = sh_stage_in(s, k, v, s->dv[k], op == OP_JTPROD ? LAYOUT_C : LAYOUT_X)) return rc; }
    int rc = PDENLP_OK;
    switch (op) {
      case OP_GRAD: rc = pdenlp_grad(s->m[k], s->dx[k], s->dout[k]); break;
      case OP_CONS: rc = pdenlp_cons(s->m[k], s->dx[k], s->dout[k]); break;
      case OP_JPROD: rc = pdenlp_jprod(s->m[k], s->dx[k], s->dv[k], s->dout[k]); break;
      case OP_JTPROD: rc = pdenlp_jtprod(s->m[k], s->dx[k], s->dv[k], s->dout[k]); break;
      // (a NoFETerm objective fk(kappa) is not a sum over cells: only the first slab contributes it)
      case OP_HPROD: rc = pdenlp_hprod(s->m[k], s->dx[k], lam ? s->dlam[k] : nullptr, (s->m[k]->nofe && k > 0) ? 0.0 : ow, s->dv[k], s->dout[k]); break;
      default: rc = fail(PDENLP_EINVAL, "bad vector op");
    }
    if (rc) return rc;
    if (op == OP_GRAD && s->m[k]->nofe && k > 0 && s->p > 0) CU(cudaMemsetAsync(s->dout[k], 0, (size_t)s->p * 8, s->st[k]));
  }
  return sh_gather_vector(s, out, out_x);
}

}  // namespace

extern "C" {

int pdenlp_sharded_destroy(pdenlp_sharded* s) {
  if (!s) return PDENLP_OK;
  for (int k = 0; k < (int)s->m.size(); ++k) {
    DevGuard g(s->dev[k]);
    for (double* b : {s->dx[k], s->dlam[k], s->dv[k], s->dout[k], s->dcoo[k]})
      if (b) cudaFree(b);
    if (s->hstage[k]) cudaFreeHost(s->hstage[k]);
    if (s->m[k]) pdenlp_destroy(s->m[k]);
    if (s->st[k]) cudaStreamDestroy(s->st[k]);
  }
  delete s;
  return PDENLP_OK;
}

int pdenlp_sharded_create(const pdenlp_desc* desc, int32_t ndev, const int32_t* devices, pdenlp_sharded** out) {
  if (!desc || !out) return fail(PDENLP_EINVAL, "null argument");
  *out = nullptr;
  int have = 0;
  CU(cudaGetDeviceCount(&have));
  if (have == 0) return fail(PDENLP_ECUDA, "no CUDA device (libpdenlp_cuda has no CPU fallback)");
  if (ndev < 1 || ndev > have) return fail(PDENLP_EINVAL, "ndev must be between 1 and the number of CUDA devices");
  const int64_t nc = desc->ncells, ntiles = (nc + 31) / 32;
  if (ntiles < ndev) return fail(PDENLP_EINVAL, "fewer 32-cell tiles than devices");
  auto* s = new pdenlp_sharded();
  s->ndev = ndev;
  s->dev.resize(ndev); s->m.assign(ndev, nullptr); s->st.assign(ndev, nullptr);
  s->c0.resize(ndev); s->c1.resize(ndev);
  s->dx.assign(ndev, nullptr); s->dlam.assign(ndev, nullptr); s->dv.assign(ndev, nullptr); s->dout.assign(ndev, nullptr); s->dcoo.assign(ndev, nullptr);
  s->hstage.assign(ndev, nullptr); s->hstage_len.assign(ndev, 0);
  s->off_jy.assign(ndev + 1, 0); s->off_ju.assign(ndev + 1, 0); s->off_h.assign(ndev + 1, 0); s->off_ho.assign(ndev + 1, 0);
  s->p = desc->nparam; s->ny = desc->nfree_y; s->nu = desc->nfree_u;
  const int nfy = desc_nfy(*desc), nfu = desc_nfu(*desc);
  auto bail = [&](int code) { pdenlp_sharded_destroy(s); return code; };
  for (int k = 0; k < ndev; ++k) {
    s->dev[k] = devices ? devices[k] : k;
    if (s->dev[k] < 0 || s->dev[k] >= have) return bail(fail(PDENLP_EINVAL, "bad device ordinal"));
    const int64_t t0 = ntiles * k / ndev, t1 = ntiles * (k + 1) / ndev;
    s->c0[k] = std::min(t0 * 32, nc); s->c1[k] = std::min(t1 * 32, nc);
    const int64_t a = s->c0[k], ncl = s->c1[k] - a;
    pdenlp_desc sub = *desc;
    sub.device = s->dev[k];
    sub.memspace = PDENLP_MEM_DEVICE;
    sub.ncells = ncl;
    sub.cell_dofs_y = desc->cell_dofs_y + a * nfy * desc->ny;
    sub.cell_dofs_u = desc->cell_dofs_u ? desc->cell_dofs_u + a * nfu * desc->nu : nullptr;
    std::vector<double> cc;
    if (desc->coef && desc->ncoef > 0) {  // [field][cell][q]: the slab's rows of every field
      cc.resize((size_t)desc->ncoef * ncl * desc->nq);
      for (int f = 0; f < desc->ncoef; ++f)
        memcpy(cc.data() + (size_t)f * ncl * desc->nq, desc->coef + ((size_t)f * nc + a) * desc->nq, (size_t)ncl * desc->nq * 8);
      sub.coef = cc.data();
    }
    if (desc->cell_jinvT) {
      const int64_t gq = desc->geom_nq, dd = (int64_t)desc->dim * desc->dim;
      sub.cell_jinvT = desc->cell_jinvT + a * gq * dd;
      sub.cell_detj = desc->cell_detj + a * gq;
    }
    if (int rc = pdenlp_create(&sub, &s->m[k])) return bail(rc);
    DevGuard g(s->dev[k]);
    if (cudaStreamCreateWithFlags(&s->st[k], cudaStreamNonBlocking) != cudaSuccess) return bail(fail(PDENLP_ECUDA, "cudaStreamCreate failed"));
    pdenlp_set_stream(s->m[k], s->st[k]);
    const pdenlp_info& I = s->m[k]->info;
    s->off_jy[k + 1] = s->off_jy[k] + I.nnzj_y;
    s->off_ju[k + 1] = s->off_ju[k] + I.nnzj_u;
    s->off_h[k + 1] = s->off_h[k] + I.nnzh_fe;
    s->off_ho[k + 1] = s->off_ho[k] + I.nnzh_fe_obj;
    auto alloc = [&](double** b, int64_t n) { return cudaMalloc((void**)b, (size_t)std::max<int64_t>(n, 2) * 8) == cudaSuccess; };
    if (!alloc(&s->dx[k], I.nvar) || !alloc(&s->dlam[k], I.ncon) || !alloc(&s->dv[k], I.nvar) || !alloc(&s->dout[k], I.nvar) ||
        !alloc(&s->dcoo[k], std::max(I.nnzj, I.nnzh)))
      return bail(fail(PDENLP_ENOMEM, "cudaMalloc failed"));
    // stale entries outside the staged ranges are never read by the slab; keep them finite anyway
    cudaMemsetAsync(s->dx[k], 0, (size_t)I.nvar * 8, s->st[k]);
    cudaMemsetAsync(s->dlam[k], 0, (size_t)std::max<int64_t>(I.ncon, 2) * 8, s->st[k]);
    cudaMemsetAsync(s->dv[k], 0, (size_t)I.nvar * 8, s->st[k]);
    if (k > 0 && (s->m[k]->yr[0] < s->m[k - 1]->yr[0] || s->m[k]->ur[0] < s->m[k - 1]->ur[0])) s->ordered = false;
  }
  const pdenlp_info& I0 = s->m[0]->info;
  pdenlp_info& G = s->gi;
  G = I0;
  G.nnzj_y = s->off_jy[ndev]; G.nnzj_u = s->off_ju[ndev];
  G.nnzj = G.nnzj_k + G.nnzj_y + G.nnzj_u;
  G.nnzh_fe = s->off_h[ndev]; G.nnzh_fe_obj = s->off_ho[ndev];
  G.nnzh_obj = G.nnzh_k_obj + G.nnzh_fe_obj;
  G.nnzh = G.nnzh_obj + G.nnzh_k_con + G.nnzh_fe;
  if (int rc = sh_sync_all(s)) return bail(rc);
  *out = s;
  return PDENLP_OK;
}

int pdenlp_sharded_get_info(const pdenlp_sharded* s, pdenlp_info* info) {
  if (!s || !info) return fail(PDENLP_EINVAL, "null argument");
  *info = s->gi;
  return PDENLP_OK;
}
int pdenlp_sharded_slab(const pdenlp_sharded* s, int32_t k, int64_t* c0, int64_t* c1) {
  if (!s || !c0 || !c1 || k < 0 || k >= s->ndev) return fail(PDENLP_EINVAL, "bad argument");
  *c0 = s->c0[k]; *c1 = s->c1[k];
  return PDENLP_OK;
}

int pdenlp_sharded_obj(pdenlp_sharded* s, const double* x, double* f) {
  if (!s || !x || !f) return fail(PDENLP_EINVAL, "null argument");
  for (int k = 0; k < s->ndev; ++k) {
    DevGuard g(s->dev[k]);
    if (int rc = sh_stage_in(s, k, x, s->dx[k], LAYOUT_X)) return rc;
  }
  double sum = 0.0;
  for (int k = 0; k < s->ndev; ++k) {
    if (s->m[k]->nofe && k > 0) continue;  // a NoFETerm objective fk(kappa) is not a sum over cells
    double fk = 0.0;
    if (int rc = pdenlp_obj(s->m[k], s->dx[k], &fk)) return rc;
    sum += fk;
  }
  *f = sum;
  return PDENLP_OK;
}
int pdenlp_sharded_grad(pdenlp_sharded* s, const double* x, double* g) {
  return sh_vector_call(s, OP_GRAD, x, nullptr, nullptr, 1.0, g);
}
int pdenlp_sharded_cons(pdenlp_sharded* s, const double* x, double* c) { return sh_vector_call(s, OP_CONS, x, nullptr, nullptr, 1.0, c); }
int pdenlp_sharded_jprod(pdenlp_sharded* s, const double* x, const double* v, double* Jv) {
  if (!v) return fail(PDENLP_EINVAL, "null argument");
  return sh_vector_call(s, OP_JPROD, x, nullptr, v, 1.0, Jv);
}
int pdenlp_sharded_jtprod(pdenlp_sharded* s, const double* x, const double* v, double* Jtv) {
  if (!v) return fail(PDENLP_EINVAL, "null argument");
  return sh_vector_call(s, OP_JTPROD, x, nullptr, v, 1.0, Jtv);
}
int pdenlp_sharded_hprod(pdenlp_sharded* s, const double* x, const double* lambda, double obj_weight, const double* v, double* Hv) {
  if (!v) return fail(PDENLP_EINVAL, "null argument");
  return sh_vector_call(s, OP_HPROD, x, lambda, v, obj_weight, Hv);
}

int pdenlp_sharded_jac_coord(pdenlp_sharded* s, const double* x, double* vals) {
  if (!s || !x || !vals) return fail(PDENLP_EINVAL, "null argument");
  const pdenlp_info& G = s->gi;
  for (int k = 0; k < s->ndev; ++k) {
    DevGuard g(s->dev[k]);
    const pdenlp_info& I = s->m[k]->info;
    if (int rc = sh_stage_in(s, k, x, s->dx[k], LAYOUT_X)) return rc;
    if (int rc = pdenlp_jac_coord(s->m[k], s->dx[k], s->dcoo[k])) return rc;
    CU(cudaMemcpyAsync(vals + G.nnzj_k + s->off_jy[k], s->dcoo[k] + I.nnzj_k, (size_t)I.nnzj_y * 8, cudaMemcpyDeviceToHost, s->st[k]));
    if (I.nnzj_u > 0)
      CU(cudaMemcpyAsync(vals + G.nnzj_k + G.nnzj_y + s->off_ju[k], s->dcoo[k] + I.nnzj_k + I.nnzj_y, (size_t)I.nnzj_u * 8, cudaMemcpyDeviceToHost, s->st[k]));
  }
  if (G.nnzj_k > 0) {  // dense ncon x p block, column-major: slab k contributes its state rows
    std::vector<std::vector<std::pair<int64_t, int64_t>>> rows(s->ndev);
    std::vector<const double*> blk(s->ndev);
    for (int k = 0; k < s->ndev; ++k) { rows[k].push_back({s->m[k]->yr[0], s->m[k]->yr[1]}); blk[k] = s->dcoo[k]; }
    const int64_t ncon = G.ncon;
    return sh_reduce_block(s, vals, G.nnzj_k, (int)s->p, [&](int j) { return (int64_t)j * ncon; }, [](int) { return (int64_t)0; }, rows, blk);
  }
  return sh_sync_all(s);
}

int pdenlp_sharded_hess_coord(pdenlp_sharded* s, const double* x, const double* lambda, double obj_weight, double* vals) {
  if (!s || !x || !vals) return fail(PDENLP_EINVAL, "null argument");
  const pdenlp_info& G = s->gi;
  double* const gko = vals;
  double* const gfo = gko + G.nnzh_k_obj;
  double* const gkc = gfo + G.nnzh_fe_obj;
  double* const gfc = gkc + G.nnzh_k_con;
  const bool cons_zero = !lambda || s->m[0]->ops->res_fe_hessian_zero();
  for (int k = 0; k < s->ndev; ++k) {
    DevGuard g(s->dev[k]);
    const pdenlp_info& I = s->m[k]->info;
    if (int rc = sh_stage_in(s, k, x, s->dx[k], LAYOUT_X)) return rc;
    if (lambda) { if (int rc = sh_stage_in(s, k, lambda, s->dlam[k], LAYOUT_C)) return rc; }
    if (int rc = pdenlp_hess_coord(s->m[k], s->dx[k], lambda ? s->dlam[k] : nullptr, obj_weight, s->dcoo[k])) return rc;
    if (I.nnzh_fe_obj > 0)
      CU(cudaMemcpyAsync(gfo + s->off_ho[k], s->dcoo[k] + I.nnzh_k_obj, (size_t)I.nnzh_fe_obj * 8, cudaMemcpyDeviceToHost, s->st[k]));
    if (!cons_zero)
      CU(cudaMemcpyAsync(gfc + s->off_h[k], s->dcoo[k] + I.nnzh_obj + I.nnzh_k_con, (size_t)I.nnzh_fe * 8, cudaMemcpyDeviceToHost, s->st[k]));
  }
  if (cons_zero) host_zero_parallel(gfc, G.nnzh_fe);  // structurally zero block: filled on the host while the copies run
  if (s->p > 0) {
    const int64_t n = G.nvar, p = s->p;
    const auto xrows = sh_x_rows(s);
    std::vector<std::vector<std::pair<int64_t, int64_t>>> krows(s->ndev);
    for (int k = 0; k < s->ndev; ++k) krows[k].push_back({0, p});
    std::vector<const double*> blk(s->ndev);
    // objective block: p x p corner (inde energies, NoFETerm) or the full trapezoid
    const bool corner = s->m[0]->inde;
    const int64_t prows_o = corner ? p : n;
    for (int k = 0; k < s->ndev; ++k) blk[k] = s->dcoo[k];
    if (s->m[0]->nofe) {  // fk(kappa) is not a sum over cells: the block of one slab is the block
      DevGuard g(s->dev[0]);
      CU(cudaMemcpyAsync(gko, s->dcoo[0], (size_t)G.nnzh_k_obj * 8, cudaMemcpyDeviceToHost, s->st[0]));
    } else if (int rc = sh_reduce_block(s, gko, G.nnzh_k_obj, (int)p, [&](int j) { return (int64_t)j * prows_o - (int64_t)j * (j - 1) / 2; },
                                        [](int j) { return (int64_t)j; }, corner ? krows : xrows, blk)) return rc;
    for (int k = 0; k < s->ndev; ++k) blk[k] = s->dcoo[k] + s->m[k]->info.nnzh_obj;
    if (lambda) {
      if (int rc = sh_reduce_block(s, gkc, G.nnzh_k_con, (int)p, [&](int j) { return (int64_t)j * n - (int64_t)j * (j - 1) / 2; },
                                   [](int j) { return (int64_t)j; }, xrows, blk)) return rc;
    } else {
      memset(gkc, 0, (size_t)G.nnzh_k_con * 8);
    }
  }
  return sh_sync_all(s);
}

static int sh_structure(pdenlp_sharded* s, bool hess, int64_t* rows, int64_t* cols) {
  if (!s || !rows || !cols) return fail(PDENLP_EINVAL, "null argument");
  const pdenlp_info& G = s->gi;
  for (int k = 0; k < s->ndev; ++k) {
    DevGuard g(s->dev[k]);
    const pdenlp_info& I = s->m[k]->info;
    const int64_t nl = hess ? I.nnzh : I.nnzj;
    int64_t *dr = nullptr, *dc = nullptr;
    CU(cudaMalloc((void**)&dr, (size_t)std::max<int64_t>(nl, 2) * 8));
    if (cudaMalloc((void**)&dc, (size_t)std::max<int64_t>(nl, 2) * 8) != cudaSuccess) { cudaFree(dr); return fail(PDENLP_ENOMEM, "cudaMalloc failed"); }
    int rc = hess ? pdenlp_hess_structure(s->m[k], dr, dc) : pdenlp_jac_structure(s->m[k], dr, dc);
    auto get = [&](int64_t dst, int64_t src, int64_t n) {
      if (rc || n <= 0) return;
      cudaError_t e = cudaMemcpyAsync(rows + dst, dr + src, (size_t)n * 8, cudaMemcpyDeviceToHost, s->st[k]);
      if (e == cudaSuccess) e = cudaMemcpyAsync(cols + dst, dc + src, (size_t)n * 8, cudaMemcpyDeviceToHost, s->st[k]);
      if (e != cudaSuccess) rc = fail(PDENLP_ECUDA, cudaGetErrorString(e));
    };
    if (!hess) {
      if (k == 0) get(0, 0, G.nnzj_k);  // the dense kappa block has the same pattern on every slab
      get(G.nnzj_k + s->off_jy[k], I.nnzj_k, I.nnzj_y);
      get(G.nnzj_k + G.nnzj_y + s->off_ju[k], I.nnzj_k + I.nnzj_y, I.nnzj_u);
    } else {
      if (k == 0) { get(0, 0, G.nnzh_k_obj); get(G.nnzh_obj, I.nnzh_obj, G.nnzh_k_con); }
      get(G.nnzh_k_obj + s->off_ho[k], I.nnzh_k_obj, I.nnzh_fe_obj);
      get(G.nnzh_obj + G.nnzh_k_con + s->off_h[k], I.nnzh_obj + I.nnzh_k_con, I.nnzh_fe);
    }
    cudaError_t e = cudaStreamSynchronize(s->st[k]);
    cudaFree(dr); cudaFree(dc);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(PDENLP_ECUDA, cudaGetErrorString(e));
  }
  return PDENLP_OK;
}
int pdenlp_sharded_jac_structure(pdenlp_sharded* s, int64_t* rows, int64_t* cols) { return sh_structure(s, false, rows, cols); }
int pdenlp_sharded_hess_structure(pdenlp_sharded* s, int64_t* rows, int64_t* cols) { return sh_structure(s, true, rows, cols); }

}  // extern "C"
