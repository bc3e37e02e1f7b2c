// compaction.cu — COO -> compressed sparse (CSC / CSR) with duplicates summed, on the device.
//
// What consumes the COO streams of jac_coord!/hess_coord! next:
//   * NLPModels `jac(nlp,x)` / `hess(nlp,x,y)` = sparse(rows, cols, vals): duplicates are summed
//     (the obj and cons Hessian segments overlap; cells share dofs);
//   * the linear API of an AffineFEOperator model: `jac_lin_structure!`/`jac_lin_coord!` return
//     findnz(get_matrix(op)) — the assembled matrix in CSC order, rows ascending inside a column
//     (/root/reference/src/jacobian_functions.jl:11-21, src/GridapPDENLPModel.jl:597-608).
// The sparsity pattern is static, so the plan (a stable sort of the COO entries by destination and
// the segment offsets) is built once; applying it is a deterministic gather-sum with no atomics:
// each output entry adds its contributions in COO (= cell) order, like the reference assembly.
#include "../../include/pdenlp_cuda.h"

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <string>

extern thread_local std::string pdenlp_g_err;  // defined in pdenlp_cuda.cu
namespace {
int cfail(int code, const std::string& msg) {
  pdenlp_g_err = msg;
  return code;
}
#define CCU(call)                                                                                \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) return cfail(PDENLP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

__global__ void k_make_keys(int64_t nnz, const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                            int64_t major_dim_unused, int64_t minor_dim, int csc, uint64_t* __restrict__ keys,
                            int64_t* __restrict__ idx) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int64_t r = rows[k] - 1, c = cols[k] - 1;
  // CSC: major = column, minor = row; CSR: major = row, minor = column
  keys[k] = csc ? (uint64_t)c * (uint64_t)minor_dim + (uint64_t)r : (uint64_t)r * (uint64_t)minor_dim + (uint64_t)c;
  idx[k] = k;
}
__global__ void k_flag_heads(int64_t nnz, const uint64_t* __restrict__ keys, int64_t* __restrict__ flags) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}
// offsets[e] = first sorted position of output entry e; flags_scan is the inclusive scan of the heads
__global__ void k_offsets(int64_t nnz, const int64_t* __restrict__ incl, const uint64_t* __restrict__ keys,
                          int64_t minor_dim, int csc, int64_t* __restrict__ offsets, int64_t* __restrict__ orow,
                          int64_t* __restrict__ ocol) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const bool head = (k == 0) || incl[k] != incl[k - 1];
  if (head) {
    const int64_t e = incl[k] - 1;
    offsets[e] = k;
    const int64_t major = (int64_t)(keys[k] / (uint64_t)minor_dim), minor = (int64_t)(keys[k] % (uint64_t)minor_dim);
    orow[e] = (csc ? minor : major) + 1;
    ocol[e] = (csc ? major : minor) + 1;
  }
}
// one thread per output entry: segments are short (<= a handful of cells share a dof pair)
__global__ void k_apply(int64_t nout, const int64_t* __restrict__ offsets, const int64_t* __restrict__ perm,
                        const double* __restrict__ vals, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nout) return;
  double a = 0.0;
  for (int64_t k = offsets[e]; k < offsets[e + 1]; ++k) a += __ldg(vals + perm[k]);
  out[e] = a;
}
// pointer array (colptr for CSC, rowptr for CSR), 1-based like Julia's SparseMatrixCSC.colptr
__global__ void k_major_ptr(int64_t nout, const int64_t* __restrict__ major_ids /*1-based*/, int64_t nmajor,
                            int64_t* __restrict__ ptr) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > nout) return;
  // ptr[m] = 1 + number of entries with major id <= m  (m = 0..nmajor), filled by the boundaries
  const int64_t cur = e < nout ? major_ids[e] : nmajor + 1;
  const int64_t prev = e > 0 ? major_ids[e - 1] : 0;
  for (int64_t m = prev; m < cur; ++m) ptr[m] = e + 1;
}
inline int grid_of(int64_t n) { return (int)((n + 255) / 256); }
}  // namespace

struct pdenlp_compaction {
  int device = 0;
  int csc = 1;
  int64_t nnz_in = 0, nnz_out = 0, nrows = 0, ncols = 0;
  int64_t* perm = nullptr;     // [nnz_in] COO index of the k-th sorted entry
  int64_t* offsets = nullptr;  // [nnz_out + 1]
  int64_t* orow = nullptr;     // [nnz_out] 1-based
  int64_t* ocol = nullptr;
};

extern "C" {

int pdenlp_compaction_create(int64_t nnz, const int64_t* rows, const int64_t* cols, int64_t nrows, int64_t ncols,
                             int32_t order, int32_t rows_cols_on_device, pdenlp_compaction** out) {
  if (!out || !rows || !cols || nnz < 0 || nrows <= 0 || ncols <= 0) return cfail(PDENLP_EINVAL, "bad argument");
  if (order != PDENLP_ORDER_CSC && order != PDENLP_ORDER_CSR) return cfail(PDENLP_EINVAL, "order must be CSC or CSR");
  *out = nullptr;
  int dev = 0;
  CCU(cudaGetDevice(&dev));
  auto* P = new pdenlp_compaction();
  P->device = dev;
  P->csc = order == PDENLP_ORDER_CSC;
  P->nnz_in = nnz;
  P->nrows = nrows;
  P->ncols = ncols;
  const int64_t minor = P->csc ? nrows : ncols;
  const int64_t n1 = nnz > 0 ? nnz : 1;
  int64_t *drows = nullptr, *dcols = nullptr, *idx_in = nullptr, *flags = nullptr, *incl = nullptr;
  uint64_t *keys_in = nullptr, *keys_out = nullptr;
  void* tmp = nullptr;
  int rc = PDENLP_OK;
  auto cleanup = [&]() {
    if (!rows_cols_on_device) { cudaFree(drows); cudaFree(dcols); }
    cudaFree(idx_in); cudaFree(flags); cudaFree(incl); cudaFree(keys_in); cudaFree(keys_out); cudaFree(tmp);
  };
#define TRY(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      rc = cfail(PDENLP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));  \
      cleanup();                                                                      \
      pdenlp_compaction_destroy(P);                                                   \
      return rc;                                                                      \
    }                                                                                 \
  } while (0)
  if (rows_cols_on_device) {
    // the caller may have produced rows/cols on a non-blocking stream (pdenlp_*_structure on the model's stream), which
    // is not ordered with the legacy default stream used below: wait for the device once (setup-time only)
    TRY(cudaDeviceSynchronize());
    drows = const_cast<int64_t*>(rows);
    dcols = const_cast<int64_t*>(cols);
  } else {
    TRY(cudaMalloc((void**)&drows, n1 * 8));
    TRY(cudaMalloc((void**)&dcols, n1 * 8));
    TRY(cudaMemcpy(drows, rows, nnz * 8, cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(dcols, cols, nnz * 8, cudaMemcpyHostToDevice));
  }
  TRY(cudaMalloc((void**)&keys_in, n1 * 8));
  TRY(cudaMalloc((void**)&keys_out, n1 * 8));
  TRY(cudaMalloc((void**)&idx_in, n1 * 8));
  TRY(cudaMalloc((void**)&P->perm, n1 * 8));
  if (nnz > 0) {
    k_make_keys<<<grid_of(nnz), 256>>>(nnz, drows, dcols, 0, minor, P->csc, keys_in, idx_in);
    TRY(cudaGetLastError());
    size_t tb = 0;
    TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, idx_in, P->perm, (int64_t)nnz));  // stable
    TRY(cudaMalloc(&tmp, tb));
    TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys_out, idx_in, P->perm, (int64_t)nnz));
    cudaFree(tmp); tmp = nullptr;
    TRY(cudaMalloc((void**)&flags, n1 * 8));
    TRY(cudaMalloc((void**)&incl, n1 * 8));
    k_flag_heads<<<grid_of(nnz), 256>>>(nnz, keys_out, flags);
    TRY(cudaGetLastError());
    tb = 0;
    TRY(cub::DeviceScan::InclusiveSum(nullptr, tb, flags, incl, (int64_t)nnz));
    TRY(cudaMalloc(&tmp, tb));
    TRY(cub::DeviceScan::InclusiveSum(tmp, tb, flags, incl, (int64_t)nnz));
    TRY(cudaMemcpy(&P->nnz_out, incl + (nnz - 1), 8, cudaMemcpyDeviceToHost));
  }
  TRY(cudaMalloc((void**)&P->offsets, (P->nnz_out + 1) * 8));
  TRY(cudaMalloc((void**)&P->orow, (P->nnz_out > 0 ? P->nnz_out : 1) * 8));
  TRY(cudaMalloc((void**)&P->ocol, (P->nnz_out > 0 ? P->nnz_out : 1) * 8));
  if (nnz > 0) {
    k_offsets<<<grid_of(nnz), 256>>>(nnz, incl, keys_out, minor, P->csc, P->offsets, P->orow, P->ocol);
    TRY(cudaGetLastError());
  }
  TRY(cudaMemcpy(P->offsets + P->nnz_out, &nnz, 8, cudaMemcpyHostToDevice));
  TRY(cudaDeviceSynchronize());
  cleanup();
#undef TRY
  *out = P;
  return PDENLP_OK;
}

int pdenlp_compaction_destroy(pdenlp_compaction* P) {
  if (!P) return PDENLP_OK;
  cudaFree(P->perm); cudaFree(P->offsets); cudaFree(P->orow); cudaFree(P->ocol);
  delete P;
  return PDENLP_OK;
}

int pdenlp_compaction_nnz(const pdenlp_compaction* P, int64_t* nnz_out) {
  if (!P || !nnz_out) return cfail(PDENLP_EINVAL, "null argument");
  *nnz_out = P->nnz_out;
  return PDENLP_OK;
}

int pdenlp_compaction_structure(pdenlp_compaction* P, int64_t* rows, int64_t* cols, int64_t* major_ptr, int32_t on_device) {
  if (!P) return cfail(PDENLP_EINVAL, "null argument");
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (rows) CCU(cudaMemcpy(rows, P->orow, P->nnz_out * 8, kind));
  if (cols) CCU(cudaMemcpy(cols, P->ocol, P->nnz_out * 8, kind));
  if (major_ptr) {
    const int64_t nmajor = P->csc ? P->ncols : P->nrows;
    int64_t* dptr = major_ptr;
    if (!on_device) CCU(cudaMalloc((void**)&dptr, (nmajor + 1) * 8));
    k_major_ptr<<<grid_of(P->nnz_out + 1), 256>>>(P->nnz_out, P->csc ? P->ocol : P->orow, nmajor, dptr);
    CCU(cudaGetLastError());
    if (!on_device) {
      cudaError_t e = cudaMemcpy(major_ptr, dptr, (nmajor + 1) * 8, cudaMemcpyDeviceToHost);
      cudaFree(dptr);
      if (e != cudaSuccess) return cfail(PDENLP_ECUDA, cudaGetErrorString(e));
    }
  }
  // device outputs were written on the legacy default stream; a consumer on a non-blocking stream is not ordered
  // with it, so the call returns only when they are complete
  if (on_device) CCU(cudaDeviceSynchronize());
  return PDENLP_OK;
}

int pdenlp_compaction_apply(pdenlp_compaction* P, const double* coo_vals, double* out_vals, void* stream) {
  if (!P || !coo_vals || !out_vals) return cfail(PDENLP_EINVAL, "null argument");
  if (P->nnz_out == 0) return PDENLP_OK;
  k_apply<<<grid_of(P->nnz_out), 256, 0, (cudaStream_t)stream>>>(P->nnz_out, P->offsets, P->perm, coo_vals, out_vals);
  CCU(cudaGetLastError());
  return PDENLP_OK;
}

}  // extern "C"
