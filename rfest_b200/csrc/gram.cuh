// K5: Gram matrix of the concatenated design, for the least-squares initialisation
// (`compute_mle`, rfest/GLM/_glm.py:438-537: XtX = A^T A, Xty = A^T y on the ones-augmented
// column stack; the ones columns only need column sums, so the device computes
//     G = XS^T XS  (Ctot x Ctot),   colsum = XS^T 1,   xty = XS^T y
// and the host assembles the normal equations).  The reference forms them in float64 (its ones
// are float64 under jax_enable_x64), so products and sums are float64 here as well: fp32 inputs,
// DFMA accumulation, fixed-order reduction over row slabs -- deterministic.
#pragma once

#include "common.cuh"

namespace rfb {

constexpr int kGramTile = 64;   // columns per tile side
constexpr int kGramKC = 32;     // rows per shared-memory stage

struct GramArgs {
  const float* slot_ptr[kMaxSlots];
  long long slot_ld[kMaxSlots];
  int slot_col0[kMaxSlots + 1];   // prefix sums of (padded) columns per slot
  int n_slots;
  int ctot;
  long long n_rows;
  const float* y;                 // may be nullptr
  const float* wgt;               // optional per-row weights: G = XS^T diag(w) XS (colsum, xty weighted too)
  int col_lo;                     // column window [col_lo, col_lo + ncols) of the concatenated design
  int ncols;
  int n_tiles;                    // ceil(ncols / 64)
  int n_pairs;                    // n_tiles * (n_tiles + 1) / 2
  int n_slabs;
  double* part;                   // [n_slabs][ctot * ctot + 2 * ctot]
};

__device__ __forceinline__ float gram_load(const GramArgs& a, long long row, int col) {
  if (col >= a.ncols || row >= a.n_rows) return 0.f;
  col += a.col_lo;
  int s = 0;
  while (s + 1 < a.n_slots && col >= a.slot_col0[s + 1]) ++s;
  const float* p = a.slot_ptr[s];
  return p ? __ldg(p + row * a.slot_ld[s] + (col - a.slot_col0[s])) : 0.f;
}

__global__ void __launch_bounds__(256) gram_kernel(const GramArgs a) {
  __shared__ float Pa[kGramKC][kGramTile + 1];
  __shared__ float Pb[kGramKC][kGramTile + 1];
  __shared__ float Py[kGramKC];
  // tile pair (ti <= tj) from the linear pair index
  int pair = blockIdx.x, ti = 0;
  while (pair >= a.n_tiles - ti) { pair -= a.n_tiles - ti; ++ti; }
  const int tj = ti + pair;
  const int slab = blockIdx.y;
  const long long rows_per_slab = (a.n_rows + a.n_slabs - 1) / a.n_slabs;
  const long long r_begin = (long long)slab * rows_per_slab;
  const long long r_end = r_begin + rows_per_slab < a.n_rows ? r_begin + rows_per_slab : a.n_rows;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double cs = 0.0, cy = 0.0;       // diagonal tiles: column sum / x^T y of column ti*64 + threadIdx.x (< 64)
  const bool diag = ti == tj;

  for (long long r0 = r_begin; r0 < r_end; r0 += kGramKC) {
    for (int e = threadIdx.x; e < kGramKC * kGramTile; e += 256) {
      const int k = e / kGramTile, c = e % kGramTile;
      const long long row = r0 + k;
      const bool ok = row < r_end;
      const float wv = (a.wgt != nullptr && ok) ? __ldg(a.wgt + row) : 1.f;
      const float xa = ok ? gram_load(a, row, ti * kGramTile + c) : 0.f;
      Pa[k][c] = xa * wv;                                   // the weight rides on the left factor
      Pb[k][c] = diag ? xa : (ok ? gram_load(a, row, tj * kGramTile + c) : 0.f);
    }
    if (threadIdx.x < kGramKC) {
      const long long row = r0 + threadIdx.x;
      Py[threadIdx.x] = (a.y != nullptr && row < r_end) ? __ldg(a.y + row) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kGramKC; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = (double)Pa[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = (double)Pb[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    if (diag && threadIdx.x < kGramTile) {
      for (int k = 0; k < kGramKC; ++k) {
        const double v = (double)Pa[k][threadIdx.x];
        cs += v;
        cy = fma(v, (double)Py[k], cy);
      }
    }
    __syncthreads();
  }
  double* out = a.part + (size_t)slab * ((size_t)a.ncols * a.ncols + 2 * (size_t)a.ncols);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ti * kGramTile + ty * 4 + i, c = tj * kGramTile + tx * 4 + j;
      if (r < a.ncols && c < a.ncols) {
        out[(size_t)r * a.ncols + c] = acc[i][j];
        if (!diag) out[(size_t)c * a.ncols + r] = acc[i][j];
      }
    }
  if (diag && threadIdx.x < kGramTile) {
    const int c = ti * kGramTile + threadIdx.x;
    if (c < a.ncols) {
      out[(size_t)a.ncols * a.ncols + c] = cs;
      out[(size_t)a.ncols * a.ncols + a.ncols + c] = cy;
    }
  }
}

// w <- 1 / r^2 in place (the weights of the Poisson filter covariance, _glm.py:1129-1131)
__global__ void __launch_bounds__(256) inv_square_kernel(float* r, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = r[i];
    r[i] = 1.0f / (v * v);
  }
}

// out[e] = sum over slabs, fixed order
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* part, int n_slabs, long long n_entries, double* out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  double s = 0.0;
  for (int k = 0; k < n_slabs; ++k) s += part[(size_t)k * n_entries + e];
  out[e] = s;
}

}  // namespace rfb
