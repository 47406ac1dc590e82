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

// address of the four consecutive columns [col, col + 4) of one row (col a multiple of 4 inside the
// window: slots start at multiples of 4 columns and rows at 16-byte boundaries, so this is one aligned
// 16-byte piece), or nullptr outside the data
__device__ __forceinline__ const float* gram_src4(const GramArgs& a, long long row, int col) {
  if (col >= a.ncols) return nullptr;
  const int gc = col + a.col_lo;
  int s = 0;
  while (s + 1 < a.n_slots && gc >= a.slot_col0[s + 1]) ++s;
  const float* p = a.slot_ptr[s];
  return p ? p + row * a.slot_ld[s] + (gc - a.slot_col0[s]) : nullptr;
}
// 16-byte / 4-byte asynchronous copies global -> shared; a null source zero-fills
__device__ __forceinline__ void gram_cp16(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = src ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src ? src : dst), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gram_cp4(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = src ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src ? src : dst), "r"(bytes) : "memory");
}

// Tile pair (ti <= tj) of 64 x 64 outputs per CTA of 64 threads, 8 x 8 threads with 8 x 8 outputs each:
// 64 DFMA per 8 16-byte shared loads, which keeps the FP64 pipe rather than the shared-memory pipe busy.
// Thread (ty, tx) owns rows 8 ty .. 8 ty + 7 of the left tile and columns
// {16 g + 2 tx, 16 g + 2 tx + 1 : g < 4} of the right one, so that the 8 tx lanes of a quarter warp read
// 128 contiguous bytes and the ty lanes broadcast.
// Row stages of 32 rows travel global -> shared as fp32 with cp.async, issued one stage ahead so that the
// copy of stage s + 1 lands while stage s is being multiplied (the first version loaded synchronously and
// spent 40 % of its samples waiting on those loads); a short pass then applies the row weight and
// converts to float64 once, so the inner loop has no conversions.
constexpr int kGramThreads = 64;
constexpr size_t kGramSmem = 2 * sizeof(double) * kGramKC * kGramTile     // Pa, Pb
                             + 2 * sizeof(float) * kGramKC * kGramTile    // Fa, Fb
                             + 3 * sizeof(float) * kGramKC;               // row weights, responses (x2)
__global__ void __launch_bounds__(kGramThreads) gram_kernel(const GramArgs a) {
  extern __shared__ __align__(16) unsigned char gram_smem[];
  double (*Pa)[kGramTile] = reinterpret_cast<double (*)[kGramTile]>(gram_smem);
  double (*Pb)[kGramTile] = Pa + kGramKC;
  float (*Fa)[kGramTile] = reinterpret_cast<float (*)[kGramTile]>(gram_smem + 2 * sizeof(double) * kGramKC * kGramTile);
  float (*Fb)[kGramTile] = Fa + kGramKC;
  float* Fw = reinterpret_cast<float*>(Fb + kGramKC);
  float* Fy = Fw + kGramKC;
  float* Py = Fy + kGramKC;             // responses of the stage being multiplied
  // tile pair (ti <= tj) from the linear pair index
  int pair = blockIdx.x, ti = 0;
  while (pair >= a.n_tiles - ti) { pair -= a.n_tiles - ti; ++ti; }
  const int tj = ti + pair;
  const int slab = blockIdx.y;
  const long long rows_per_slab = (a.n_rows + a.n_slabs - 1) / a.n_slabs;
  const long long r_begin = (long long)slab * rows_per_slab;
  const long long r_end = r_begin + rows_per_slab < a.n_rows ? r_begin + rows_per_slab : a.n_rows;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  double cs = 0.0, cy = 0.0;       // diagonal tiles: column sum / x^T y of column ti*64 + threadIdx.x
  const bool diag = ti == tj;

  // 32 rows x 16 pieces of 16 bytes per tile, eight per thread
  auto issue = [&](long long r0) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int idx = threadIdx.x + kGramThreads * h;
      const int k = idx >> 4, sc = (idx & 15) * 4;
      const long long row = r0 + k;
      const bool ok = row < r_end;
      gram_cp16(&Fa[k][sc], ok ? gram_src4(a, row, ti * kGramTile + sc) : nullptr);
      if (!diag) gram_cp16(&Fb[k][sc], ok ? gram_src4(a, row, tj * kGramTile + sc) : nullptr);
    }
    if (threadIdx.x < kGramKC) {
      const long long row = r0 + threadIdx.x;
      gram_cp4(&Fw[threadIdx.x], (a.wgt != nullptr && row < r_end) ? a.wgt + row : nullptr);
      gram_cp4(&Fy[threadIdx.x], (a.y != nullptr && row < r_end) ? a.y + row : nullptr);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto masked = [&](float4 v, int col) {      // columns at or beyond ncols are padding
    if (col + 0 >= a.ncols) v.x = 0.f;
    if (col + 1 >= a.ncols) v.y = 0.f;
    if (col + 2 >= a.ncols) v.z = 0.f;
    if (col + 3 >= a.ncols) v.w = 0.f;
    return v;
  };

  if (r_begin < r_end) issue(r_begin);
  for (long long r0 = r_begin; r0 < r_end; r0 += kGramKC) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                 // stage landed for everyone; everyone finished the previous products
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int idx = threadIdx.x + kGramThreads * h;
      const int k = idx >> 4, sc = (idx & 15) * 4;
      const float wv = a.wgt != nullptr ? Fw[k] : 1.f;
      const float4 xa = masked(*reinterpret_cast<const float4*>(&Fa[k][sc]), ti * kGramTile + sc);
      const float4 xb = diag ? xa : masked(*reinterpret_cast<const float4*>(&Fb[k][sc]), tj * kGramTile + sc);
      // the weight rides on the left factor (product formed in fp32, as XS.T * U does in the reference)
      *reinterpret_cast<double2*>(&Pa[k][sc]) = make_double2((double)(xa.x * wv), (double)(xa.y * wv));
      *reinterpret_cast<double2*>(&Pa[k][sc + 2]) = make_double2((double)(xa.z * wv), (double)(xa.w * wv));
      *reinterpret_cast<double2*>(&Pb[k][sc]) = make_double2((double)xb.x, (double)xb.y);
      *reinterpret_cast<double2*>(&Pb[k][sc + 2]) = make_double2((double)xb.z, (double)xb.w);
    }
    if (threadIdx.x < kGramKC) Py[threadIdx.x] = Fy[threadIdx.x];
    __syncthreads();                 // float64 stage complete, fp32 stage free again
    if (r0 + kGramKC < r_end) issue(r0 + kGramKC);
#pragma unroll 2
    for (int k = 0; k < kGramKC; ++k) {
      double av[8], bv[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const double2 p = *reinterpret_cast<const double2*>(&Pa[k][ty * 8 + 2 * g]);
        const double2 q = *reinterpret_cast<const double2*>(&Pb[k][16 * g + 2 * tx]);
        av[2 * g] = p.x; av[2 * g + 1] = p.y;
        bv[2 * g] = q.x; bv[2 * g + 1] = q.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    if (diag) {
      for (int k = 0; k < kGramKC; ++k) {
        const double v = Pa[k][threadIdx.x];
        cs += v;
        cy = fma(v, (double)Py[k], cy);
      }
    }
  }
  double* out = a.part + (size_t)slab * ((size_t)a.ncols * a.ncols + 2 * (size_t)a.ncols);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ti * kGramTile + ty * 8 + i;
      const int c = tj * kGramTile + 16 * (j >> 1) + 2 * tx + (j & 1);
      if (r < a.ncols && c < a.ncols) {
        out[(size_t)r * a.ncols + c] = acc[i][j];
        if (!diag) out[(size_t)c * a.ncols + r] = acc[i][j];
      }
    }
  if (diag) {
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
