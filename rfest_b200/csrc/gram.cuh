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

// ---- the same tile pair on the FP64 tensor-core path (round 2) -------------------------------------------------
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the DFMA rate on B200 (18.5 vs 17.8 T/s measured,
// tools/micro/dmma_rate.cu) but needs one 8-byte shared load per 256 multiply-adds where the register-tile kernel
// above needs eight 16-byte loads per 2048: the FP64 pipe, not operand delivery, becomes the limit.
// 4 warps per CTA.  Off the diagonal warp (wi, wj) owns the 32 x 32 quadrant of the 64 x 64 tile pair = 4 x 4 DMMA
// blocks; on it only block columns >= block row are computed (the rest is the mirror image): warp w takes block
// rows w and 7 - w, 8 - w and w + 1 blocks, nine per warp.  32 accumulator doubles per lane either way.
// With the stage stored [row k][column] both fragments are the same access,
//   A[i = lane / 4][k = lane % 4] = P[k0 + lane % 4][i0 + lane / 4],
// and a row stride of 68 doubles (= 32 B mod 128 B) puts the four k rows a half warp reads (4 columns = 32 B each)
// into four different bank groups: conflict-free 8-byte loads.
// Columns past the window are NOT masked: output (r, c) depends on columns r and c only and outputs outside the
// window are never written, so whatever the padding holds stays in the padding.
constexpr int kGramMmaThreads = 128;
constexpr int kGramMK = 16;                  // rows per stage of the tensor-core kernel
constexpr int kGramLd = kGramTile + 4;
// Shared memory: the float64 stage is double-buffered, so one barrier per stage suffices -- a warp that has
// finished its products converts the next stage into the other buffer while slower warps still multiply.  Every
// thread converts exactly the 16-byte pieces it copied itself (cp.async), so the fp32 staging area needs no
// barrier at all.
constexpr size_t kGramMmaSmem = 2 * 2 * sizeof(double) * kGramMK * kGramLd   // Pa, Pb x 2 buffers
                                + 2 * sizeof(float) * kGramMK * kGramTile    // Fa, Fb
                                + 2 * sizeof(float) * kGramMK;               // responses x 2 buffers
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// one stage of an off-diagonal tile pair: NI x NJ blocks of this warp's quadrant
template <int NI, int NJ>
__device__ __forceinline__ void gram_stage_offdiag(double (&acc)[4][4][2], const double (*Pa)[kGramLd],
                                                   const double (*Pb)[kGramLd], int fk, int ca, int cb) {
#pragma unroll
  for (int k0 = 0; k0 < kGramMK; k0 += 4) {
    double av[NI], bv[NJ];
#pragma unroll
    for (int i = 0; i < NI; ++i) av[i] = Pa[k0 + fk][ca + i * 8];
#pragma unroll
    for (int j = 0; j < NJ; ++j) bv[j] = Pb[k0 + fk][cb + j * 8];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) dmma_8x8x4(acc[i][j], av[i], bv[j]);
  }
}
// one stage of a diagonal tile for warp W: block rows W and 7 - W against block columns j >= row, j < ND.
// acc[2 r + (j >> 2)][j & 3] with r = 0 for row W, 1 for row 7 - W.
template <int W, int ND>
__device__ __forceinline__ void gram_stage_diag(double (&acc)[4][4][2], const double (*Pa)[kGramLd],
                                                const double (*Pb)[kGramLd], int fk, int fc) {
#pragma unroll
  for (int k0 = 0; k0 < kGramMK; k0 += 4) {
    double av[2], bv[8];
    av[0] = Pa[k0 + fk][W * 8 + fc];             // (Pa carries the row weight, Pb does not)
    av[1] = Pa[k0 + fk][(7 - W) * 8 + fc];
#pragma unroll
    for (int j = W; j < ND; ++j) bv[j] = Pb[k0 + fk][j * 8 + fc];
#pragma unroll
    for (int j = W; j < ND; ++j) {
      if (W < ND) dmma_8x8x4(acc[j >> 2][j & 3], av[0], bv[j]);
      if (j >= 7 - W) dmma_8x8x4(acc[2 + (j >> 2)][j & 3], av[1], bv[j]);
    }
  }
}
template <int ND>
__device__ __forceinline__ void gram_stage_diag_w(int warp, double (&acc)[4][4][2], const double (*Pa)[kGramLd],
                                                  const double (*Pb)[kGramLd], int fk, int fc) {
  switch (warp) {
    case 0: gram_stage_diag<0, ND>(acc, Pa, Pb, fk, fc); break;
    case 1: gram_stage_diag<1, ND>(acc, Pa, Pb, fk, fc); break;
    case 2: gram_stage_diag<2, ND>(acc, Pa, Pb, fk, fc); break;
    default: gram_stage_diag<3, ND>(acc, Pa, Pb, fk, fc); break;
  }
}

__global__ void __launch_bounds__(kGramMmaThreads, 4) gram_mma_kernel(const GramArgs a) {
  extern __shared__ __align__(16) unsigned char gram_smem[];
  typedef double (*PStage)[kGramLd];
  PStage P0 = reinterpret_cast<PStage>(gram_smem);            // [buffer][Pa | Pb][row][kGramLd]
  constexpr int kPRows = 2 * kGramMK;                          // rows of one buffer: Pa then Pb
  float (*Fa)[kGramTile] = reinterpret_cast<float (*)[kGramTile]>(gram_smem + 2 * sizeof(double) * kPRows * kGramLd);
  float (*Fb)[kGramTile] = Fa + kGramMK;
  float* Py = reinterpret_cast<float*>(Fb + kGramMK);          // [buffer][row]
  int pair = blockIdx.x, ti = 0;
  while (pair >= a.n_tiles - ti) { pair -= a.n_tiles - ti; ++ti; }
  const int tj = ti + pair;
  const int slab = blockIdx.y;
  const long long rows_per_slab = (a.n_rows + a.n_slabs - 1) / a.n_slabs;
  const long long r_begin = (long long)slab * rows_per_slab;
  const long long r_end = r_begin + rows_per_slab < a.n_rows ? r_begin + rows_per_slab : a.n_rows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wi = warp >> 1, wj = warp & 1;
  const int fk = lane & 3, fc = lane >> 2;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double cs = 0.0, cy = 0.0;
  const bool diag = ti == tj;
  // 8 x 8 blocks that hold real columns (the window's last tile is usually partial): the rest is skipped
  const int ni = min(4, max(0, (a.ncols - ti * kGramTile - wi * 32 + 7) / 8));
  const int nj = min(4, max(0, (a.ncols - tj * kGramTile - wj * 32 + 7) / 8));
  const int nd = min(8, (a.ncols - ti * kGramTile + 7) / 8);

  // a thread owns the four-column piece sc_t of rows k_t and k_t + 8 of every stage: it copies and converts them
  const int k_t = threadIdx.x >> 4, sc_t = (threadIdx.x & 15) * 4;
  const float* src_a = gram_src4(a, 0, ti * kGramTile + sc_t);
  const float* src_b = diag ? nullptr : gram_src4(a, 0, tj * kGramTile + sc_t);
  long long ld_a = 0, ld_b = 0;
  {
    auto ld_of = [&](int col) {
      const int gc = col + a.col_lo;
      int s = 0;
      while (s + 1 < a.n_slots && gc >= a.slot_col0[s + 1]) ++s;
      return a.slot_ld[s];
    };
    if (src_a) ld_a = ld_of(ti * kGramTile + sc_t);
    if (src_b) ld_b = ld_of(tj * kGramTile + sc_t);
  }
  auto issue = [&](long long r0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k_t + 8 * h;
      const long long row = r0 + k;
      const bool ok = row < r_end;
      gram_cp16(&Fa[k][sc_t], (ok && src_a) ? src_a + row * ld_a : nullptr);
      if (!diag) gram_cp16(&Fb[k][sc_t], (ok && src_b) ? src_b + row * ld_b : nullptr);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (r_begin < r_end) issue(r_begin);
  int buf = 0;
  for (long long r0 = r_begin; r0 < r_end; r0 += kGramMK, buf ^= 1) {
    PStage Pa = P0 + (size_t)buf * kPRows, Pb = Pa + kGramMK;
    asm volatile("cp.async.wait_group 0;" ::: "memory");       // this thread's own pieces have landed
    // fp32 -> fp64 of the own pieces (the row weight rides on the left factor, product formed in fp32 as
    // XS.T * U does in the reference)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k_t + 8 * h;
      const long long row = r0 + k;
      const float wv = (a.wgt != nullptr && row < r_end) ? __ldg(a.wgt + row) : 1.f;
      const float4 xa = *reinterpret_cast<const float4*>(&Fa[k][sc_t]);
      const float4 xb = diag ? xa : *reinterpret_cast<const float4*>(&Fb[k][sc_t]);
      *reinterpret_cast<double2*>(&Pa[k][sc_t]) = make_double2((double)(xa.x * wv), (double)(xa.y * wv));
      *reinterpret_cast<double2*>(&Pa[k][sc_t + 2]) = make_double2((double)(xa.z * wv), (double)(xa.w * wv));
      *reinterpret_cast<double2*>(&Pb[k][sc_t]) = make_double2((double)xb.x, (double)xb.y);
      *reinterpret_cast<double2*>(&Pb[k][sc_t + 2]) = make_double2((double)xb.z, (double)xb.w);
    }
    if (diag && threadIdx.x < kGramMK) {
      const long long row = r0 + threadIdx.x;
      Py[buf * kGramMK + threadIdx.x] = (a.y != nullptr && row < r_end) ? __ldg(a.y + row) : 0.f;
    }
    if (r0 + kGramMK < r_end) issue(r0 + kGramMK);             // own pieces only: nobody else reads them
    // ONE barrier per stage: this buffer is complete, and everyone has finished the products of the stage before
    // last, which used the buffer the NEXT conversion will overwrite
    __syncthreads();
    if (!diag) {
      const int ca = wi * 32 + fc, cb = wj * 32 + fc;
      if (ni == 4 && nj == 4) gram_stage_offdiag<4, 4>(acc, Pa, Pb, fk, ca, cb);
      else if (ni == 4 && nj == 1) gram_stage_offdiag<4, 1>(acc, Pa, Pb, fk, ca, cb);
      else if (ni == 4 && nj == 2) gram_stage_offdiag<4, 2>(acc, Pa, Pb, fk, ca, cb);
      else if (ni == 4 && nj == 3) gram_stage_offdiag<4, 3>(acc, Pa, Pb, fk, ca, cb);
      // (ni < 4 only happens in the last tile, which is never the left one of an off-diagonal pair; nj == 0: nothing)
    } else {
      if (nd > 4) gram_stage_diag_w<8>(warp, acc, Pa, Pb, fk, fc);   // (a partial tile runs whole: padding unused)
      else gram_stage_diag_w<4>(warp, acc, Pa, Pb, fk, fc);
      if (threadIdx.x < kGramTile) {
#pragma unroll
        for (int k = 0; k < kGramMK; ++k) {
          const double v = Pa[k][threadIdx.x];
          cs += v;
          cy = fma(v, (double)Py[buf * kGramMK + k], cy);
        }
      }
    }
  }
  double* out = a.part + (size_t)slab * ((size_t)a.ncols * a.ncols + 2 * (size_t)a.ncols);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int r, c;
        bool mirror = true;
        if (!diag) {
          r = ti * kGramTile + wi * 32 + i * 8 + fc;
          c = tj * kGramTile + wj * 32 + j * 8 + 2 * fk + e;
        } else {
          const int br = i < 2 ? warp : 7 - warp, bj = 4 * (i & 1) + j;
          if (bj < br) continue;                       // below the diagonal: never computed
          mirror = bj > br;                            // a diagonal block holds both of its triangles itself
          r = ti * kGramTile + br * 8 + fc;
          c = ti * kGramTile + bj * 8 + 2 * fk + e;
        }
        if (r < a.ncols && c < a.ncols) {
          out[(size_t)r * a.ncols + c] = acc[i][j][e];
          if (mirror) out[(size_t)c * a.ncols + r] = acc[i][j][e];
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
