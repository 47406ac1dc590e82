// K1: fused time-lag + spline projection, never materialising the lagged (Hankel) design.
//
// Replaces build_design_matrix(...)[burn_in:] @ build_spline_matrix(...)
// (rfest/utils.py:5-67, rfest/splines.py:278-305, rfest/GLM/_glm.py:250-258,297-298).
// Because S = kron(St, Sxy), the product separates:
//   K1a  Z[s, q]            = sum_p stim[s, p] * Sxy[p, q]            (spatial contraction)
//   K1b  XS[t, a*n_q + q]   = sum_i Z[t + i - front, q] * St[i, a]    (temporal, Hankel)
// The integer lag/shift/padding map (which raw row feeds design row t, lag i) is exactly the
// reference's: src = t + i - front, front = nlag + shift - 1 (0 when nlag + shift <= 0), zero
// outside [0, n_raw_total).
#pragma once

#include "common.cuh"

namespace rfb {

// ---- K1a: Z = stim @ Sxy, fp32 FMA, shared-memory tiled ------------------------------------------
// Block tile: 64 rows x 128 columns of Z; K tile 32.  Thread (ty, tx), 16 x 32 threads:
// rows ty*4..ty*4+3, columns tx + 32*j (j < 4).
constexpr int kSpTM = 64, kSpTN = 128, kSpTK = 32;
__global__ void __launch_bounds__(512)
spatial_project_kernel(const float* __restrict__ stim, long long n_rows, int n_pix,
                       const float* __restrict__ Sxy, int n_q, float* __restrict__ Z, int ldz) {
  __shared__ float As[kSpTM][kSpTK + 1];
  __shared__ float Bs[kSpTK][kSpTN];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long row0 = (long long)blockIdx.x * kSpTM;
  const int q0 = blockIdx.y * kSpTN;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;

  for (int k0 = 0; k0 < n_pix; k0 += kSpTK) {
    // stim tile: 64 x 32, thread loads 4 elements, coalesced along k
    for (int e = threadIdx.x; e < kSpTM * kSpTK; e += blockDim.x) {
      const int r = e / kSpTK, k = e % kSpTK;
      const long long row = row0 + r;
      As[r][k] = (row < n_rows && k0 + k < n_pix) ? stim[row * n_pix + k0 + k] : 0.f;
    }
    for (int e = threadIdx.x; e < kSpTK * kSpTN; e += blockDim.x) {
      const int k = e / kSpTN, q = e % kSpTN;
      Bs[k][q] = (k0 + k < n_pix && q0 + q < n_q) ? Sxy[(size_t)(k0 + k) * n_q + q0 + q] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kSpTK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = As[ty * 4 + r][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx + 32 * j];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(av[r], bv[j], acc[r][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long row = row0 + ty * 4 + r;
    if (row >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + tx + 32 * j;
      if (q < n_q) Z[row * ldz + q] = acc[r][j];
    }
  }
}

// ---- K1b: temporal Hankel contraction ---------------------------------------------------------------
// Work item = (group of TT consecutive design rows, one q).  Z rows are addressed relative to
// z_row0 (the raw row held in Z's first row); raw rows outside [0, n_raw_total) contribute zero.
struct TemporalArgs {
  const float* Z;        // [n_z x ldz]
  long long z_src0;      // raw row index of Z row 0
  long long n_z;
  int ldz;
  long long n_raw_total;
  float* out;            // block storage
  long long out_ld;
  long long out_row0;    // first block row to write
  long long n_out;
  int col0;
  long long first_t;     // design row (untrimmed) of block row out_row0
  long long front;       // zero rows in front (nlag + shift - 1 or 0)
  int nlag;
  int df_t;
  int n_q;
  const float* St;       // [nlag x DFT] padded with zero columns, device
};

template <int DFT, int TT>
__global__ void __launch_bounds__(128) temporal_project_kernel(const TemporalArgs a) {
  const long long n_groups = (a.n_out + TT - 1) / TT;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_groups * a.n_q) return;
  const long long grp = item / a.n_q;
  const int q = (int)(item - grp * a.n_q);
  const long long k0 = grp * TT;               // first row (relative to out_row0) of this group

  float acc[TT][DFT];
#pragma unroll
  for (int tt = 0; tt < TT; ++tt)
#pragma unroll
    for (int d = 0; d < DFT; ++d) acc[tt][d] = 0.f;

  for (int i = 0; i < a.nlag; ++i) {
    float st[DFT];
#pragma unroll
    for (int d = 0; d < DFT; ++d) st[d] = __ldg(a.St + i * DFT + d);
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
      const long long src = a.first_t + k0 + tt + i - a.front;     // raw row feeding (t, lag i)
      float z = 0.f;
      if (src >= 0 && src < a.n_raw_total) {
        const long long zr = src - a.z_src0;
        if (zr >= 0 && zr < a.n_z) z = __ldg(a.Z + zr * a.ldz + q);
      }
#pragma unroll
      for (int d = 0; d < DFT; ++d) acc[tt][d] = fmaf(z, st[d], acc[tt][d]);
    }
  }
#pragma unroll
  for (int tt = 0; tt < TT; ++tt) {
    const long long k = k0 + tt;
    if (k >= a.n_out) break;
    float* dst = a.out + (a.out_row0 + k) * a.out_ld + a.col0 + q;
#pragma unroll
    for (int d = 0; d < DFT; ++d)
      if (d < a.df_t) dst[(size_t)d * a.n_q] = acc[tt][d];
  }
}

// Faster variant for nlag <= 64: the temporal basis sits in constant memory (zero rows past nlag), and the
// lag loop is unrolled by 8 over a 15-value register window of Z, of which only 8 values are new per
// step: one global load per 64 FMAs instead of 16.  A thread whose whole window of raw rows is held in Z
// (all but the first and last few of a recording) takes the FAST path: no bounds checks, one pointer
// walking down Z's column -- in the first version a quarter of all instructions were the 64-bit index
// arithmetic and range tests of the loads.  Both paths add the lags in the same order, so results do not
// depend on which one a row takes (sharded == unsharded bit for bit).  The last, partial step of 8 lags
// only issues the lags that exist, and the 8 values a step adds to the window are loaded one step ahead.
constexpr int kTempMaxLag = 64;
__constant__ float c_st[kTempMaxLag * 16];

// Packed arithmetic: the accumulators of two neighbouring basis columns (d, d + 1) share a 64-bit register
// pair and are updated by one fma.rn.f32x2 (SASS FFMA2) -- two IEEE fused multiply-adds, bit-identical to the
// scalar ones, for one issue slot.  The kernel was issue-bound (65 % of its instructions were FFMAs at 75 %
// issue utilisation); halving the FMA instructions puts it on the FMA pipe itself.
// `zload(k)` returns the value of the thread's column in the k-th raw row of its window (k = 0 .. nlag + 6).
// (The basis is read from constant memory through uniform registers; a shared-memory copy feeding ordinary
// registers was tried and measured no faster.)
template <int DFT, class Load>
__device__ __forceinline__ void temporal_accumulate_from(int nlag, Load zload, float (&acc)[8][DFT]) {
  constexpr int TT = 8;
  constexpr bool PACKED = (DFT % 2) == 0;
  constexpr int DP = PACKED ? DFT / 2 : 1;
  u64_t acc2[TT][DP];
  if (PACKED) {
#pragma unroll
    for (int tt = 0; tt < TT; ++tt)
#pragma unroll
      for (int dp = 0; dp < DP; ++dp) acc2[tt][dp] = 0ull;
  }
  // one lag: acc[tt][:] += w[tt + j] * st[j][:]
  auto lag = [&](const float (&w)[2 * TT - 1], const float* cs, int j) {
    if (PACKED) {
      u64_t ww[TT];
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) ww[tt] = pack2(w[tt + j], w[tt + j]);
#pragma unroll
      for (int dp = 0; dp < DP; ++dp) {
        const float2 sv = *reinterpret_cast<const float2*>(cs + j * DFT + 2 * dp);
        const u64_t st = pack2(sv.x, sv.y);
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) acc2[tt][dp] = ffma2(ww[tt], st, acc2[tt][dp]);
      }
    } else {
#pragma unroll
      for (int d = 0; d < DFT; ++d) {
        const float st = cs[j * DFT + d];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) acc[tt][d] = fmaf(w[tt + j], st, acc[tt][d]);
      }
    }
  };

  float w[2 * TT - 1];                              // Z[src0 + i0 + k], k = 0 .. 14
  float nx[TT];                                     // the 8 values the NEXT step adds, loaded one step ahead
  const int last = nlag + TT - 2;                   // last raw row (relative) any lag of this thread reads
#pragma unroll
  for (int k = 0; k < TT - 1; ++k) w[k] = zload(k);
#pragma unroll
  for (int k = 0; k < TT; ++k) nx[k] = zload(min(TT - 1 + k, last));
  int i0 = 0;
#pragma unroll 1
  for (; i0 + TT <= nlag; i0 += TT) {
#pragma unroll
    for (int k = 0; k < TT; ++k) w[TT - 1 + k] = nx[k];
#pragma unroll
    for (int k = 0; k < TT; ++k) nx[k] = zload(min(i0 + 2 * TT - 1 + k, last));
    const float* cs = c_st + i0 * DFT;
#pragma unroll
    for (int j = 0; j < TT; ++j) lag(w, cs, j);
#pragma unroll
    for (int k = 0; k < TT - 1; ++k) w[k] = w[k + TT];
  }
  const int rem = nlag - i0;                        // 0 .. 7 lags left
  if (rem > 0) {
#pragma unroll
    for (int k = 0; k < TT - 1; ++k) w[TT - 1 + k] = nx[k];
    const float* cs = c_st + i0 * DFT;
#pragma unroll
    for (int j = 0; j < TT - 1; ++j)
      if (j < rem) lag(w, cs, j);
  }
  if (PACKED) {
#pragma unroll
    for (int tt = 0; tt < TT; ++tt)
#pragma unroll
      for (int dp = 0; dp < DP; ++dp) unpack2(acc2[tt][dp], acc[tt][2 * dp], acc[tt][2 * dp + 1]);
  }
}

template <int DFT, bool FAST>
__device__ __forceinline__ void temporal_accumulate(const TemporalArgs& a, long long src0, int q, float (&acc)[8][DFT]) {
  const float* zp = nullptr;
  const long long ldz = a.ldz;
  if (FAST) zp = a.Z + (src0 - a.z_src0) * ldz + q;
  auto zload = [&](int k) -> float {                // raw row src0 + k of column q
    if (FAST) return __ldg(zp + k * ldz);
    const long long src = src0 + k;
    if (src < 0 || src >= a.n_raw_total) return 0.f;
    const long long zr = src - a.z_src0;
    return (zr >= 0 && zr < a.n_z) ? __ldg(a.Z + zr * ldz + q) : 0.f;
  };
  temporal_accumulate_from<DFT>(a.nlag, zload, acc);
}

template <int DFT>
__global__ void __launch_bounds__(128) temporal_project_const_kernel(const TemporalArgs a) {
  constexpr int TT = 8;
  const long long n_groups = (a.n_out + TT - 1) / TT;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_groups * a.n_q) return;
  const long long grp = item / a.n_q;
  const int q = (int)(item - grp * a.n_q);
  const long long k0 = grp * TT;
  const long long src0 = a.first_t + k0 - a.front;     // raw row feeding (row k0, lag 0)

  float acc[TT][DFT];
#pragma unroll
  for (int tt = 0; tt < TT; ++tt)
#pragma unroll
    for (int d = 0; d < DFT; ++d) acc[tt][d] = 0.f;
  // raw rows [src0, src0 + nlag + TT - 1) feed this thread's 8 design rows
  const long long n_win = a.nlag + TT - 1;
  const long long zr0 = src0 - a.z_src0;
  if (src0 >= 0 && src0 + n_win <= a.n_raw_total && zr0 >= 0 && zr0 + n_win <= a.n_z)
    temporal_accumulate<DFT, true>(a, src0, q, acc);
  else
    temporal_accumulate<DFT, false>(a, src0, q, acc);

  float* dst = a.out + (a.out_row0 + k0) * a.out_ld + a.col0 + q;
  const bool full = k0 + TT <= a.n_out;
#pragma unroll
  for (int tt = 0; tt < TT; ++tt) {
    if (!full && k0 + tt >= a.n_out) break;
#pragma unroll
    for (int d = 0; d < DFT; ++d)
      if (d < a.df_t) dst[(size_t)d * a.n_q] = acc[tt][d];
    dst += a.out_ld;
  }
}

// ---- block @ matrix: out[r, j] = sum_c X[r, c] * B[c, j] for K <= 8 columns of B at a time ------------------
// (`model.XS[kind][name] @ b` of user code, rfest/GLM/_glm.py:562,1127: one warp per row, float4 loads,
// B through the read-only cache; HBM-bound: the block is read once per 8 columns of B)
template <int K>
__global__ void __launch_bounds__(256)
block_matmul_kernel(const float* __restrict__ X, long long n_rows, int n_cols, long long ld,
                    const float* __restrict__ B, int ldb, int j0, int kw, float* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nch = (n_cols + 3) / 4;                       // rows are padded to a multiple of 4 floats
  for (long long r = warp; r < n_rows; r += n_warps) {
    const float4* row = reinterpret_cast<const float4*>(X + r * ld);
    float acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.f;
    for (int c4 = lane; c4 < nch; c4 += 32) {
      const float4 x = __ldg(row + c4);
      const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c4 * 4 + u;
        if (c < n_cols) {
#pragma unroll
          for (int j = 0; j < K; ++j)
            if (j < kw) acc[j] = fmaf(xs[u], __ldg(B + (size_t)c * ldb + j0 + j), acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float v = acc[j];
      for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (lane == 0 && j < kw) out[r * ldo + j0 + j] = v;
    }
  }
}

// sum y, sum y^2 in fp64 (constants of the metrics, rfest/metrics.py:8-12)
__global__ void __launch_bounds__(256) y_moments_kernel(const float* y, long long n, double* out2) {
  __shared__ double s0[8], s1[8];
  double a = 0.0, b = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)y[i];
    a += v;
    b += v * v;
  }
  for (int m = 16; m >= 1; m >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, m);
    b += __shfl_xor_sync(0xffffffffu, b, m);
  }
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < 8; ++k) { ta += s0[k]; tb += s1[k]; }
    out2[2 * blockIdx.x] = ta;
    out2[2 * blockIdx.x + 1] = tb;
  }
}

}  // namespace rfb
