// K6: confidence band of the predicted response (GLM._get_response_variance,
// rfest/GLM/_glm.py:1166-1232).
//
// Per filter f with coefficient covariance V_f (p x p) the reference forms
//     y_se_f[t] = sqrt( sum_j (XS_f @ V_f)[t, j] * XS_f[t, j] )  =  sqrt( xs_t^T V_f xs_t )
// and pushes  XS_f b_f + c_f  and  X_f w_f + c_f -/+ 2 y_se_f  through the filter and output
// nonlinearities.  Since w_f = S_f b_f, the raw-design product X_f w_f is XS_f b_f, so the band never
// needs the n x prod(dims) lagged matrix -- only the projected block that already sits in HBM.
//
//   quadform_kernel      lin_f[t] = xs_t . b_f,  se_f[t] = sqrt(xs_t^T V_f xs_t)   (64-row tiles; the
//                        row tile times V_f is a small fp32 GEMM whose result is contracted with
//                        the same rows straight from registers)
//   band_combine_kernel  the three nonlinearity chains per row
#pragma once

#include "common.cuh"
#include "sweep.cuh"

namespace rfb {

constexpr int kQfRows = 64, kQfCols = 64, kQfK = 32;

struct QuadformArgs {
  const float* X;       // block base of the filter's first column, row-major
  long long ld;
  long long n_rows;
  int p;                // coefficients of the filter
  const float* V;       // [p x p] row-major, device
  const float* b;       // [p], device
  float* lin;           // [n_rows]
  float* se;            // [n_rows]
};

__global__ void __launch_bounds__(256) quadform_kernel(const QuadformArgs a) {
  __shared__ float Xs[kQfRows][kQfK + 1];
  __shared__ float Vs[kQfK][kQfCols];
  __shared__ float bs[kQfK];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16 threads, 4 x 4 outputs of X V each
  const long long row0 = (long long)blockIdx.x * kQfRows;
  float q[4] = {0.f, 0.f, 0.f, 0.f};
  float lin[4] = {0.f, 0.f, 0.f, 0.f};

  for (int j0 = 0; j0 < a.p; j0 += kQfCols) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < a.p; k0 += kQfK) {
      for (int e = threadIdx.x; e < kQfRows * kQfK; e += 256) {
        const int r = e / kQfK, k = e % kQfK;
        const long long row = row0 + r;
        Xs[r][k] = (row < a.n_rows && k0 + k < a.p) ? __ldg(a.X + row * a.ld + k0 + k) : 0.f;
      }
      for (int e = threadIdx.x; e < kQfK * kQfCols; e += 256) {
        const int k = e / kQfCols, j = e % kQfCols;
        Vs[k][j] = (k0 + k < a.p && j0 + j < a.p) ? __ldg(a.V + (size_t)(k0 + k) * a.p + j0 + j) : 0.f;
      }
      if (j0 == 0 && threadIdx.x < kQfK) bs[threadIdx.x] = (k0 + threadIdx.x < a.p) ? __ldg(a.b + k0 + threadIdx.x) : 0.f;
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < kQfK; ++k) {
        float xv[4], vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = Xs[ty * 4 + i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) vv[j] = Vs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], vv[j], acc[i][j]);
      }
      if (j0 == 0) {          // the linear part rides on the first column tile: lane tx takes k = tx, tx + 16
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lin[i] = fmaf(Xs[ty * 4 + i][tx], bs[tx], lin[i]);
          lin[i] = fmaf(Xs[ty * 4 + i][tx + 16], bs[tx + 16], lin[i]);
        }
      }
      __syncthreads();
    }
    // contract (X V)[t, j] with X[t, j] for this column tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long row = row0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + tx * 4 + j;
        const float x = (row < a.n_rows && col < a.p) ? __ldg(a.X + row * a.ld + col) : 0.f;
        q[i] = fmaf(acc[i][j], x, q[i]);
      }
    }
  }
  // the 16 lanes of a half warp share ty: fixed-order butterfly
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) {
      q[i] += __shfl_xor_sync(0xffffffffu, q[i], m);
      lin[i] += __shfl_xor_sync(0xffffffffu, lin[i], m);
    }
    const long long row = row0 + ty * 4 + i;
    if (tx == 0 && row < a.n_rows) {
      a.lin[row] = lin[i];
      a.se[row] = sqrtf(q[i]);       // NaN for a negative form, as jnp.sqrt gives
    }
  }
}

struct BandArgs {
  long long n_rows;
  int n_filters;                      // filters with data in this set
  const float* lin[kMaxHeads * 2];    // per filter [n_rows]
  const float* se[kMaxHeads * 2];     // per filter [n_rows], nullptr = no covariance (band of width 0)
  float icpt[kMaxHeads * 2];
  int nl[kMaxHeads * 2];
  float glob_icpt;
  int out_nl;
  float* pred;                        // [n_rows] each
  float* upper;
  float* lower;
  // 'softmax' (a normalisation over ALL samples, _glm.py:134-140) applies to each of the three lines on its
  // own, as the reference's three fnl calls do (_glm.py:1187-1219):
  //   mode 1: block partials of sum exp over the three drives of filter `which`
  //   mode 2: block partials of sum exp over the three output drives (filters already normalised)
  //   mode 0: write the three lines
  int mode, which;
  const float* fscale;                // [3 n_filters] device: 1 / sum for softmax filters
  const float* oscale;                // [3] device: 1 / sum for a softmax output
  double* part;                       // [3 gridDim.x] (modes 1, 2)
};

constexpr int kBandBlocks = 592;

__global__ void __launch_bounds__(256) band_combine_kernel(const BandArgs a) {
  double s[3] = {0.0, 0.0, 0.0};
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n_rows;
       t += (long long)gridDim.x * blockDim.x) {
    float line[3] = {0.f, 0.f, 0.f}, f, d;
    if (a.mode == 1) {
      const int i = a.which;
      const float base = a.lin[i][t] + a.icpt[i];
      const float s2 = a.se[i] ? 2.0f * a.se[i][t] : 0.f;
      s[0] += (double)expf(base);
      s[1] += (double)expf(base + s2);
      s[2] += (double)expf(base - s2);
      continue;
    }
    for (int i = 0; i < a.n_filters; ++i) {
      const float base = a.lin[i][t] + a.icpt[i];
      const float s2 = a.se[i] ? 2.0f * a.se[i][t] : 0.f;
      const bool smx = a.nl[i] == RFB_NL_SOFTMAX;
      nl_eval(a.nl[i], base, f, d);
      line[0] += smx ? f * __ldg(a.fscale + 3 * i) : f;
      nl_eval(a.nl[i], base + s2, f, d);
      line[1] += smx ? f * __ldg(a.fscale + 3 * i + 1) : f;
      nl_eval(a.nl[i], base - s2, f, d);
      line[2] += smx ? f * __ldg(a.fscale + 3 * i + 2) : f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      nl_eval(a.out_nl, line[k] + a.glob_icpt, f, d);
      if (a.mode == 2) s[k] += (double)f;
      else if (a.out_nl == RFB_NL_SOFTMAX) line[k] = f * __ldg(a.oscale + k);
      else line[k] = f;
    }
    if (a.mode == 0) {
      a.pred[t] = line[0];
      a.upper[t] = line[1];
      a.lower[t] = line[2];
    }
  }
  if (a.mode != 0) {
    __shared__ double sh[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v = warp_sum_f64(s[k]);
      if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
      a.part[3 * blockIdx.x + threadIdx.x] = v;
    }
  }
}

// fixed-order sum of the block partials -> sums[3]
__global__ void band_sum_kernel(const double* part, int n_part, double* sums) {
  if (blockIdx.x == 0 && threadIdx.x < 3) {
    double v = 0.0;
    for (int k = 0; k < n_part; ++k) v += part[3 * k + threadIdx.x];
    sums[threadIdx.x] = v;
  }
}
__global__ void band_scale_kernel(const double* sums, float* scale) {
  if (blockIdx.x == 0 && threadIdx.x < 3) scale[threadIdx.x] = (float)(1.0 / sums[threadIdx.x]);
}

}  // namespace rfb
