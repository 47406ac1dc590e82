// Output nonlinearity 'softmax' (GLM.fnl, rfest/GLM/_glm.py:134-140): r = exp(u) / sum(exp(u)) over ALL
// samples of the data set -- the one cross-sample operation of the path.  One evaluation becomes
//   1. a forward sweep with the exponential in place of the output nonlinearity: z_t = exp(u_t) is kept
//      (n floats) and Z = sum_t z_t comes out of the sweep's scalar sums (all-reduced over ranks);
//   2. (gradient only) softmax_stats_kernel over the n-vector: D = sum_t (dL/dr_t) r_t / sum_t r_t, the
//      scalar the softmax Jacobian needs (both sums all-reduced; sum_t r_t is 1 up to the rounding of
//      1 / Z, and dividing by it makes sum_t dL/du_t cancel as it does in exact arithmetic);
//   3. the ordinary sweep with {1 / Z, D}: r_t = exp(u_t) / Z, loss and metric sums as always, and
//      dL/du_t = r_t (dL/dr_t - D).
// Everything stays on the stream (and inside the captured fit iteration).
#pragma once

#include "common.cuh"
#include "sweep.cuh"

namespace rfb {

constexpr int kSoftmaxBlocks = 296;

// part[2 block] = sum over the block's rows of (dL/dr_t) r_t, part[2 block + 1] = sum of r_t, with
// r_t = z_t / Z  (loss.py:4-16)
__global__ void __launch_bounds__(256)
softmax_stats_kernel(const float* __restrict__ z, const float* __restrict__ y, long long n, const double* Zp,
                     int distr, float two_over_n, double* part) {
  __shared__ double sh[8], sh_r[8];
  const float inv_z = (float)(1.0 / *Zp);
  double acc = 0.0, acc_r = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float r = z[t] * inv_z;
    const float yv = y[t];
    float dldr;
    if (distr == RFB_DISTR_POISSON) {
      const bool fin = r != CUDART_INF_F;
      const float r1 = fin ? r : 0.f;
      const bool live = fin && (r1 > 1e-20f);
      const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
      dldr = live ? (1.0f - fast_div(yv, r2)) : 0.f;
    } else {
      dldr = -(yv - r) * two_over_n;
    }
    acc += (double)(dldr * r);
    acc_r += (double)r;
  }
  acc = warp_sum_f64(acc);
  acc_r = warp_sum_f64(acc_r);
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = acc;
    sh_r[threadIdx.x >> 5] = acc_r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, sr = 0.0;
    for (int k = 0; k < 8; ++k) {
      s += sh[k];
      sr += sh_r[k];
    }
    part[2 * blockIdx.x] = s;
    part[2 * blockIdx.x + 1] = sr;
  }
}

// Loss terms of given predictions r (GLM.cost(..., precomputed=r), _glm.py:596-603): per block
// part[2 block] = sum of -log(r') y (Poisson, loss.py:4-10) or (y - r)^2 (Gaussian, loss.py:14-16),
// part[2 block + 1] = sum of r' (Poisson) -- the same fp32 per-sample terms as the sweeps, fp64 sums.
__global__ void __launch_bounds__(256)
loss_terms_kernel(const float* __restrict__ r, const float* __restrict__ y, long long n, int distr, double* part) {
  __shared__ double sh_a[8], sh_b[8];
  double acc_a = 0.0, acc_b = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float rv = r[t];
    const float yv = y[t];
    if (distr == RFB_DISTR_POISSON) {
      const bool fin = rv != CUDART_INF_F;
      const float r1 = fin ? rv : 0.f;
      const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
      acc_a += (double)(-logf(r2) * yv);
      acc_b += (double)r2;
    } else {
      const float err = yv - rv;
      acc_a += (double)(err * err);
    }
  }
  acc_a = warp_sum_f64(acc_a);
  acc_b = warp_sum_f64(acc_b);
  if ((threadIdx.x & 31) == 0) {
    sh_a[threadIdx.x >> 5] = acc_a;
    sh_b[threadIdx.x >> 5] = acc_b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int k = 0; k < 8; ++k) {
      sa += sh_a[k];
      sb += sh_b[k];
    }
    part[2 * blockIdx.x] = sa;
    part[2 * blockIdx.x + 1] = sb;
  }
}

// fixed-order sums of the block partials: out = {sum (dL/dr) r, sum r}
__global__ void softmax_sum_kernel(const double* part, int n_part, double* out) {
  if (threadIdx.x < 2 && blockIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < n_part; ++k) s += part[2 * k + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// sm = {1 / Z, D} in the sweep's arithmetic type; Dp = {sum (dL/dr) r, sum r} or nullptr (forward only)
__global__ void softmax_finish_kernel(const double* Zp, const double* Dp, float* sm) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    sm[0] = (float)(1.0 / *Zp);
    sm[1] = (Dp && Dp[1] != 0.0) ? (float)(Dp[0] / Dp[1]) : 0.f;
  }
}

// ---- 'softmax' as a filter nonlinearity (head_softmax_fix in sweep.cuh) ---------------------------
struct SmxHeads {
  int n_heads;
  int is_smx[kMaxHeads];
};

// prep[tail + S_GHEAD + h] = Z_h (all ranks) -> scale[h] = 1 / Z_h in the sweep's arithmetic type
__global__ void smx_finish_kernel(const double* prep, int tail, SmxHeads hs, float* scale) {
  const int h = threadIdx.x;
  if (blockIdx.x == 0 && h < hs.n_heads) scale[h] = hs.is_smx[h] ? (float)(1.0 / prep[tail + S_GHEAD + h]) : 1.f;
}

// fin: the summed gradient vector of the ordinary sweep, prep: that of the preparatory sweep.  For every softmax
// head  G_h -= D_h E_h / Z_h  with D_h = fin[tail + S_GHEAD + h]; its intercept gradient is exactly 0 (a shift of
// a_h does not change the head's output), reported as such rather than as rounding noise.
__global__ void __launch_bounds__(256)
smx_fix_kernel(double* fin, const double* prep, int ctot, int tail, SmxHeads hs) {
  for (int h = 0; h < hs.n_heads; ++h) {
    if (!hs.is_smx[h]) continue;
    const double k = fin[tail + S_GHEAD + h] / prep[tail + S_GHEAD + h];
    for (int c = threadIdx.x; c < ctot; c += blockDim.x) fin[(size_t)h * ctot + c] -= k * prep[(size_t)h * ctot + c];
  }
  __syncthreads();
  for (int h = threadIdx.x; h < hs.n_heads; h += blockDim.x)
    if (hs.is_smx[h]) fin[tail + S_GHEAD + h] = 0.0;
}

}  // namespace rfb
