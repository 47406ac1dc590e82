// K3: trace bookkeeping, elastic-net penalty and the Adam update, one CTA.
//
// Follows rfest/GLM/_glm.py:692-734 (what is stored where), rfest/loss.py:19-23 (penalty),
// rfest/metrics.py:8-20 (scores from sums) and the Adam recurrence of
// jax.example_libraries.optimizers.adam (b1 = 0.9, b2 = 0.999, eps = 1e-8; the iteration
// index passed to the update is 0-based, _glm.py:721).
//
// Trace semantics: the sweep of iteration `it` ran at the parameters
// AFTER update it-1, so it yields cost_train[it] (pre-update `it`) and at the same time
// metric_train[it-1], cost_dev[it-1], metric_dev[it-1] (post-update it-1).
#pragma once

#include "common.cuh"

namespace rfb {

struct UpdateArgs {
  int P;      // filter coefficients
  int F;      // filters; intercepts = F + 1 (last = global)
  int H;      // heads of the sweep
  int ctot;   // padded columns per head
  const int* entry_of;  // [P]  coefficient -> entry (h * ctot + column) in W / the gradient
  const int* head_of;   // [F]  filter -> head
  float* theta;         // [P]
  float* m;
  float* v;
  double* ic;           // [F + 1]
  double* mi;
  double* vi;
  float* W;             // [H * ctot]
  float* head_icpt;     // [H]
  float* glob_icpt;     // [1]
  const double* fin_train;  // [H * ctot + kNumScalars] reduced (and all-reduced) sums
  const double* fin_dev;    // [kNumScalars] or nullptr
  double n_train, sy_train, syy_train;
  double n_dev, sy_dev, syy_dev;
  int distr;
  int metric;
  float alpha, beta;    // elastic net, in the parameter dtype like the reference's weak floats
  int penalise;         // beta != 0
  const double* step_sizes;  // [max_iters]
  // Adam bias corrections per iteration, tabulated by the host at rfb_fit_begin (a double-precision pow
  // per use was the longest dependent chain of the whole update): bc_f[2 it + {0, 1}] =
  // 1 - (float)b^(it + 1) in the parameter dtype for the coefficients, bc_d likewise in double for the
  // float64 intercept leaves
  const float* bc_f;
  const double* bc_d;
  double* cost_train;
  double* cost_dev;
  double* metric_train;
  double* metric_dev;
  float* ring_b;        // [max_iters][P]
  double* ring_ic;      // [max_iters][F + 1]
  int* it;              // device-side iteration counter
  int max_iters;
  int final_only;       // 1: only complete trace entry it-1 (no cost, no update)
  XchgArgs x;           // multi-GPU one-shot exchange (x.world == 1: fin_* already hold the totals)
  double* fin_local;    // [sum_hi] where the exchanged totals are written (= fin_train)
  long long sum_lo, sum_hi;   // entries of the exchanged vector this call needs
};

__device__ inline double score_from_sums(int metric, const double* s, double n, double sy, double syy) {
  const double sr = s[S_R], srr = s[S_RR], syr = s[S_YR], sse = s[S_SQERR];
  switch (metric) {
    case RFB_METRIC_MSE:
      return sse / n;                                          // metrics.py:15-16
    case RFB_METRIC_R2:
      return 1.0 - sse / (syy - sy * sy / n);                  // metrics.py:8-12
    case RFB_METRIC_CORRCOEF: {
      const double cyy = syy - sy * sy / n, crr = srr - sr * sr / n;
      return (syr - sy * sr / n) / sqrt(cyy * crr);            // metrics.py:19-20
    }
    default:
      return CUDART_NAN;
  }
}

__device__ inline double data_loss_from_sums(int distr, const double* s, double n) {
  return distr == RFB_DISTR_POISSON ? s[S_LOSS_A] + s[S_LOSS_B]     // loss.py:8-10
                                    : s[S_LOSS_A] / n;              // loss.py:15
}

constexpr int kUpdateThreads = 256;

__device__ inline double block_sum(double v, double* sm) {
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < kUpdateThreads / 32; ++k) t += sm[k];
  return t;
}

// W / head intercepts from theta / ic (used on its own after set_params, and by the update).
__device__ inline void publish_params(const UpdateArgs& a) {
  for (int j = threadIdx.x; j < a.P; j += blockDim.x) a.W[a.entry_of[j]] = a.theta[j];
  if (threadIdx.x == 0) {
    for (int h = 0; h < a.H; ++h) {
      float s = 0.f;
      for (int f = 0; f < a.F; ++f)
        if (a.head_of[f] == h) s += (float)a.ic[f];
      a.head_icpt[h] = s;
    }
    a.glob_icpt[0] = (float)a.ic[a.F];
  }
}

__global__ void __launch_bounds__(kUpdateThreads) publish_params_kernel(const UpdateArgs a) {
  publish_params(a);
}

__device__ inline unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One iteration's bookkeeping + penalty + Adam by ONE CTA of kUpdateThreads threads.  Shared by the graph
// path (update_kernel) and the persistent small-model kernel (fit_small.cuh), so both leave identical state.
__device__ inline void update_body(const UpdateArgs& a, double* sm, float* s_bc) {
  if (a.x.world > 1) {
    // wait until every rank has published this iteration's vector in my buffer, then add the
    // slots in rank order (identical on all ranks)
    const unsigned long long seq = *a.x.seq + 1ull;
    const int parity = (int)(seq & 1ull);
    unsigned char* mine = a.x.peer[a.x.rank];
    if ((int)threadIdx.x < a.x.world) {
      const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(mine) + parity * kMaxRanks + threadIdx.x;
      while (ld_acquire_sys(flag) < seq) {
      }
    }
    __syncthreads();
    const double* data = reinterpret_cast<const double*>(mine + kXchgDataOffset) + (long long)parity * a.x.world * a.x.cap;
    for (long long e = a.sum_lo + threadIdx.x; e < a.sum_hi; e += blockDim.x) {
      double t = 0.0;
      for (int q = 0; q < a.x.world; ++q) t += __ldcv(data + (long long)q * a.x.cap + e);
      a.fin_local[e] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.x.seq = seq;
  }
  const int it = *a.it;
  const double* ft = a.fin_train;
  const double* sc = ft + (size_t)a.H * a.ctot;

  // monitoring of the previous update: metric_train / cost_dev / metric_dev [it - 1]
  if (threadIdx.x == 0 && it >= 1 && it - 1 < a.max_iters) {
    a.metric_train[it - 1] = score_from_sums(a.metric, sc, a.n_train, a.sy_train, a.syy_train);
    if (a.fin_dev != nullptr) {
      a.cost_dev[it - 1] = data_loss_from_sums(a.distr, a.fin_dev, a.n_dev);
      a.metric_dev[it - 1] = score_from_sums(a.metric, a.fin_dev, a.n_dev, a.sy_dev, a.syy_dev);
    }
  }
  if (a.final_only || it >= a.max_iters) return;

  // elastic net: beta * ((1 - alpha) * ||w||_2 + alpha * ||w||_1)   (loss.py:19-23)
  double l1 = 0.0, l2sq = 0.0;
  if (a.penalise) {
    for (int j = threadIdx.x; j < a.P; j += blockDim.x) {
      const double t = (double)a.theta[j];
      l1 += fabs(t);
      l2sq += t * t;
    }
    l1 = block_sum(l1, sm);
    l2sq = block_sum(l2sq, sm);
  }
  const float l2 = (float)sqrt(l2sq);
  const bool use_l1 = a.penalise && a.alpha != 0.0f;
  const bool use_l2 = a.penalise && a.alpha != 1.0f;
  if (threadIdx.x == 0) {
    double cost = data_loss_from_sums(a.distr, sc, a.n_train);
    if (a.penalise) {
      const float pen = a.beta * ((1.0f - a.alpha) * (use_l2 ? l2 : 0.f) + a.alpha * (use_l1 ? (float)l1 : 0.f));
      cost += (double)pen;
    }
    a.cost_train[it] = cost;
  }
  const float lr = (float)a.step_sizes[it];
  const float b1 = 0.9f, b2 = 0.999f, c1 = (float)(1.0 - 0.9), c2 = (float)(1.0 - 0.999);
  const float eps = 1e-8f;
  const float bc1 = __ldg(a.bc_f + 2 * it), bc2 = __ldg(a.bc_f + 2 * it + 1);
  const float wl1 = a.beta * a.alpha, wl2 = a.beta * (1.0f - a.alpha);

  for (int j = threadIdx.x; j < a.P; j += blockDim.x) {
    const float t = a.theta[j];
    float g = (float)ft[a.entry_of[j]];
    if (use_l1) g += wl1 * ((t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f));
    if (use_l2) g += wl2 * (t / l2);
    const float mj = c1 * g + b1 * a.m[j];
    const float vj = c2 * (g * g) + b2 * a.v[j];
    const float mhat = mj / bc1;
    const float vhat = vj / bc2;
    const float tn = t - lr * mhat / (sqrtf(vhat) + eps);
    a.m[j] = mj;
    a.v[j] = vj;
    a.theta[j] = tn;
    a.ring_b[(size_t)it * a.P + j] = tn;
  }
  // intercepts: float64 leaves (weak-typed Python floats in the reference), fp32 gradients
  if (threadIdx.x <= a.F) {
    const int k = threadIdx.x;
    const double g = (k == a.F) ? (double)(float)sc[S_GGLOBAL]
                                : (double)(float)sc[S_GHEAD + a.head_of[k]];
    const double mk = (1.0 - 0.9) * g + 0.9 * a.mi[k];
    const double vk = (1.0 - 0.999) * (g * g) + 0.999 * a.vi[k];
    const double mhat = mk / __ldg(a.bc_d + 2 * it);
    const double vhat = vk / __ldg(a.bc_d + 2 * it + 1);
    const double xn = a.ic[k] - a.step_sizes[it] * mhat / (sqrt(vhat) + 1e-8);
    a.mi[k] = mk;
    a.vi[k] = vk;
    a.ic[k] = xn;
    a.ring_ic[(size_t)it * (a.F + 1) + k] = xn;
  }
  __syncthreads();
  publish_params(a);
  if (threadIdx.x == 0) *a.it = it + 1;
}

__global__ void __launch_bounds__(kUpdateThreads) update_kernel(const UpdateArgs a) {
  __shared__ double sm[kUpdateThreads / 32];
  __shared__ float s_bc[2];
  update_body(a, sm, s_bc);
}

}  // namespace rfb
