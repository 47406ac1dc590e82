// Kp: the whole fit loop of a SMALL model as ONE persistent cooperative kernel.
//
// BASELINE config 1 (the reference README recipe: ~20k samples x 16 spline columns, 1000 iterations) is
// pure launch latency on the graph path: four kernels of 8-12 us for 1.3 MB of data (22-25 us per
// iteration).  Here the rows of the train and dev sets are loaded into shared memory ONCE, every CTA keeps
// its slice resident, and an iteration is
//     sweep over the resident rows (forward, loss, gradient, metric sums; dev: forward only)
//     -> fixed-order fp64 reduction inside the CTA -> per-CTA partial row in global memory
//     -> ONE grid barrier (cooperative groups; a flag-per-CTA exchange is kept as an option) -> every CTA adds the partial rows in CTA order and applies the bookkeeping +
//     penalty + Adam arithmetic of update_body (update.cuh) to ITS OWN shared-memory copy of the optimiser
//     state; the copies stay bit-identical, so nobody waits for a broadcast and the next sweep starts at
//     once.  CTA 0 records the traces and the parameter ring; the partial rows alternate between two
//     parities, which is what makes one barrier per iteration enough.
// No relaunch, no atomics in the sums (bit-reproducible), same trace semantics (rfest/GLM/_glm.py:720-734).
// The state is read from and written back to the device buffers the graph path uses, so a fit may switch
// between the two paths between batches of iterations.
//
// Mapping: a row is owned by LPR lanes (the next power of two >= chunks per row, <= 16: up to 64 design
// columns); lane = (row slot, float4 chunk).  Per-row arithmetic is the row_epilogue of sweep.cuh.
#pragma once

#include <cooperative_groups.h>

#include "sweep.cuh"
#include "update.cuh"

namespace rfb {

namespace cg = cooperative_groups;

constexpr int kSmallThreads = kUpdateThreads;
constexpr int kSmallMaxChunks = 16;

struct SmallArgs {
  SweepArgs tr, dv;            // dv.n_rows == 0: no dev set
  UpdateArgs ua;               // the fit's state in global memory (read at entry, written back at exit)
  double* part;                // [2 parities][grid][NE + kNumScalars] per-CTA partial sums
  unsigned long long* flags;   // [grid]: epoch + (iterations of this launch whose partial row is complete)
  unsigned long long epoch;    // iterations launched before this batch (flags only ever grow)
  int n_iters;
  int rows_tr, rows_dv;        // rows per CTA
  int NE;                      // H * ctot + kNumScalars
};

// dynamic shared memory layout (all CTAs identical), in this order:
//   rows of the train / dev slices, their responses, W image [H][ctot], head + global intercepts,
//   optimiser state (theta, m, v [P] float; entry_of [P] int; ic, mi, vi [F + 1] double; head_of [F] int),
//   totals fin [NE + kNumScalars] double, cross-warp reduction scratch [warps][NE] double
__host__ __device__ inline size_t fit_small_smem_bytes(int rows_tr, int rows_dv, int ctot, int H, int P, int F) {
  const size_t rows = (size_t)rows_tr + rows_dv;
  size_t b = rows * ctot * sizeof(float);
  b += ((size_t)rows_tr + 3) / 4 * 16 + ((size_t)rows_dv + 3) / 4 * 16;
  b += (size_t)H * ctot * sizeof(float) + 64;
  b += (size_t)4 * ((P + 3) / 4 * 4) * 4;                       // theta, m, v, entry_of
  b += (size_t)((F + 1 + 3) / 4 * 4) * 4;                       // head_of
  b = (b + 15) & ~(size_t)15;
  b += (size_t)3 * (F + 2) * sizeof(double);                    // ic, mi, vi
  const size_t NV = (size_t)H * ctot + kNumScalars;
  b += (NV + kNumScalars) * sizeof(double);                     // fin
  b += (size_t)(kSmallThreads / 32) * NV * sizeof(double);      // reduction scratch
  return b + 16;
}

__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int H>
__global__ void __launch_bounds__(kSmallThreads, 1) fit_small_kernel(const __grid_constant__ SmallArgs a) {
  extern __shared__ __align__(16) unsigned char small_smem[];
  __shared__ double upd_sm[kUpdateThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = kSmallThreads / 32;
  const UpdateArgs& u = a.ua;
  const int NC = a.tr.n_chunks, ctot = NC * 4;
  const int NV = H * ctot + kNumScalars;
  const int P = u.P, F = u.F, P4 = (P + 3) / 4 * 4;
  const bool has_dev = a.dv.n_rows > 0;
  const bool cta0 = blockIdx.x == 0;

  float* xs_tr = reinterpret_cast<float*>(small_smem);
  float* xs_dv = xs_tr + (size_t)a.rows_tr * ctot;
  float* y_tr = xs_dv + (size_t)a.rows_dv * ctot;
  float* y_dv = y_tr + (a.rows_tr + 3) / 4 * 4;
  float* Wsm = y_dv + (a.rows_dv + 3) / 4 * 4;
  float* icpt = Wsm + H * ctot;                                  // [H] head intercepts, then the global one
  float* theta = icpt + 16;
  float* am = theta + P4;
  float* av = am + P4;
  int* entry_of = reinterpret_cast<int*>(av + P4);
  int* head_of = entry_of + P4;
  double* ic = reinterpret_cast<double*>(
      small_smem + ((reinterpret_cast<unsigned char*>(head_of + (F + 1 + 3) / 4 * 4) - small_smem + 15) & ~(size_t)15));
  double* mi = ic + (F + 2);
  double* vi = mi + (F + 2);
  double* fin = vi + (F + 2);
  double* red = fin + NV + kNumScalars;

  // ---- this CTA's rows -> shared memory (once) ------------------------------------------------------
  const long long tr0 = (long long)blockIdx.x * a.rows_tr, dv0 = (long long)blockIdx.x * a.rows_dv;
  const int n_tr = (int)max(0LL, min((long long)a.rows_tr, a.tr.n_rows - tr0));
  const int n_dv = (int)max(0LL, min((long long)a.rows_dv, a.dv.n_rows - dv0));
  for (int set = 0; set < 2; ++set) {
    const SweepArgs& d = set == 0 ? a.tr : a.dv;
    const int nr = set == 0 ? n_tr : n_dv;
    const long long r0 = set == 0 ? tr0 : dv0;
    float* xs = set == 0 ? xs_tr : xs_dv;
    float* ys = set == 0 ? y_tr : y_dv;
    for (int s = 0; s < d.n_slots; ++s) {
      const int c0 = d.slot_chunk0[s], cw = d.slot_chunk0[s + 1] - c0;
      for (int i = tid; i < nr * cw; i += kSmallThreads) {
        const int r = i / cw, c = i - r * cw;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.slot_ptr[s] != nullptr)
          v = *reinterpret_cast<const float4*>(d.slot_ptr[s] + (r0 + r) * d.slot_ld[s] + 4 * c);
        *reinterpret_cast<float4*>(xs + (size_t)r * ctot + 4 * (c0 + c)) = v;
      }
    }
    for (int i = tid; i < nr; i += kSmallThreads) ys[i] = d.y != nullptr ? d.y[r0 + i] : 0.f;
  }
  // ---- optimiser state -> shared memory: every CTA carries an identical copy -------------------------
  for (int j = tid; j < P; j += kSmallThreads) {
    theta[j] = u.theta[j];
    am[j] = u.m[j];
    av[j] = u.v[j];
    entry_of[j] = u.entry_of[j];
  }
  if (tid <= F) {
    ic[tid] = u.ic[tid];
    mi[tid] = u.mi[tid];
    vi[tid] = u.vi[tid];
    if (tid < F) head_of[tid] = u.head_of[tid];
  }
  for (int i = tid; i < H * ctot; i += kSmallThreads) Wsm[i] = 0.f;
  const int it0 = *u.it;
  __syncthreads();
  auto publish = [&]() {                                         // publish_params (update.cuh) into shared memory
    for (int j = tid; j < P; j += kSmallThreads) Wsm[entry_of[j]] = theta[j];
    if (tid == 0) {
      for (int h = 0; h < H; ++h) {
        float s = 0.f;
        for (int f = 0; f < F; ++f)
          if (head_of[f] == h) s += (float)ic[f];
        icpt[h] = s;
      }
      icpt[H] = (float)ic[F];
    }
    __syncthreads();
  };
  publish();

  // lanes per row, this lane's place
  int LPR = 1;
  while (LPR < NC) LPR <<= 1;
  const int chunk = lane & (LPR - 1), slot = lane / LPR, RPW = 32 / LPR;      // rows per warp pass
  const bool chunk_ok = chunk < NC;
  const bool poisson = a.tr.distr == RFB_DISTR_POISSON;
  const size_t prow = (size_t)a.NE + kNumScalars;

  for (int k = 0; k < a.n_iters; ++k) {
    const int it = it0 + k;
    float4 w[H];
    float hic[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      w[h] = chunk_ok ? *reinterpret_cast<const float4*>(Wsm + h * ctot + 4 * chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
      hic[h] = icpt[h];
    }
    const float gic = icpt[H];
    double* out = a.part + ((size_t)(k & 1) * gridDim.x + blockIdx.x) * prow;

    for (int set = 0; set < 2; ++set) {
      const bool bwd = set == 0;
      const SweepArgs& d = bwd ? a.tr : a.dv;
      const int nr = bwd ? n_tr : n_dv;
      if (!bwd && !has_dev) break;
      const float* xs = bwd ? xs_tr : xs_dv;
      const float* ys = bwd ? y_tr : y_dv;
      float4 g[H];
      float sc[kNumScalars];
#pragma unroll
      for (int h = 0; h < H; ++h) g[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kNumScalars; ++q) sc[q] = 0.f;
      const int passes = (nr + NW * RPW - 1) / (NW * RPW);
      for (int ps = 0; ps < passes; ++ps) {
        const int r = (ps * NW + wid) * RPW + slot;
        const bool rowok = r < nr;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rowok && chunk_ok) x = *reinterpret_cast<const float4*>(xs + (size_t)r * ctot + 4 * chunk);
        float tot[H], dh[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float pdot = fmaf(x.x, w[h].x, fmaf(x.y, w[h].y, fmaf(x.z, w[h].z, x.w * w[h].w)));
          for (int m = 1; m < LPR; m <<= 1) pdot += __shfl_xor_sync(0xffffffffu, pdot, m);
          tot[h] = pdot;
        }
        float r_out;
        row_epilogue<H, 0>(d, tot, hic, gic, poisson, rowok ? ys[r] : 0.f, rowok, rowok && chunk == 0, dh, r_out, sc);
        if (bwd) {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            g[h].x = fmaf(dh[h], x.x, g[h].x);
            g[h].y = fmaf(dh[h], x.y, g[h].y);
            g[h].z = fmaf(dh[h], x.z, g[h].z);
            g[h].w = fmaf(dh[h], x.w, g[h].w);
          }
        }
      }
      // ---- fixed-order reduction: row slots of a warp, then the CTA's warps, in fp64 --------------------
      double* my = red + (size_t)wid * NV;
      // (a warp's lanes hold at most a few dozen rows: their sum stays in fp32, as in the streaming sweeps,
      // which add up to 256 rows in fp32 before going to fp64)
      if (bwd) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float v[4] = {g[h].x, g[h].y, g[h].z, g[h].w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            for (int m = LPR; m < 32; m <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], m);
          if (slot == 0 && chunk_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) my[h * ctot + 4 * chunk + q] = (double)v[q];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < S_GHEAD + H; ++q) {
        float v = sc[q];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) my[H * ctot + q] = (double)v;
      }
      if (lane == 0)
        for (int q = S_GHEAD + H; q < kNumScalars; ++q) my[H * ctot + q] = 0.0;
      __syncthreads();
      const int e_lo = bwd ? 0 : H * ctot;
      for (int e = e_lo + tid; e < NV; e += kSmallThreads) {
        double s = 0.0;
        for (int wv = 0; wv < NW; ++wv) s += red[(size_t)wv * NV + e];
        if (bwd) out[e] = s;
        else out[a.NE + (e - H * ctot)] = s;
      }
      __syncthreads();
    }
    // ---- exchange = barrier: every CTA publishes "my partial row of iteration `it` is written" in a flag and
    // waits for everybody else's; the rows alternate between two parities (a CTA can be at most one iteration
    // ahead of the slowest one, because it needs that CTA's flag of the previous iteration to get there)
    // (measured at config 1, 64 CTAs: 12.0 us per iteration with the flags, 9.7 us with the cooperative-groups
    // grid barrier -- 64 polling threads per CTA cost more than one counter; the flags stay as an option)
    if (a.flags != nullptr) {
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release_gpu(a.flags + blockIdx.x, (unsigned long long)a.epoch + (unsigned long long)(k + 1));
      }
      for (unsigned c = tid; c < gridDim.x; c += kSmallThreads)
        while (ld_acquire_gpu(a.flags + c) < (unsigned long long)a.epoch + (unsigned long long)(k + 1)) {
        }
      __syncthreads();
    } else {
      cg::this_grid().sync();
    }

    // ---- every CTA: totals in CTA order, then bookkeeping + penalty + Adam on its own copy of the state ---
    {
      const double* base = a.part + (size_t)(k & 1) * gridDim.x * prow;
      const int n_fin = a.NE + (has_dev ? kNumScalars : 0);
      for (int e = tid; e < n_fin; e += kSmallThreads) {
        double s = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) s += __ldcg(base + (size_t)c * prow + e);
        fin[e] = s;
      }
      __syncthreads();
    }
    const double* sc = fin + (size_t)H * ctot;
    const double* fdev = fin + a.NE;
    // (the statements below are update_body of update.cuh on the shared-memory copy; CTA 0 records)
    if (cta0 && tid == 0 && it >= 1 && it - 1 < u.max_iters) {
      u.metric_train[it - 1] = score_from_sums(u.metric, sc, u.n_train, u.sy_train, u.syy_train);
      if (has_dev) {
        u.cost_dev[it - 1] = data_loss_from_sums(u.distr, fdev, u.n_dev);
        u.metric_dev[it - 1] = score_from_sums(u.metric, fdev, u.n_dev, u.sy_dev, u.syy_dev);
      }
    }
    double l1 = 0.0, l2sq = 0.0;
    if (u.penalise) {
      for (int j = tid; j < P; j += kSmallThreads) {
        const double t = (double)theta[j];
        l1 += fabs(t);
        l2sq += t * t;
      }
      l1 = block_sum(l1, upd_sm);
      l2sq = block_sum(l2sq, upd_sm);
    }
    const float l2 = (float)sqrt(l2sq);
    const bool use_l1 = u.penalise && u.alpha != 0.0f;
    const bool use_l2 = u.penalise && u.alpha != 1.0f;
    if (tid == 0) {
      if (cta0) {
        double cost = data_loss_from_sums(u.distr, sc, u.n_train);
        if (u.penalise) {
          const float pen = u.beta * ((1.0f - u.alpha) * (use_l2 ? l2 : 0.f) + u.alpha * (use_l1 ? (float)l1 : 0.f));
          cost += (double)pen;
        }
        u.cost_train[it] = cost;
      }
    }
    const double step = __ldg(u.step_sizes + it);
    const float lr = (float)step;
    const float b1 = 0.9f, b2 = 0.999f, c1 = (float)(1.0 - 0.9), c2 = (float)(1.0 - 0.999);
    const float eps = 1e-8f;
    const float bc1 = __ldg(u.bc_f + 2 * it), bc2 = __ldg(u.bc_f + 2 * it + 1);
    const float wl1 = u.beta * u.alpha, wl2 = u.beta * (1.0f - u.alpha);
    for (int j = tid; j < P; j += kSmallThreads) {
      const float t = theta[j];
      float gj = (float)fin[entry_of[j]];
      if (use_l1) gj += wl1 * ((t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f));
      if (use_l2) gj += wl2 * (t / l2);
      const float mj = c1 * gj + b1 * am[j];
      const float vj = c2 * (gj * gj) + b2 * av[j];
      const float mhat = mj / bc1;
      const float vhat = vj / bc2;
      const float tn = t - lr * mhat / (sqrtf(vhat) + eps);
      am[j] = mj;
      av[j] = vj;
      theta[j] = tn;
      if (cta0) u.ring_b[(size_t)it * P + j] = tn;
    }
    if (tid >= 64 && tid - 64 <= F) {                    // intercepts: float64 leaves, fp32 gradients
      const int q = tid - 64;
      const double gq = (q == F) ? (double)(float)sc[S_GGLOBAL] : (double)(float)sc[S_GHEAD + head_of[q]];
      const double mk = (1.0 - 0.9) * gq + 0.9 * mi[q];
      const double vk = (1.0 - 0.999) * (gq * gq) + 0.999 * vi[q];
      const double mhat = mk / __ldg(u.bc_d + 2 * it);
      const double vhat = vk / __ldg(u.bc_d + 2 * it + 1);
      const double xn = ic[q] - step * mhat / (sqrt(vhat) + 1e-8);
      mi[q] = mk;
      vi[q] = vk;
      ic[q] = xn;
      if (cta0) u.ring_ic[(size_t)it * (F + 1) + q] = xn;
    }
    __syncthreads();
    publish();
  }

  // ---- CTA 0 hands the state back to the fit's device buffers (what the graph path continues from) ------
  if (cta0) {
    for (int j = tid; j < P; j += kSmallThreads) {
      u.theta[j] = theta[j];
      u.m[j] = am[j];
      u.v[j] = av[j];
    }
    if (tid <= F) {
      u.ic[tid] = ic[tid];
      u.mi[tid] = mi[tid];
      u.vi[tid] = vi[tid];
    }
    for (int i = tid; i < H * ctot; i += kSmallThreads) u.W[i] = Wsm[i];
    if (tid < H) u.head_icpt[tid] = icpt[tid];
    if (tid == 0) {
      u.glob_icpt[0] = icpt[H];
      *u.it = it0 + a.n_iters;
    }
  }
}

}  // namespace rfb
