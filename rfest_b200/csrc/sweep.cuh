// K2 / K4: the streaming sweep over the projected design XS.
//
// One read of XS per iteration computes everything `value_and_grad(cost)` and the
// monitoring forward passes of the reference need (rfest/GLM/_glm.py:539-614,692-696,
// 726-734; rfest/loss.py:4-16):
//   a_h = XS . W_h + c_h        (one dot product per head; 'none' filters are merged)
//   u   = sum_h phi_h(a_h) + c_global,   r = phi_out(u)
//   loss sums, metric sums, delta = dL/du, gradient  G_h += delta * phi_h'(a_h) * XS_row
//
// Thread mapping (HBM-bound; no tensor cores -- 1 flop/byte):
//   a warp owns whole rows.  Lane l owns the float4 column chunks l, l+32, ... (NCH of
//   them), keeps the head weights and the fp32 gradient accumulators of those columns in
//   registers, and processes RB rows per tile:
//     1. RB x NCH independent 16-byte streaming loads (ld.global.nc.L1::no_allocate),
//        512 contiguous bytes per warp instruction;
//     2. per-row partial dot products, then a transposed butterfly reduction so that lane
//        (l mod RB) ends up with the full sum of row (l mod RB): RB-1+log2(32/RB) shuffles
//        for RB rows instead of 5 per row;
//     3. nonlinearity / loss / delta once per row, lane = row;
//     4. delta broadcast by shuffle; G += delta * x from the SAME registers (XS is never
//        re-read, neither from HBM nor from shared memory).
//   Sums that span many rows are accumulated in fp32 for at most `flush_tiles` tiles and
//   then added to per-warp fp64 accumulators (L2-resident scratch); the kernel ends with a
//   fixed-order fp64 reduction over the CTA's warps.  No atomics: bit-reproducible.
#pragma once

#include "common.cuh"

namespace rfb {

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p));
  return v;
}

// phi(x) and phi'(x).  GLM.fnl, rfest/GLM/_glm.py:103-170.  Softplus is the reference's
// naive log(1 + exp(x)) + 1e-7 in fp32 (NOT log1p): for x < -16.6 it is exactly 1e-7.
__device__ __forceinline__ void nl_eval(int kind, float x, float& f, float& d) {
  switch (kind) {
    case RFB_NL_SOFTPLUS: {
      const float e = expf(x);
      const float s = 1.0f + e;
      f = logf(s) + 1e-7f;
      d = (e == CUDART_INF_F) ? 1.0f : e / s;
      break;
    }
    case RFB_NL_EXPONENTIAL: {
      const float e = expf(x);
      f = e;
      d = e;
      break;
    }
    case RFB_NL_SIGMOID: {
      const float e = expf(-x);
      const float s = 1.0f + e;
      f = 1.0f / s;
      d = (e == CUDART_INF_F) ? 0.0f : e / (s * s);
      break;
    }
    case RFB_NL_TANH: {
      const float t = tanhf(x);
      f = t;
      d = 1.0f - t * t;
      break;
    }
    case RFB_NL_RELU:
      f = x > 0.0f ? x : 1e-7f;
      d = x > 0.0f ? 1.0f : 0.0f;
      break;
    case RFB_NL_LEAKY_RELU:
      f = x > 0.0f ? x : x * 0.01f;
      d = x > 0.0f ? 1.0f : 0.01f;
      break;
    case kNlAbsent:   // a filter without data in this data set contributes nothing
      f = 0.0f;
      d = 0.0f;
      break;
    default:
      f = x;
      d = 1.0f;
      break;
  }
}

// p[j] holds this lane's partial sum for row j.  Returns, in every lane, the warp-wide sum
// of row (lane & (RB - 1)).
template <int RB>
__device__ __forceinline__ float transpose_reduce(float (&p)[RB], int lane) {
#pragma unroll
  for (int s = RB / 2; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float keep = up ? p[j + s] : p[j];
      const float send = up ? p[j] : p[j + s];
      p[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  float v = p[0];
#pragma unroll
  for (int m = RB; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

constexpr int kSweepThreads = 128;

template <int H, int NCH, int RB, bool BWD>
__global__ void __launch_bounds__(kSweepThreads) sweep_kernel(const __grid_constant__ SweepArgs a) {
  static_assert(RB >= 1 && RB <= 32 && (RB & (RB - 1)) == 0, "RB must be a power of two");
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpc + (threadIdx.x >> 5);
  const long long GW = (long long)gridDim.x * wpc;
  const int ctot = a.n_chunks * 4;
  const int NE = H * ctot + kNumScalars;

  // ---- column ownership -------------------------------------------------------------
  const float* cptr[NCH];
  long long cld[NCH];
  bool cvalid[NCH];
  float4 w[H][NCH];
  float4 g[H][NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    cvalid[i] = c < a.n_chunks;
    cptr[i] = nullptr;
    cld[i] = 0;
    if (cvalid[i]) {
      int s = 0;
      while (s + 1 < a.n_slots && c >= a.slot_chunk0[s + 1]) ++s;
      cvalid[i] = a.slot_ptr[s] != nullptr;   // slot without a block in this data set
      cptr[i] = a.slot_ptr[s] + 4 * (c - a.slot_chunk0[s]);
      cld[i] = a.slot_ld[s];
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      w[h][i] = cvalid[i] ? *reinterpret_cast<const float4*>(a.W + (size_t)h * ctot + 4 * c)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      g[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float hic[H];
#pragma unroll
  for (int h = 0; h < H; ++h) hic[h] = a.head_icpt[h];
  const float gic = a.glob_icpt[0];
  const bool poisson = a.distr == RFB_DISTR_POISSON;

  float sc[kNumScalars];
#pragma unroll
  for (int k = 0; k < kNumScalars; ++k) sc[k] = 0.f;

  double* my_acc = a.warp_acc + (size_t)gw * NE;
  bool first_flush = true;
  int since_flush = 0;

  auto flush = [&]() {
    if (BWD) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          if (cvalid[i]) {
            double* p = my_acc + (size_t)h * ctot + 4 * (lane + 32 * i);
            double2* p2 = reinterpret_cast<double2*>(p);
            double2 lo = make_double2((double)g[h][i].x, (double)g[h][i].y);
            double2 hi = make_double2((double)g[h][i].z, (double)g[h][i].w);
            if (!first_flush) {
              const double2 olo = p2[0], ohi = p2[1];
              lo.x += olo.x; lo.y += olo.y; hi.x += ohi.x; hi.y += ohi.y;
            }
            p2[0] = lo;
            p2[1] = hi;
          }
          g[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kNumScalars; ++k) {
      if (k < S_GHEAD + H) {
        double v = warp_sum_f64((double)sc[k]);
        if (lane == 0) {
          double* p = my_acc + (size_t)H * ctot + k;
          *p = first_flush ? v : (*p + v);
        }
      } else if (first_flush && lane == 0) {
        my_acc[(size_t)H * ctot + k] = 0.0;
      }
      sc[k] = 0.f;
    }
    first_flush = false;
    since_flush = 0;
  };

  // ---- stream the rows ------------------------------------------------------------------
  const long long ntiles = (a.n_rows + RB - 1) / RB;
  const int myr = lane & (RB - 1);
  for (long long tile = gw; tile < ntiles; tile += GW) {
    const long long row0 = tile * RB;
    float4 x[RB][NCH];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const long long row = row0 + r;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        x[r][i] = (cvalid[i] && row < a.n_rows) ? ldg_stream_f4(cptr[i] + row * cld[i])
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const long long myrow = row0 + myr;
    const bool rowok = myrow < a.n_rows;
    const float yv = (rowok && a.y != nullptr) ? __ldg(a.y + myrow) : 0.f;

    // forward: per-head dot products, transposed reduction
    float tot[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float p[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          acc = fmaf(x[r][i].x, w[h][i].x, acc);
          acc = fmaf(x[r][i].y, w[h][i].y, acc);
          acc = fmaf(x[r][i].z, w[h][i].z, acc);
          acc = fmaf(x[r][i].w, w[h][i].w, acc);
        }
        p[r] = acc;
      }
      tot[h] = transpose_reduce<RB>(p, lane);
    }

    // nonlinearities, loss, delta: lane = row
    float dh[H];
    float u = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float f;
      nl_eval(a.head_nl[h], tot[h] + hic[h], f, dh[h]);
      u += f;
    }
    u += gic;
    float r, dr;
    nl_eval(a.out_nl, u, r, dr);

    float la, lb, dldr;
    const float err = yv - r;
    if (poisson) {                                   // rfest/loss.py:4-11
      const bool fin = r != CUDART_INF_F;
      const float r1 = fin ? r : 0.f;
      const bool live = fin && (r1 > 1e-20f);
      const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
      la = -logf(r2) * yv;
      lb = r2;
      dldr = live ? (1.0f - yv / r2) : 0.f;
    } else {                                         // rfest/loss.py:14-16
      la = err * err;
      lb = 0.f;
      dldr = -err * a.two_over_n;
    }
    float delta = (dldr == 0.f) ? 0.f : dldr * dr;
    const bool count = rowok && (lane < RB);
    if (!rowok) delta = 0.f;
    if (count) {
      sc[S_LOSS_A] += la;
      sc[S_LOSS_B] += lb;
      sc[S_R] += r;
      sc[S_RR] += r * r;
      sc[S_YR] += yv * r;
      sc[S_SQERR] += err * err;
      sc[S_GGLOBAL] += delta;
      sc[S_BAD] += (r != r) ? 1.f : 0.f;
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      dh[h] = (delta == 0.f) ? 0.f : delta * dh[h];
      if (count) sc[S_GHEAD + h] += dh[h];
    }
    if (a.r_out != nullptr && count) a.r_out[myrow] = r;

    // backward: G_h += delta_h(row) * x(row)
    if (BWD) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float d = __shfl_sync(0xffffffffu, dh[h], r);
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            g[h][i].x = fmaf(d, x[r][i].x, g[h][i].x);
            g[h][i].y = fmaf(d, x[r][i].y, g[h][i].y);
            g[h][i].z = fmaf(d, x[r][i].z, g[h][i].z);
            g[h][i].w = fmaf(d, x[r][i].w, g[h][i].w);
          }
        }
      }
    }
    if (++since_flush >= a.flush_tiles) flush();
  }
  flush();  // every warp writes its row, even with no tiles

  // ---- fixed-order fp64 reduction over the CTA's warps ---------------------------------------
  __syncthreads();
  const double* cta_rows = a.warp_acc + (size_t)blockIdx.x * wpc * NE;
  double* out = a.cta_part + (size_t)blockIdx.x * NE;
  // (a forward-only sweep leaves the gradient part of its rows untouched: nobody reads it)
  for (int e = (BWD ? 0 : H * ctot) + threadIdx.x; e < NE; e += blockDim.x) {
    double s = 0.0;
    for (int wv = 0; wv < wpc; ++wv) s += cta_rows[(size_t)wv * NE + e];
    out[e] = s;
  }
}

// out[e] = sum_p part[p][e]: thread x = entry, thread y = strided slice of the partials.
constexpr int kReduceEx = 32, kReduceEy = 8;
struct ReduceSeg {
  const double* part;   // first entry of partial row 0
  int n_parts;
  long long stride;     // doubles between consecutive partial rows
  int n_entries;
  double* out;
};
__global__ void __launch_bounds__(kReduceEx * kReduceEy)
reduce_partials_kernel(ReduceSeg s0, ReduceSeg s1, int blocks0) {
  const ReduceSeg s = (int)blockIdx.x < blocks0 ? s0 : s1;
  const int b = (int)blockIdx.x < blocks0 ? blockIdx.x : blockIdx.x - blocks0;
  __shared__ double sm[kReduceEy][kReduceEx];
  const int e = b * kReduceEx + threadIdx.x;
  double acc = 0.0;
  if (e < s.n_entries) {
    for (int p = threadIdx.y; p < s.n_parts; p += kReduceEy) acc += s.part[(size_t)p * s.stride + e];
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && e < s.n_entries) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kReduceEy; ++k) t += sm[k][threadIdx.x];
    s.out[e] = t;
  }
}

}  // namespace rfb
