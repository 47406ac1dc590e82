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

// ---- bulk-async (TMA 1-D) staging helpers: cp.async.bulk + mbarrier, sm_90+ PTX ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// x / y to <= 2 ulp (MUFU.RCP based) for |y| < 2^126 -- one extra rounding next to the IEEE
// quotient, far inside the parity tolerance, and ~10 instructions shorter on the per-row chain.
__device__ __forceinline__ float fast_div(float x, float y) { return __fdividef(x, y); }

// phi(x) and phi'(x).  GLM.fnl, rfest/GLM/_glm.py:103-170.  Softplus is the reference's
// naive log(1 + exp(x)) + 1e-7 in fp32 (NOT log1p): for x < -16.6 it is exactly 1e-7.
__device__ __forceinline__ void nl_eval(int kind, float x, float& f, float& d) {
  switch (kind) {
    case RFB_NL_SOFTPLUS: {
      const float e = expf(x);
      const float s = 1.0f + e;
      f = logf(s) + 1e-7f;
      d = (e > 1e30f) ? 1.0f : fast_div(e, s);
      break;
    }
    case RFB_NL_SOFTMAX:        // exp(x) here; the epilogues divide by the global sum (softmax_fix)
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

constexpr int kSweepThreads = 128;   // 4 warps per CTA

// resident CTAs per SM the compiler must leave registers for
__host__ __device__ constexpr int sweep_min_blocks(int H, int NCH, bool staged) {
  return !staged ? 1 : (H * NCH <= 3 ? 2 : 1);   // 2 x 4 warps per SM measured best: up to 255 registers
}

// bytes of dynamic shared memory of the staged variant: per warp and stage the RB-row tile of
// every design block, RB responses (padded to 32 B) and one mbarrier
__host__ __device__ inline size_t sweep_stage_bytes(int RB, int ctot) {
  return (size_t)RB * ctot * sizeof(float) + 32;
}
__host__ __device__ inline size_t sweep_smem_bytes(int warps, int stages, int RB, int ctot) {
  return (size_t)warps * stages * (sweep_stage_bytes(RB, ctot) + sizeof(unsigned long long));
}

// Output softmax over all samples (GLM.fnl 'softmax', _glm.py:134-140): r_t = exp(u_t) / Z with
// Z = sum_s exp(u_s) over every sample of the data set (all ranks).  Z is produced by a first
// forward sweep with the exponential in its place; a.softmax = {1 / Z, D} with
// D = sum_s (dL/dr_s) r_s, so that  dL/du_t = r_t (dL/dr_t - D)  (the softmax Jacobian applied to the
// loss cotangent).  sum_t dL/du_t vanishes identically (the output does not depend on a shift of u), so
// the gradients of the global intercept and of the intercepts of 'none' filters are reported as exactly 0
// rather than as rounding noise, which Adam would normalise into a random walk of step_size per iteration.
__device__ __forceinline__ void softmax_scale(const SweepArgs& a, int out_kind, float& r) {
  if (out_kind == RFB_NL_SOFTMAX) r *= __ldg(a.softmax);
}
__device__ __forceinline__ float softmax_delta(const SweepArgs& a, int out_kind, float delta, float r, float dldr,
                                               bool rowok) {
  if (out_kind != RFB_NL_SOFTMAX) return delta;
  return rowok ? r * (dldr - __ldg(a.softmax + 1)) : 0.f;
}

// 'softmax' as a FILTER nonlinearity (GLM.fnl via forwardpass, _glm.py:134-140,561-565): head h's output is
// o_t = exp(a_t) / Z_h with Z_h = sum_s exp(a_s) over every sample of the data set (all ranks), so
//   dL/da_s = o_s (g_s - D_h),  g_s = dL/do_s = delta_s,  D_h = sum_t g_t o_t
// and the coefficient gradient is  sum_s delta_s o_s x_s  -  D_h * (sum_s exp(a_s) x_s) / Z_h.
// A preparatory sweep (a.prep: delta == 1, heads unscaled) yields Z_h in the head-intercept slot and
// E_h = sum exp(a) x in the gradient rows; the ordinary sweep then runs with phi = phi' = exp(a) / Z_h, which
// puts D_h into the head-intercept slot, and smx_fix_kernel (softmax.cuh) subtracts D_h E_h / Z_h.
__device__ __forceinline__ void head_softmax_fix(const SweepArgs& a, int kind, int h, float& f, float& d) {
  if (kind == RFB_NL_SOFTMAX) {
    if (!a.prep) f *= __ldg(a.head_scale + h);
    d = f;
  }
}

// ---- per-row epilogue shared by both variants: nonlinearities, loss terms, delta -------------------
// `tot[h]` = this lane's row's dot products.  Fills r, delta-scaled head derivatives dh[] and adds
// the row's terms to the fp32 scalar partials when `count`.
// SPEC fixes the nonlinearities / distribution at compile time for the common models (the switch
// and its branch resolution leave the per-row dependency chain):
//   0 run-time kinds; 1 heads none, softplus output, Poisson (LNP); 2 heads none, exponential
//   output, Poisson; 3 heads none, identity output, Gaussian (LG).
constexpr int kNumSpecs = 4;
template <int H, int SPEC>
__device__ __forceinline__ void row_epilogue(const SweepArgs& a, const float (&tot)[H], const float (&hic)[H],
                                             float gic, bool poisson_rt, float yv, bool rowok, bool count,
                                             float (&dh)[H], float& r_out, float (&sc)[kNumScalars]) {
  const bool poisson = SPEC == 0 ? poisson_rt : (SPEC != 3);
  const int out_kind = SPEC == 0 ? a.out_nl
                                 : (SPEC == 1 ? RFB_NL_SOFTPLUS : (SPEC == 2 ? RFB_NL_EXPONENTIAL : RFB_NL_NONE));
  float u = 0.f;
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float f;
    nl_eval(SPEC == 0 ? a.head_nl[h] : RFB_NL_NONE, tot[h] + hic[h], f, dh[h]);
    if (SPEC == 0) head_softmax_fix(a, a.head_nl[h], h, f, dh[h]);
    u += f;
  }
  u += gic;
  float r, dr;
  nl_eval(out_kind, u, r, dr);
  if (SPEC == 0) softmax_scale(a, out_kind, r);
  float la, lb, dldr;
  const float err = yv - r;
  if (poisson) {                                   // rfest/loss.py:4-11
    const bool fin = r != CUDART_INF_F;
    const float r1 = fin ? r : 0.f;
    const bool live = fin && (r1 > 1e-20f);
    const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
    la = -logf(r2) * yv;
    lb = r2;
    dldr = live ? (1.0f - fast_div(yv, r2)) : 0.f;
  } else {                                         // rfest/loss.py:14-16
    la = err * err;
    lb = 0.f;
    dldr = -err * a.two_over_n;
  }
  float delta = (dldr == 0.f || !rowok) ? 0.f : dldr * dr;
  if (SPEC == 0) delta = softmax_delta(a, out_kind, delta, r, dldr, rowok);
  if (SPEC == 0 && a.prep) delta = rowok ? 1.f : 0.f;
  if (count) {
    sc[S_LOSS_A] += la;
    sc[S_LOSS_B] += lb;
    sc[S_R] += r;
    sc[S_RR] += r * r;
    sc[S_YR] += yv * r;
    sc[S_SQERR] += err * err;
    if (SPEC != 0 || out_kind != RFB_NL_SOFTMAX) sc[S_GGLOBAL] += delta;
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    dh[h] = (delta == 0.f) ? 0.f : delta * dh[h];
    if (count && !(SPEC == 0 && out_kind == RFB_NL_SOFTMAX && a.head_nl[h] == RFB_NL_NONE)) sc[S_GHEAD + h] += dh[h];
  }
  r_out = r;
}

// ---- multi-head epilogue (staged kernel, H > 1) --------------------------------------------------
// After the transposed reduction lane L holds the dot product of (head, row) = (L / RB, L % RB)
// (VP = next power of two >= H * RB lanes; the rest idle).  Every lane evaluates ONE filter
// nonlinearity -- all H * RB of them in parallel instead of H in a row -- the per-row sums over
// heads are three shuffles, the output nonlinearity / loss / delta are replicated per row.
// Fills delta_hr (this lane's dL/da_{h,r}), and accumulates the row scalars in lanes < RB and
// the head-intercept gradient of this lane's head in `sc_head`.
template <int H, int RB, int VP>
__device__ __forceinline__ void multihead_epilogue(const SweepArgs& a, float tot, float hic_lane, int kind_lane,
                                                   bool lane_valid, float gic, bool poisson, float yv,
                                                   bool rowok, int lane, float& delta_hr, float& r_out,
                                                   float (&sc)[kNumScalars], float& sc_head) {
  float f, dphi;
  nl_eval(kind_lane, tot + hic_lane, f, dphi);
  head_softmax_fix(a, kind_lane, min(lane / RB, H - 1), f, dphi);
  float u = lane_valid ? f : 0.f;
#pragma unroll
  for (int m = RB; m < VP; m <<= 1) u += __shfl_xor_sync(0xffffffffu, u, m);
  u += gic;
  float r, dr;
  nl_eval(a.out_nl, u, r, dr);
  softmax_scale(a, a.out_nl, r);
  float la, lb, dldr;
  const float err = yv - r;
  if (poisson) {                                   // rfest/loss.py:4-11
    const bool fin = r != CUDART_INF_F;
    const float r1 = fin ? r : 0.f;
    const bool live = fin && (r1 > 1e-20f);
    const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
    la = -logf(r2) * yv;
    lb = r2;
    dldr = live ? (1.0f - fast_div(yv, r2)) : 0.f;
  } else {                                         // rfest/loss.py:14-16
    la = err * err;
    lb = 0.f;
    dldr = -err * a.two_over_n;
  }
  float delta = softmax_delta(a, a.out_nl, (dldr == 0.f || !rowok) ? 0.f : dldr * dr, r, dldr, rowok);
  if (a.prep) delta = rowok ? 1.f : 0.f;
  if (rowok && lane < RB) {
    sc[S_LOSS_A] += la;
    sc[S_LOSS_B] += lb;
    sc[S_R] += r;
    sc[S_RR] += r * r;
    sc[S_YR] += yv * r;
    sc[S_SQERR] += err * err;
    if (a.out_nl != RFB_NL_SOFTMAX) sc[S_GGLOBAL] += delta;
  }
  delta_hr = (delta == 0.f || !lane_valid) ? 0.f : delta * dphi;
  if (lane < H * RB && !(a.out_nl == RFB_NL_SOFTMAX && kind_lane == RFB_NL_NONE)) sc_head += delta_hr;
  r_out = r;
}

// fp32 partials -> this warp's fp64 accumulator row (L2 resident).  The row was zeroed at kernel
// start, so this is always read-modify-write.  Scalars live in lanes < RB only.
template <int H, int NCH, int RB, bool BWD>
__device__ __forceinline__ void flush_partials(double* my_acc, int ctot, int lane, const bool (&cvalid)[NCH],
                                               float4 (&g)[H][NCH], float (&sc)[kNumScalars]) {
  if (BWD) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        if (cvalid[i]) {
          double2* p2 = reinterpret_cast<double2*>(my_acc + (size_t)h * ctot + 4 * (lane + 32 * i));
          double2 lo = p2[0], hi = p2[1];
          lo.x += (double)g[h][i].x; lo.y += (double)g[h][i].y;
          hi.x += (double)g[h][i].z; hi.y += (double)g[h][i].w;
          p2[0] = lo;
          p2[1] = hi;
        }
        g[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < S_GHEAD + H; ++k) {
    double v = (double)sc[k];
#pragma unroll
    for (int m = RB / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) my_acc[(size_t)H * ctot + k] += v;
    sc[k] = 0.f;
  }
}

// head-intercept gradients of the multi-head epilogue: lane L carries head L / RB
template <int H, int RB>
__device__ __forceinline__ void flush_head_scalar(double* my_acc, int ctot, int lane, float& sc_head) {
  double v = (double)sc_head;
#pragma unroll
  for (int m = RB / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  if (lane < H * RB && (lane & (RB - 1)) == 0) my_acc[(size_t)H * ctot + S_GHEAD + lane / RB] += v;
  sc_head = 0.f;
}

template <int H, int NCH, bool BWD>
__device__ __forceinline__ void cta_reduce(const SweepArgs& a, int ctot, int NE) {
  // fixed-order fp64 reduction over the CTA's warps (a forward-only sweep leaves the gradient
  // part of its rows untouched: nobody reads it)
  __syncthreads();
  const int wpc = blockDim.x >> 5;
  const double* cta_rows = a.warp_acc + (size_t)blockIdx.x * wpc * NE;
  double* out = a.cta_part + (size_t)blockIdx.x * NE;
  for (int e = (BWD ? 0 : H * ctot) + threadIdx.x; e < NE; e += blockDim.x) {
    double s = 0.0;
    for (int wv = 0; wv < wpc; ++wv) s += cta_rows[(size_t)wv * NE + e];
    out[e] = s;
  }
}

// ================================================================================================
// Variant A -- direct: rows go HBM -> registers with streaming 16-byte loads.
// ================================================================================================
template <int H, int NCH, int RB, bool BWD>
__global__ void __launch_bounds__(kSweepThreads) sweep_direct_kernel(const __grid_constant__ SweepArgs a) {
  sweep_stamp(a, 0);
  static_assert(RB >= 1 && RB <= 32 && (RB & (RB - 1)) == 0, "RB must be a power of two");
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpc + (threadIdx.x >> 5);
  const long long GW = (long long)gridDim.x * wpc;
  const int ctot = a.n_chunks * 4;
  const int NE = H * ctot + kNumScalars;

  const float* cptr[NCH];
  long long cld[NCH];
  bool cvalid[NCH];
  float4 w[H][NCH], g[H][NCH];
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
  for (int e = lane; e < NE; e += 32) my_acc[e] = 0.0;
  __syncwarp();
  int since_flush = 0;

  const long long ntiles = (a.n_rows + RB - 1) / RB;
  const int myr = lane & (RB - 1);
  for (long long tile = gw; tile < ntiles; tile += GW) {
    const long long row0 = tile * RB;
    float4 x[RB][NCH];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const long long row = row0 + r;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        x[r][i] = (cvalid[i] && row < a.n_rows) ? ldg_stream_f4(cptr[i] + row * cld[i])
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long myrow = row0 + myr;
    const bool rowok = myrow < a.n_rows;
    const float yv = (rowok && a.y != nullptr) ? __ldg(a.y + myrow) : 0.f;

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
    float dh[H], r;
    const bool count = rowok && (lane < RB);
    row_epilogue<H, 0>(a, tot, hic, gic, poisson, yv, rowok, count, dh, r, sc);
    if (a.r_out != nullptr && count) a.r_out[myrow] = r;
    if (BWD) {
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float d = __shfl_sync(0xffffffffu, dh[h], rr);
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            g[h][i].x = fmaf(d, x[rr][i].x, g[h][i].x);
            g[h][i].y = fmaf(d, x[rr][i].y, g[h][i].y);
            g[h][i].z = fmaf(d, x[rr][i].z, g[h][i].z);
            g[h][i].w = fmaf(d, x[rr][i].w, g[h][i].w);
          }
        }
      }
    }
    if (++since_flush >= a.flush_tiles) {
      flush_partials<H, NCH, RB, BWD>(my_acc, ctot, lane, cvalid, g, sc);
      since_flush = 0;
    }
  }
  flush_partials<H, NCH, RB, BWD>(my_acc, ctot, lane, cvalid, g, sc);
  cta_reduce<H, NCH, BWD>(a, ctot, NE);
  sweep_stamp(a, 1);
}

// ================================================================================================
// Variant B -- staged (default): every warp runs its own `n_stages`-deep ring of bulk-async
// copies (cp.async.bulk: one contiguous RB x ld chunk per design block and tile, plus the RB
// responses), waits on the tile's mbarrier, pulls the rows into registers with conflict-free
// LDS.128 and immediately re-arms the buffer with the tile `n_stages` ahead.  The copy engine
// keeps the HBM pipe full while the warp computes; nothing in the loop touches global memory.
// ================================================================================================
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                           uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

template <int H, int NCH, int RB, bool BWD, int SPEC>
__global__ void __launch_bounds__(kSweepThreads, sweep_min_blocks(H, NCH, true))
sweep_staged_kernel(const __grid_constant__ SweepArgs a) {
  sweep_stamp(a, 0);
  static_assert(RB >= 1 && RB <= 8 && (RB & (RB - 1)) == 0, "RB must be 1, 2, 4 or 8");
  extern __shared__ __align__(128) unsigned char sweep_smem[];
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpc + wic;
  const long long GW = (long long)gridDim.x * wpc;
  const int ctot = a.n_chunks * 4;
  const int NE = H * ctot + kNumScalars;
  const int NS = a.n_stages;
  const uint32_t stage_bytes = (uint32_t)sweep_stage_bytes(RB, ctot);
  const uint32_t y_off = (uint32_t)(RB * ctot * sizeof(float));       // responses sit after the rows

  // ---- column ownership.  A lane's unused chunk slots alias its first chunk with zero weights,
  // so the inner loops carry no predicates (the loaded values are finite data, times zero).
  uint32_t laddr[NCH], lstride[NCH];
  bool cvalid[NCH];
  float4 w[H][NCH], g[H][NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    cvalid[i] = c < a.n_chunks;
    laddr[i] = 0;
    lstride[i] = 0;
    if (cvalid[i]) {
      int s = 0;
      while (s + 1 < a.n_slots && c >= a.slot_chunk0[s + 1]) ++s;
      cvalid[i] = a.slot_ptr[s] != nullptr;
      laddr[i] = (uint32_t)((RB * 4 * a.slot_chunk0[s] + 4 * (c - a.slot_chunk0[s])) * sizeof(float));
      lstride[i] = (uint32_t)(a.slot_ld[s] * sizeof(float));
    }
    if (!cvalid[i]) {
      laddr[i] = laddr[0];
      lstride[i] = lstride[0];
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
  for (int e = lane; e < NE; e += 32) my_acc[e] = 0.0;

  // ---- the copies: lane s moves design block s, lane n_slots moves the responses ------------------
  const char* my_src = nullptr;
  uint32_t my_rowbytes = 0, my_dst = 0, all_rowbytes = 0;
  for (int s = 0; s < a.n_slots; ++s) {
    if (a.slot_ptr[s] == nullptr) continue;
    all_rowbytes += (uint32_t)(a.slot_ld[s] * sizeof(float));
    if (lane == s) {
      my_src = reinterpret_cast<const char*>(a.slot_ptr[s]);
      my_rowbytes = (uint32_t)(a.slot_ld[s] * sizeof(float));
      my_dst = (uint32_t)(RB * 4 * a.slot_chunk0[s] * sizeof(float));
    }
  }
  constexpr bool kStageY = RB >= 4;   // a 32-byte bulk copy needs a 16-byte aligned source
  const bool has_y = a.y != nullptr;
  if (kStageY && has_y && lane == a.n_slots) {
    my_src = reinterpret_cast<const char*>(a.y);
    my_rowbytes = sizeof(float);
    my_dst = y_off;
  }
  const bool y_lane = kStageY && has_y && lane == a.n_slots;
  const uint32_t tile_y_bytes = (kStageY && has_y) ? 32u : 0u;   // RB <= 8 responses as 32 B (y is padded)

  const uint32_t wbase = smem_u32(sweep_smem) + (uint32_t)wic * NS * stage_bytes;
  const uint32_t bbase = smem_u32(sweep_smem) + (uint32_t)wpc * NS * stage_bytes + (uint32_t)wic * NS * 8u;
  const uint64_t policy = l2_evict_first_policy();
  const long long ntiles = (a.n_rows + RB - 1) / RB;

  auto issue = [&](long long tile, uint32_t stage_addr, uint32_t bar) {
    const long long row0 = tile * RB;
    const long long left = a.n_rows - row0;
    const uint32_t rows = left < RB ? (uint32_t)left : (uint32_t)RB;
    if (lane == 0) mbar_arrive_expect_tx_a(bar, rows * all_rowbytes + tile_y_bytes);
    if (my_rowbytes != 0) {
      const uint32_t bytes = y_lane ? 32u : rows * my_rowbytes;
      bulk_g2s_a(stage_addr + my_dst, my_src + row0 * my_rowbytes, bytes, bar, policy);
    }
  };

  // rows past the end of the data keep whatever the buffer held: make that finite
  for (uint32_t o = lane * 4u; o < (uint32_t)NS * stage_bytes; o += 128u)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(wbase + o), "r"(0u) : "memory");
  if (lane == 0)
    for (int s = 0; s < NS; ++s) mbar_init_a(bbase + 8u * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  for (int s = 0; s < NS; ++s)
    if (gw + (long long)s * GW < ntiles) issue(gw + (long long)s * GW, wbase + s * stage_bytes, bbase + 8u * s);

  // multi-head layout: lane L <-> (head, row) = (L / RBS, L % RBS) among the first VP lanes.  When
  // H * RB exceeds the warp, the tile is handled as SUB sub-tiles of RBS rows: one stage, one barrier
  // wait and one refill per tile, SUB independent epilogue chains that the scheduler interleaves.
  constexpr int RBS = (H * RB <= 32) ? RB : RB / 2;
  constexpr int SUB = RB / RBS;
  constexpr int V = H * RBS;
  constexpr int VP = V <= 1 ? 1 : V <= 2 ? 2 : V <= 4 ? 4 : V <= 8 ? 8 : V <= 16 ? 16 : 32;
  static_assert(V <= 32, "H * RBS must not exceed the warp size");
  static_assert(H > 1 || SUB == 1, "sub-tiles are a multi-head device");
  const int my_idx = lane & (VP - 1);
  const bool lane_valid = my_idx < V;
  const int my_h = lane_valid ? my_idx / RBS : 0;
  const float hic_lane = a.head_icpt[my_h];
  const int kind_lane = lane_valid ? a.head_nl[my_h] : RFB_NL_NONE;
  float sc_head = 0.f;
  bool use[H][NCH];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int i = 0; i < NCH; ++i) use[h][i] = (H == 1) || ((a.head_slotmask[h] >> i) & 1u);

  const int myr = (H == 1) ? (lane & (RB - 1)) : (my_idx % RBS);     // row inside the (sub-)tile
  int since_flush = 0;
  int buf = 0;
  uint32_t phase = 0;
  for (long long tile = gw; tile < ntiles; tile += GW) {
    const uint32_t stage_addr = wbase + (uint32_t)buf * stage_bytes;
    const uint32_t bar = bbase + 8u * (uint32_t)buf;
    mbar_wait_a(bar, phase);

    // Three heads: weights + gradient accumulators take 72 registers, and holding 8 rows x 3 chunks of the
    // tile from the forward to the backward pass on top of that does not fit.  REREAD: the rows are read
    // from the shared-memory stage twice (once per pass, cheap) and the stage goes back to the copy engine
    // after the backward pass -- which buys 8-row tiles (half as many per-tile epilogue chains per row;
    // with 4-row tiles this shape ran at 60 % of the HBM rate while two heads with 8-row tiles saturate it).
    constexpr bool REREAD = BWD && H == 3;   // (5 heads: measured slower with it, 6.9 vs 5.2 ms at config 4)
    float4 x[REREAD ? 1 : RB][NCH];
    if constexpr (!REREAD) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        uint32_t ad = stage_addr + laddr[i];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          x[r][i] = lds_f4(ad);
          ad += lstride[i];
        }
      }
    }
    float yv[SUB];
#pragma unroll
    for (int sb = 0; sb < SUB; ++sb) {
      yv[sb] = 0.f;
      if (has_y) {
        if (kStageY) yv[sb] = lds_f1(stage_addr + y_off + 4u * (uint32_t)(sb * RBS + myr));
        else if (tile * RB + sb * RBS + myr < a.n_rows) yv[sb] = __ldg(a.y + tile * RB + sb * RBS + myr);
      }
    }

    // forward: per (head, row) partial dot products (chunk slots a head has no columns in are
    // skipped warp-uniformly), then ONE transposed reduction over all H * RB values
    float pv[SUB][VP];
    if constexpr (!REREAD) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            if (use[h][i]) {
              acc = fmaf(x[r][i].x, w[h][i].x, acc);
              acc = fmaf(x[r][i].y, w[h][i].y, acc);
              acc = fmaf(x[r][i].z, w[h][i].z, acc);
              acc = fmaf(x[r][i].w, w[h][i].w, acc);
            }
          }
          pv[r / RBS][h * RBS + (r % RBS)] = acc;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[0][i] = lds_f4(stage_addr + laddr[i] + (uint32_t)r * lstride[i]);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            if (use[h][i]) {
              acc = fmaf(x[0][i].x, w[h][i].x, acc);
              acc = fmaf(x[0][i].y, w[h][i].y, acc);
              acc = fmaf(x[0][i].z, w[h][i].z, acc);
              acc = fmaf(x[0][i].w, w[h][i].w, acc);
            }
          }
          pv[r / RBS][h * RBS + (r % RBS)] = acc;
        }
      }
    }
    float tot_lane[SUB];
#pragma unroll
    for (int sb = 0; sb < SUB; ++sb) {
#pragma unroll
      for (int k = V; k < VP; ++k) pv[sb][k] = 0.f;
      tot_lane[sb] = transpose_reduce<VP>(pv[sb], lane);
    }
    // every lane's rows (and y) are in registers -- the dot products consumed them: hand the
    // buffer back to the copy engine for the tile NS ahead
    if constexpr (!REREAD) {
      const long long nt = tile + (long long)NS * GW;
      __syncwarp();
      if (nt < ntiles) issue(nt, stage_addr, bar);
    }

    const long long myrow = tile * RB + myr;
    const bool rowok = myrow < a.n_rows;
    float r;
    if constexpr (H == 1) {
      const bool count = rowok && (lane < RB);
      float tot[1] = {tot_lane[0]}, dh[1];
      row_epilogue<1, SPEC>(a, tot, hic, gic, poisson, rowok ? yv[0] : 0.f, rowok, count, dh, r, sc);
      if (a.r_out != nullptr && count) a.r_out[myrow] = r;
      if (BWD) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
          const float d = __shfl_sync(0xffffffffu, dh[0], rr);
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            g[0][i].x = fmaf(d, x[rr][i].x, g[0][i].x);
            g[0][i].y = fmaf(d, x[rr][i].y, g[0][i].y);
            g[0][i].z = fmaf(d, x[rr][i].z, g[0][i].z);
            g[0][i].w = fmaf(d, x[rr][i].w, g[0][i].w);
          }
        }
      }
    } else {
      float delta_hr[SUB];
#pragma unroll
      for (int sb = 0; sb < SUB; ++sb) {
        const long long srow = myrow + sb * RBS;
        const bool sok = srow < a.n_rows;
        multihead_epilogue<H, RBS, VP>(a, tot_lane[sb], hic_lane, kind_lane, lane_valid, gic, poisson,
                                       sok ? yv[sb] : 0.f, sok, lane, delta_hr[sb], r, sc, sc_head);
        if (a.r_out != nullptr && sok && lane < RBS) a.r_out[srow] = r;
      }
      if (BWD) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
          if constexpr (REREAD) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[0][i] = lds_f4(stage_addr + laddr[i] + (uint32_t)rr * lstride[i]);
          }
          const int xr = REREAD ? 0 : rr;
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float d = __shfl_sync(0xffffffffu, delta_hr[rr / RBS], h * RBS + (rr % RBS));
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
              if (use[h][i]) {
                g[h][i].x = fmaf(d, x[xr][i].x, g[h][i].x);
                g[h][i].y = fmaf(d, x[xr][i].y, g[h][i].y);
                g[h][i].z = fmaf(d, x[xr][i].z, g[h][i].z);
                g[h][i].w = fmaf(d, x[xr][i].w, g[h][i].w);
              }
            }
          }
        }
      }
    }
    if constexpr (REREAD) {       // both passes have read the stage: refill it for the tile NS ahead
      const long long nt = tile + (long long)NS * GW;
      __syncwarp();
      if (nt < ntiles) issue(nt, stage_addr, bar);
    }
    if (++buf == NS) {
      buf = 0;
      phase ^= 1u;
    }
    if (++since_flush >= a.flush_tiles) {
      flush_partials<H, NCH, RBS, BWD>(my_acc, ctot, lane, cvalid, g, sc);
      if constexpr (H > 1) flush_head_scalar<H, RBS>(my_acc, ctot, lane, sc_head);
      since_flush = 0;
    }
  }
  flush_partials<H, NCH, RBS, BWD>(my_acc, ctot, lane, cvalid, g, sc);
  if constexpr (H > 1) flush_head_scalar<H, RBS>(my_acc, ctot, lane, sc_head);
  cta_reduce<H, NCH, BWD>(a, ctot, NE);
  sweep_stamp(a, 1);
}

// ================================================================================================
// Variant C -- generic fallback: any number of columns, up to 16 heads.  Nothing is kept in
// registers across rows: the head weights are read through L1, and the gradient goes straight
// into the warp's fp64 accumulator row (L2 resident) once per 4-row tile.  About 3x the memory
// traffic of the register-resident kernels; used only for shapes those cannot take (e.g. a 3-D
// filter without a spline basis: 7500 columns).
// ================================================================================================
constexpr int kGenRB = 4;
constexpr int kGenHeads = kMaxHeads;

template <bool BWD>
__global__ void __launch_bounds__(kSweepThreads) sweep_generic_kernel(const __grid_constant__ SweepArgs a, int H) {
  sweep_stamp(a, 0);
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpc + (threadIdx.x >> 5);
  const long long GW = (long long)gridDim.x * wpc;
  const int ctot = a.n_chunks * 4;
  const int NE = H * ctot + kNumScalars;
  double* my_acc = a.warp_acc + (size_t)gw * NE;
  for (int e = lane; e < NE; e += 32) my_acc[e] = 0.0;
  __syncwarp();
  const bool poisson = a.distr == RFB_DISTR_POISSON;
  const float gic = a.glob_icpt[0];
  const float hic_lane = lane < H ? a.head_icpt[lane] : 0.f;
  const int kind_lane = lane < H ? a.head_nl[lane] : RFB_NL_NONE;
  double sc[S_GHEAD];          // row scalars, lane 0
  double sc_head = 0.0;        // head-intercept gradient, lane = head
#pragma unroll
  for (int k = 0; k < S_GHEAD; ++k) sc[k] = 0.0;

  auto chunk_ptr = [&](int c, long long row) -> const float* {
    int s = 0;
    while (s + 1 < a.n_slots && c >= a.slot_chunk0[s + 1]) ++s;
    const float* p = a.slot_ptr[s];
    return p ? p + row * a.slot_ld[s] + 4 * (c - a.slot_chunk0[s]) : nullptr;
  };

  const long long ntiles = (a.n_rows + kGenRB - 1) / kGenRB;
  for (long long tile = gw; tile < ntiles; tile += GW) {
    const long long row0 = tile * kGenRB;
    // pass 1: dot products of the tile's rows with every head
    float acc[kGenHeads][kGenRB];
#pragma unroll
    for (int h = 0; h < kGenHeads; ++h)
#pragma unroll
      for (int r = 0; r < kGenRB; ++r) acc[h][r] = 0.f;
    for (int c = lane; c < a.n_chunks; c += 32) {
      float4 x[kGenRB];
#pragma unroll
      for (int r = 0; r < kGenRB; ++r) {
        const float* p = (row0 + r < a.n_rows) ? chunk_ptr(c, row0 + r) : nullptr;
        x[r] = p ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int h = 0; h < kGenHeads; ++h) {
        if (h < H) {
          const float4 w = *reinterpret_cast<const float4*>(a.W + (size_t)h * ctot + 4 * c);
#pragma unroll
          for (int r = 0; r < kGenRB; ++r)
            acc[h][r] = fmaf(x[r].x, w.x, fmaf(x[r].y, w.y, fmaf(x[r].z, w.z, fmaf(x[r].w, w.w, acc[h][r]))));
        }
      }
    }
    __syncwarp();   // the chunk loop is lane-divergent; reconverge before the shuffles
    // per row: warp sums, nonlinearities (lane = head), loss, delta
    float delta_h[kGenRB];      // lane h: dL/da_{h, r}
#pragma unroll
    for (int r = 0; r < kGenRB; ++r) {
      float mine = 0.f;         // lane h keeps the full dot product of head h
#pragma unroll
      for (int h = 0; h < kGenHeads; ++h) {
        if (h < H) {
          float v = acc[h][r];
#pragma unroll
          for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
          if (lane == h) mine = v;
        }
      }
      float f, dphi;
      nl_eval(kind_lane, mine + hic_lane, f, dphi);
      head_softmax_fix(a, kind_lane, min(lane, H - 1), f, dphi);
      float u = lane < H ? f : 0.f;
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) u += __shfl_xor_sync(0xffffffffu, u, m);
      u += gic;
      float rr, dr;
      nl_eval(a.out_nl, u, rr, dr);
      softmax_scale(a, a.out_nl, rr);
      const long long row = row0 + r;
      const bool rowok = row < a.n_rows;
      const float yv = (rowok && a.y != nullptr) ? __ldg(a.y + row) : 0.f;
      float la, lb, dldr;
      const float err = yv - rr;
      if (poisson) {
        const bool fin = rr != CUDART_INF_F;
        const float r1 = fin ? rr : 0.f;
        const bool live = fin && (r1 > 1e-20f);
        const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
        la = -logf(r2) * yv;
        lb = r2;
        dldr = live ? (1.0f - fast_div(yv, r2)) : 0.f;
      } else {
        la = err * err;
        lb = 0.f;
        dldr = -err * a.two_over_n;
      }
      float delta = softmax_delta(a, a.out_nl, (dldr == 0.f || !rowok) ? 0.f : dldr * dr, rr, dldr, rowok);
      if (a.prep) delta = rowok ? 1.f : 0.f;
      if (rowok && lane == 0) {
        sc[S_LOSS_A] += la; sc[S_LOSS_B] += lb; sc[S_R] += rr; sc[S_RR] += rr * rr;
        sc[S_YR] += yv * rr; sc[S_SQERR] += err * err;
        if (a.out_nl != RFB_NL_SOFTMAX) sc[S_GGLOBAL] += delta;
        if (a.r_out != nullptr) a.r_out[row] = rr;
      }
      delta_h[r] = (delta == 0.f || lane >= H) ? 0.f : delta * dphi;
      if (!(a.out_nl == RFB_NL_SOFTMAX && kind_lane == RFB_NL_NONE)) sc_head += (double)delta_h[r];
    }
    // pass 2: gradient of the tile, added once per chunk into the fp64 accumulator row.  The delta
    // broadcasts happen before the (lane-divergent) chunk loop: no shuffles inside it.
    if (BWD) {
      float d[kGenHeads][kGenRB];
#pragma unroll
      for (int h = 0; h < kGenHeads; ++h)
#pragma unroll
        for (int r = 0; r < kGenRB; ++r) d[h][r] = __shfl_sync(0xffffffffu, delta_h[r], h);
      for (int c = lane; c < a.n_chunks; c += 32) {
        float4 x[kGenRB];
#pragma unroll
        for (int r = 0; r < kGenRB; ++r) {
          const float* p = (row0 + r < a.n_rows) ? chunk_ptr(c, row0 + r) : nullptr;   // L1 / L2 hit
          x[r] = p ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int h = 0; h < kGenHeads; ++h) {
          if (h < H) {
            float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < kGenRB; ++r) {
              gsum.x = fmaf(d[h][r], x[r].x, gsum.x); gsum.y = fmaf(d[h][r], x[r].y, gsum.y);
              gsum.z = fmaf(d[h][r], x[r].z, gsum.z); gsum.w = fmaf(d[h][r], x[r].w, gsum.w);
            }
            double2* p2 = reinterpret_cast<double2*>(my_acc + (size_t)h * ctot + 4 * c);
            double2 lo = p2[0], hi = p2[1];
            lo.x += (double)gsum.x; lo.y += (double)gsum.y; hi.x += (double)gsum.z; hi.y += (double)gsum.w;
            p2[0] = lo;
            p2[1] = hi;
          }
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0)
    for (int k = 0; k < S_GHEAD; ++k) my_acc[(size_t)H * ctot + k] = sc[k];
  if (lane < H) my_acc[(size_t)H * ctot + S_GHEAD + lane] = sc_head;

  __syncthreads();
  const double* cta_rows = a.warp_acc + (size_t)blockIdx.x * wpc * NE;
  double* out = a.cta_part + (size_t)blockIdx.x * NE;
  for (int e = (BWD ? 0 : H * ctot) + threadIdx.x; e < NE; e += blockDim.x) {
    double s = 0.0;
    for (int wv = 0; wv < wpc; ++wv) s += cta_rows[(size_t)wv * NE + e];
    out[e] = s;
  }
  sweep_stamp(a, 1);
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

// The same reduction fused with the multi-GPU exchange: instead of a local vector, the sums are
// stored straight into this rank's slot of every peer's exchange buffer (P2P stores over
// NVLink); the last CTA to finish publishes the sequence number to all peers.
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kReduceEx * kReduceEy)
reduce_exchange_kernel(ReduceSeg s0, ReduceSeg s1, int blocks0, long long off0, long long off1, XchgArgs x) {
  const bool first = (int)blockIdx.x < blocks0;
  const ReduceSeg s = first ? s0 : s1;
  const long long off = first ? off0 : off1;
  const int b = first ? blockIdx.x : blockIdx.x - blocks0;
  __shared__ double sm[kReduceEy][kReduceEx];
  const unsigned long long seq = *x.seq + 1ull;
  const int parity = (int)(seq & 1ull);
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
    const long long slot = ((long long)parity * x.world + x.rank) * x.cap + off + e;
    for (int q = 0; q < x.world; ++q)
      reinterpret_cast<double*>(x.peer[q] + kXchgDataOffset)[slot] = t;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int prev = atomicAdd(x.done, 1u);
    if (prev == gridDim.x - 1) {
      *x.done = 0u;
      __threadfence_system();
      for (int q = 0; q < x.world; ++q)
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(x.peer[q]) + parity * kMaxRanks + x.rank, seq);
    }
  }
}

// One-CTA form of the fused reduction + exchange (round 2 experiment, `xchg = 3`).  The multi-CTA kernel above needs
// two system-scope fences in sequence (every CTA before it counts itself done, the last CTA again before it
// publishes) plus an atomic round trip.  Here one warp owns 32 consecutive entries and adds the CTA partial rows in
// row order (lanes = entries: coalesced, no shared memory, no barrier inside the sum), stores the sums into every
// peer, and after ONE __syncthreads a single thread issues ONE system-scope fence and the flags.  Measured SLOWER:
// 43 us against 20 us at 2 GPUs (rest of the step 52 vs 32 us, `profiles/xchg_variants_r2.txt`) -- 16 warps walking
// 148-296 partial rows 8 at a time are ~37 dependent L2 round trips; the multi-CTA form spreads them over 11 CTAs.
constexpr int kReduce1Threads = 512;
__global__ void __launch_bounds__(kReduce1Threads)
reduce_exchange1_kernel(ReduceSeg s0, ReduceSeg s1, long long off0, long long off1, XchgArgs x) {
  const unsigned long long seq = *x.seq + 1ull;
  const int parity = (int)(seq & 1ull);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = kReduce1Threads / 32;
  const int chunks0 = (s0.n_entries + 31) / 32, chunks1 = (s1.n_entries + 31) / 32;
  for (int c = warp; c < chunks0 + chunks1; c += n_warps) {
    const bool first = c < chunks0;
    const ReduceSeg& s = first ? s0 : s1;
    const int e = (first ? c : c - chunks0) * 32 + lane;
    if (e >= s.n_entries) continue;
    const double* col = s.part + e;
    double t = 0.0;
    int p = 0;
    for (; p + 8 <= s.n_parts; p += 8) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = col[(size_t)(p + k) * s.stride];
#pragma unroll
      for (int k = 0; k < 8; ++k) t += v[k];
    }
    for (; p < s.n_parts; ++p) t += col[(size_t)p * s.stride];
    const long long slot = ((long long)parity * x.world + x.rank) * x.cap + (first ? off0 : off1) + e;
    for (int q = 0; q < x.world; ++q)
      reinterpret_cast<double*>(x.peer[q] + kXchgDataOffset)[slot] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int q = 0; q < x.world; ++q)
      st_release_sys_u64(reinterpret_cast<unsigned long long*>(x.peer[q]) + parity * kMaxRanks + x.rank, seq);
  }
}

}  // namespace rfb
