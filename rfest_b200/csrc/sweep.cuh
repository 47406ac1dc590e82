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
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// contiguous global -> shared copy by the copy engine, completion signalled on an mbarrier;
// streamed data is marked evict-first in L2 (it is read exactly once per sweep)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
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

constexpr int kSweepMaxThreads = 128;

// resident CTAs per SM the compiler must leave registers for (staged variant: more warps to
// overlap the per-tile arithmetic; the copy ring, not the warp count, hides HBM latency)
__host__ __device__ constexpr int sweep_min_blocks(int H, int NCH, bool staged) {
  return !staged ? 1 : (H * NCH <= 3 ? 3 : (H * NCH <= 6 ? 2 : 1));
}

// bytes of dynamic shared memory the staged variant needs
__host__ __device__ inline size_t sweep_smem_bytes(int warps, int stages, int RB, int ctot) {
  return (size_t)warps * stages * ((size_t)RB * ctot * sizeof(float) + sizeof(uint64_t));
}

// STAGED = false: rows go HBM -> registers with streaming 16-byte loads.
// STAGED = true : every warp runs its own `n_stages`-deep ring of bulk-async copies (one
//   contiguous RB x ld chunk per design block and tile), waits on the tile's mbarrier, pulls the
//   rows into registers with conflict-free LDS.128 and immediately re-arms the buffer with the
//   tile `n_stages` ahead -- the copy engine keeps the HBM pipe full while the warp computes.
template <int H, int NCH, int RB, bool BWD, bool STAGED>
__global__ void __launch_bounds__(kSweepMaxThreads, sweep_min_blocks(H, NCH, STAGED))
sweep_kernel(const __grid_constant__ SweepArgs a) {
  static_assert(RB >= 1 && RB <= 32 && (RB & (RB - 1)) == 0, "RB must be a power of two");
  extern __shared__ __align__(128) unsigned char sweep_smem[];
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpc + (threadIdx.x >> 5);
  const long long GW = (long long)gridDim.x * wpc;
  const int ctot = a.n_chunks * 4;
  const int NE = H * ctot + kNumScalars;

  // ---- column ownership -------------------------------------------------------------
  const float* cptr[NCH];
  long long cld[NCH];
  int soff[NCH];   // staged: float offset of this lane's chunk inside a stage buffer
  bool cvalid[NCH];
  float4 w[H][NCH];
  float4 g[H][NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    cvalid[i] = c < a.n_chunks;
    cptr[i] = nullptr;
    cld[i] = 0;
    soff[i] = 0;
    if (cvalid[i]) {
      int s = 0;
      while (s + 1 < a.n_slots && c >= a.slot_chunk0[s + 1]) ++s;
      cvalid[i] = a.slot_ptr[s] != nullptr;   // slot without a block in this data set
      cptr[i] = a.slot_ptr[s] + 4 * (c - a.slot_chunk0[s]);
      cld[i] = a.slot_ld[s];
      soff[i] = RB * 4 * a.slot_chunk0[s] + 4 * (c - a.slot_chunk0[s]);
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

  // staged: this warp's ring of stage buffers and their mbarriers
  const int NS = a.n_stages;
  const int stage_floats = RB * ctot;
  float* wbuf = nullptr;
  uint64_t* bars = nullptr;
  uint64_t policy = 0;
  auto issue = [&](long long tile, int buf) {   // lane 0: arm the barrier, start the copies
    const long long row0 = tile * RB;
    const int rows = (int)((a.n_rows - row0 < RB) ? (a.n_rows - row0) : RB);
    uint32_t bytes = 0;
    for (int s = 0; s < a.n_slots; ++s)
      if (a.slot_ptr[s] != nullptr) bytes += (uint32_t)(rows * a.slot_ld[s] * sizeof(float));
    mbar_arrive_expect_tx(bars + buf, bytes);
    float* dst = wbuf + (size_t)buf * stage_floats;
    for (int s = 0; s < a.n_slots; ++s)
      if (a.slot_ptr[s] != nullptr)
        bulk_g2s(dst + RB * 4 * a.slot_chunk0[s], a.slot_ptr[s] + row0 * a.slot_ld[s],
                 (uint32_t)(rows * a.slot_ld[s] * sizeof(float)), bars + buf, policy);
  };
  if constexpr (STAGED) {
    const int wic = threadIdx.x >> 5;
    wbuf = reinterpret_cast<float*>(sweep_smem) + (size_t)wic * NS * stage_floats;
    bars = reinterpret_cast<uint64_t*>(sweep_smem + (size_t)wpc * NS * stage_floats * sizeof(float)) + wic * NS;
    policy = l2_evict_first_policy();
    // rows past the end of the data keep whatever the buffer held: make that finite
    for (int i = lane; i < NS * stage_floats; i += 32) wbuf[i] = 0.f;
    if (lane == 0)
      for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < NS; ++s)
        if (gw + (long long)s * GW < ntiles) issue(gw + (long long)s * GW, s);
  }

  long long kt = 0;   // tiles this warp has consumed
  for (long long tile = gw; tile < ntiles; tile += GW, ++kt) {
    const long long row0 = tile * RB;
    float4 x[RB][NCH];
    const int buf = STAGED ? (int)(kt % NS) : 0;
    if constexpr (STAGED) {
      mbar_wait(bars + buf, (uint32_t)((kt / NS) & 1));
      const float* sb = wbuf + (size_t)buf * stage_floats;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          x[r][i] = cvalid[i] ? *reinterpret_cast<const float4*>(sb + soff[i] + r * (int)cld[i])
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const long long row = row0 + r;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          x[r][i] = (cvalid[i] && row < a.n_rows) ? ldg_stream_f4(cptr[i] + row * cld[i])
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
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
    if constexpr (STAGED) {
      // every lane's rows are in registers (the dot products above consumed them): hand the
      // buffer back to the copy engine for the tile NS ahead
      const long long nt = tile + (long long)NS * GW;
      __syncwarp();
      if (nt < ntiles && lane == 0) issue(nt, buf);
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
