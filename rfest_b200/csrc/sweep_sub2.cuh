// K2s2: the subunit sweep with a CTA-shared stage pool and its gradient accumulators in TENSOR MEMORY.
//
// sweep_subunit_kernel (sweep_sub.cuh) is latency-bound at BASELINE config 4 (4 softplus subunits + history):
// no unit is saturated (issue 50 %, FMA 47 %, shared memory 60 %, DRAM 71 %) but its 255 registers leave
// only 8 warps per SM -- two per scheduler -- to cover a ~250-instruction epilogue chain per tile.  80 of
// those registers are the gradient accumulators (4 heads x 5 chunks x float4 per lane), which are touched
// once per tile.  Blackwell has 256 KB of tensor memory per SM that ordinary warps can read and write with
// tcgen05.ld / tcgen05.st (no MMA involved): the accumulators live THERE, 80 columns per warp, and pass
// through 16 registers at a time (one column chunk of the block for all heads) during the backward products.
// That brings the kernel under 160 registers, i.e. 12 consumer warps per SM.
//
// 12 warps with a private two-stage ring each would need 228 KB of shared memory, so the stages become a
// CTA-wide pool: one producer warp issues the bulk-async copies (cp.async.bulk, mbarrier complete_tx) for
// the CTA's tiles in order, stage = tile mod NS, consumer warp = tile mod 12; a consumer waits on the
// stage's `full` barrier, runs forward products / transposed reduction / epilogue / backward products
// exactly as sweep_subunit_kernel does, and arrives on the stage's `empty` barrier, which the producer
// waits on before refilling.  Arithmetic, accumulation cadence (fp32 for <= 64 tiles, then fp64) and the
// fixed-order reductions are unchanged: results are bit-reproducible run to run.
//
// Tensor-memory map: the CTA allocates 256 columns; warp w may only touch lanes [32 (w % 4), +32), so the
// three consumer warps of a lane quarter take columns [80 (w / 4), +80): column 16 j + 4 k + {0, 1, 2, 3} =
// gradient of chunk slot j, head k, floats {x, y, z, w} of this lane's float4 column chunk.
#pragma once

#include "project_tc.cuh"
#include "sweep_sub.cuh"

namespace rfb {

// consumer warps per CTA: registers are granted to a CTA in units of four warps, so 13 warps (12 + the
// producer) and 16 warps (15 + the producer) both run at <= 128 registers per thread
__host__ __device__ constexpr int sub2_threads(int ncw) { return (ncw + 1) * 32; }
__host__ __device__ constexpr int sub2_tmem_cols(int ncw) { return ncw <= 12 ? 256 : 512; }

__host__ __device__ inline size_t sweep_sub2_w_offset(int stages, int ctot, int cpl, int c0) {
  return ((size_t)stages * sweep_sub_stage_bytes(ctot, cpl, c0) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t sweep_sub2_smem_bytes(int stages, int ctot, int K, int cpl, int c0) {
  return sweep_sub2_w_offset(stages, ctot, cpl, c0) + (size_t)K * cpl * 16 + (size_t)stages * 16 + 16;
}

// SPEC 0: run-time nonlinearities / distribution; 1: softplus subunits, softplus output, Poisson.
// Backward sweeps only (the forward-only sweep has no accumulators to move and stays on sweep_subunit_kernel).
template <int K, int NJ, int SPEC, int NCW>
__global__ void __launch_bounds__(sub2_threads(NCW), 1) sweep_subunit_pool_kernel(const __grid_constant__ SweepArgs a) {
  sweep_stamp(a, 0);
  constexpr int kSub2Consumers = NCW;
  constexpr int kSub2TmemCols = sub2_tmem_cols(NCW);
  static_assert(NCW >= 4 && NCW <= 15 && ((NCW + 3) / 4) * 80 <= kSub2TmemCols, "tensor-memory columns");
  constexpr int LPR = 16, RT = kSubRows, NRG = 32 / LPR, NP = RT / NRG;
  constexpr int KP = K <= 2 ? 2 : 4, VN = NP * KP;
  static_assert(K >= 2 && K <= kSubMaxK && NJ >= 1 && NJ <= 5 && LPR * NJ <= kSubMaxChunks, "shape");
  extern __shared__ __align__(128) unsigned char sweep_smem[];
  const int lane = threadIdx.x & 31, l8 = lane & (LPR - 1), rg = lane / LPR;
  const int warp = threadIdx.x >> 5;
  const int ctot = a.n_chunks * 4;
  const int NE = a.layout_heads * ctot + kNumScalars;
  const int NS = a.n_stages;
  const int S = a.sub_slot, cs0 = a.slot_chunk0[S], C0 = a.slot_chunk0[S + 1] - cs0;
  const uint32_t stage_bytes = (uint32_t)sweep_sub_stage_bytes(ctot, LPR * NJ, C0);
  const uint32_t y_off = (uint32_t)(RT * ctot * sizeof(float));
  const uint32_t ldS = (uint32_t)(a.slot_ld[S] * sizeof(float));

  const uint32_t sbase = smem_u32(sweep_smem);
  const uint32_t wsm = sbase + (uint32_t)sweep_sub2_w_offset(NS, ctot, LPR * NJ, C0);
  constexpr uint32_t wstride = (uint32_t)LPR * NJ * 16u;
  constexpr uint32_t jstep = (uint32_t)LPR * 16u;
  const uint32_t bars = wsm + (uint32_t)K * wstride;            // full[s] at bars + 16 s, empty[s] at + 8
  const uint32_t tmem_slot = bars + (uint32_t)NS * 16u;
  const uint32_t issued_at = tmem_slot + 8u;                    // tiles the producer has armed so far

  // ---- head weights -> shared memory, barriers, tensor memory ----------------------------------------
  for (int idx = threadIdx.x; idx < K * LPR * NJ; idx += blockDim.x) {
    const int k = idx / (LPR * NJ), c = idx - k * (LPR * NJ);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C0) v = *reinterpret_cast<const float4*>(a.W + (size_t)a.sub_head[k] * ctot + 4 * (cs0 + c));
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wsm + (uint32_t)idx * 16u), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  }
  for (uint32_t o = threadIdx.x * 4u; o < (uint32_t)NS * stage_bytes; o += blockDim.x * 4u)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + o), "r"(0u) : "memory");
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init_a(bars + 16u * s, 1);
      mbar_init_a(bars + 16u * s + 8u, 1);
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(issued_at), "r"(0u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kSub2TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const long long ntiles = (a.n_rows + RT - 1) / RT;
  const long long first = blockIdx.x, G = gridDim.x;
  const long long n_local = first < ntiles ? (ntiles - first + G - 1) / G : 0;   // this CTA's tiles: first + i G

  if (warp == kSub2Consumers) {
    // ================================ producer ================================
    const char* my_src = nullptr;
    uint32_t my_rowbytes = 0, my_dst = 0, all_rowbytes = 0;
    for (int s = 0; s < a.n_slots; ++s) {
      all_rowbytes += (uint32_t)(a.slot_ld[s] * sizeof(float));
      if (lane == s) {
        my_src = reinterpret_cast<const char*>(a.slot_ptr[s]);
        my_rowbytes = (uint32_t)(a.slot_ld[s] * sizeof(float));
        my_dst = (uint32_t)(RT * 4 * a.slot_chunk0[s] * sizeof(float));
      }
    }
    const bool has_y = a.y != nullptr;
    const bool y_lane = has_y && lane == a.n_slots;
    if (y_lane) {
      my_src = reinterpret_cast<const char*>(a.y);
      my_rowbytes = sizeof(float);
      my_dst = y_off;
    }
    const uint32_t tile_y_bytes = has_y ? 32u : 0u;
    const uint64_t policy = l2_evict_first_policy();
    int s = 0;
    uint32_t use = 0;
    for (long long i = 0; i < n_local; ++i) {
      const uint32_t full = bars + 16u * (uint32_t)s, empty = full + 8u;
      if (use > 0) mbar_wait_a(empty, (use - 1u) & 1u);          // the consumer of the previous tenant is done
      __syncwarp();
      const long long row0 = (first + i * G) * RT;
      const long long left = a.n_rows - row0;
      const uint32_t rows = left < RT ? (uint32_t)left : (uint32_t)RT;
      if (lane == 0) {
        mbar_arrive_expect_tx_a(full, rows * all_rowbytes + tile_y_bytes);
        // A consumer may only parity-wait on this use of the stage once it is armed: with a pool, the warp
        // that takes tile i is not the one that took tile i - NS, so nothing else orders its wait after the
        // previous phase (a parity wait two phases early would pass at once).
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(issued_at), "r"((uint32_t)(i + 1)) : "memory");
      }
      if (my_rowbytes != 0) {
        const uint32_t bytes = y_lane ? 32u : rows * my_rowbytes;
        bulk_g2s_a(sbase + (uint32_t)s * stage_bytes + my_dst, my_src + row0 * my_rowbytes, bytes, full, policy);
      }
      if (++s == NS) {
        s = 0;
        ++use;
      }
    }
  } else {
    // ================================ consumers ================================
    const int cw = warp;
    const uint32_t wl8o = (wsm - sbase) + (uint32_t)l8 * 16u;
    // tensor-memory address: lane quarter in bits 16+, column in the low bits
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 80);

    const int E = a.n_chunks - C0;
    const int he = a.sub_extra_head;
    uint32_t eoff = (uint32_t)(RT * 16 * cs0), estride = ldS;
    float4 we = make_float4(0.f, 0.f, 0.f, 0.f), ge = make_float4(0.f, 0.f, 0.f, 0.f);
    int ecol = -1;
    if (l8 < E && he >= 0) {
      const int c = l8 < cs0 ? l8 : l8 + C0;
      int s = 0;
      while (s + 1 < a.n_slots && c >= a.slot_chunk0[s + 1]) ++s;
      eoff = (uint32_t)(RT * 16 * a.slot_chunk0[s] + 16 * (c - a.slot_chunk0[s]));
      estride = (uint32_t)(a.slot_ld[s] * sizeof(float));
      we = *reinterpret_cast<const float4*>(a.W + (size_t)he * ctot + 4 * c);
      ecol = 4 * c;
    }
    const int my_idx = l8 & (VN - 1);
    const int my_k = my_idx % KP, my_p = my_idx / KP;
    const bool head_ok = my_k < K;
    const bool lane_counts = l8 < VN;
    const int my_head = head_ok ? a.sub_head[my_k] : 0;
    const float hic_lane = head_ok ? a.head_icpt[my_head] : 0.f;
    const int kind_lane = head_ok ? a.head_nl[my_head] : RFB_NL_NONE;
    const int kind_e = he >= 0 ? a.head_nl[he] : kNlAbsent;
    const float hic_e = he >= 0 ? a.head_icpt[he] : 0.f;
    const float gic = a.glob_icpt[0];
    const bool poisson = SPEC == 1 ? true : (a.distr == RFB_DISTR_POISSON);
    const int out_kind = SPEC == 1 ? RFB_NL_SOFTPLUS : a.out_nl;
    const bool has_y = a.y != nullptr;

    float sc[S_GHEAD];
#pragma unroll
    for (int q = 0; q < S_GHEAD; ++q) sc[q] = 0.f;
    float sc_head = 0.f, sc_e = 0.f;
    double* my_acc = a.warp_acc + ((size_t)blockIdx.x * kSub2Consumers + cw) * NE;
    for (int e = lane; e < NE; e += 32) my_acc[e] = 0.0;

    auto zero_tmem = [&]() {
      uint32_t z[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) z[q] = 0u;
#pragma unroll
      for (int j = 0; j < NJ; ++j) tmem_st16(tacc + 16u * j, z);
      tmem_wait_st();
    };
    zero_tmem();

    auto flush = [&]() {
      // tensor-memory accumulators -> this warp's fp64 row (the two row groups hold partial sums of the same columns)
      tmem_wait_st();
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        uint32_t r[16];
        tc::tmem_ld16(tacc + 16u * j, r);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[q] = __uint_as_float(r[4 * k + q]);
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
          }
          if (rg == (j % NRG) && l8 + LPR * j < C0) {
            double2* p2 = reinterpret_cast<double2*>(my_acc + (size_t)a.sub_head[k] * ctot + 4 * (cs0 + l8 + LPR * j));
            double2 lo = p2[0], hi = p2[1];
            lo.x += (double)v[0]; lo.y += (double)v[1];
            hi.x += (double)v[2]; hi.y += (double)v[3];
            p2[0] = lo;
            p2[1] = hi;
          }
        }
      }
      zero_tmem();
      float v[4] = {ge.x, ge.y, ge.z, ge.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
      if (rg == 0 && ecol >= 0) {
        double2* p2 = reinterpret_cast<double2*>(my_acc + (size_t)he * ctot + ecol);
        double2 lo = p2[0], hi = p2[1];
        lo.x += (double)v[0]; lo.y += (double)v[1];
        hi.x += (double)v[2]; hi.y += (double)v[3];
        p2[0] = lo;
        p2[1] = hi;
      }
      ge = make_float4(0.f, 0.f, 0.f, 0.f);
      double* scal = my_acc + (size_t)a.layout_heads * ctot;
#pragma unroll
      for (int q = 0; q < S_GHEAD; ++q) {
        const double t = warp_sum_f64((double)sc[q]);
        if (lane == 0) scal[q] += t;
        sc[q] = 0.f;
      }
      {
        double t = (double)sc_head;
#pragma unroll
        for (int m = KP; m < 32; m <<= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (lane < K) scal[S_GHEAD + a.sub_head[lane]] += t;
        sc_head = 0.f;
        const double te = warp_sum_f64((double)sc_e);
        if (lane == 0 && he >= 0) scal[S_GHEAD + he] += te;
        sc_e = 0.f;
      }
    };

    const unsigned char* const smem_gen = sweep_smem;
    auto ld2 = [&](uint32_t off, u64_t& lo, u64_t& hi) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(smem_gen + off);
      lo = v.x;
      hi = v.y;
    };
    const uint32_t xlane = (uint32_t)(RT * 16 * cs0) + (uint32_t)rg * ldS + (uint32_t)l8 * 16u;

    int since_flush = 0;
    for (long long i = cw; i < n_local; i += kSub2Consumers) {
      const uint32_t s = (uint32_t)(i % NS), parity = (uint32_t)((i / NS) & 1);
      const uint32_t so = s * stage_bytes;
      const long long tile = first + i * G;
      {
        uint32_t seen;
        do {
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(issued_at) : "memory");
        } while (seen < (uint32_t)(i + 1));
      }
      mbar_wait_a(bars + 16u * s, parity);

      // ---- forward products (as sweep_subunit_kernel) ----------------------------------------------------
      float pv[VN], ae[NP], yv;
      float4 xe[NP];
      {
        u64_t acc[K][NP];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int p = 0; p < NP; ++p) acc[k][p] = 0ull;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          u64_t tl[NP], th[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) ld2(so + xlane + (uint32_t)(p * NRG) * ldS + jstep * j, tl[p], th[p]);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            u64_t wl, wh;
            ld2(wl8o + (uint32_t)k * wstride + jstep * j, wl, wh);
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[k][p] = ffma2(tl[p], wl, acc[k][p]);
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[k][p] = ffma2(th[p], wh, acc[k][p]);
          }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          xe[p] = *reinterpret_cast<const float4*>(smem_gen + so + eoff + (uint32_t)(p * NRG + rg) * estride);
          ae[p] = fmaf(xe[p].x, we.x, fmaf(xe[p].y, we.y, fmaf(xe[p].z, we.z, xe[p].w * we.w)));
        }
        yv = 0.f;
        if (has_y) yv = *reinterpret_cast<const float*>(smem_gen + so + y_off + 4u * (uint32_t)(my_p * NRG + rg));
#pragma unroll
        for (int q = 0; q < VN; ++q) pv[q] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float lo, hi;
            unpack2(acc[k][p], lo, hi);
            pv[k + KP * p] = lo + hi;
          }
        }
      }
      // ---- cross-lane sums ----------------------------------------------------------------------------------
      const float tot = transpose_reduce_grp<VN, LPR>(pv, lane);
#pragma unroll
      for (int st = NP / 2; st >= 1; st >>= 1) {
        const bool up = (l8 & (st * KP)) != 0;
#pragma unroll
        for (int j = 0; j < st; ++j) {
          const float keep = up ? ae[j + st] : ae[j];
          const float send = up ? ae[j] : ae[j + st];
          ae[j] = keep + __shfl_xor_sync(0xffffffffu, send, st * KP);
        }
      }
      float aet = ae[0];
#pragma unroll
      for (int m = 1; m < LPR; m <<= 1)
        if (m < KP || m >= NP * KP) aet += __shfl_xor_sync(0xffffffffu, aet, m);

      // ---- epilogue: lane = (row group, pass, head) ------------------------------------------------------------
      const long long myrow = tile * RT + my_p * NRG + rg;
      const bool rowok = myrow < a.n_rows;
      if (!rowok) yv = 0.f;
      float f, dphi;
      nl_eval(SPEC == 1 ? RFB_NL_SOFTPLUS : kind_lane, tot + hic_lane, f, dphi);
      float u = head_ok ? f : 0.f;
#pragma unroll
      for (int m = 1; m < KP; m <<= 1) u += __shfl_xor_sync(0xffffffffu, u, m);
      float fe, dpe;
      if (SPEC == 1) {
        fe = kind_e == RFB_NL_NONE ? aet + hic_e : 0.f;
        dpe = kind_e == RFB_NL_NONE ? 1.f : 0.f;
      } else {
        nl_eval(kind_e, aet + hic_e, fe, dpe);
      }
      u += fe;
      u += gic;
      float r, dr;
      nl_eval(out_kind, u, r, dr);
      if (SPEC == 0) softmax_scale(a, out_kind, r);
      float la, lb, dldr;
      const float err = yv - r;
      if (poisson) {                                     // rfest/loss.py:4-11
        const bool fin = r != CUDART_INF_F;
        const float r1 = fin ? r : 0.f;
        const bool live = fin && (r1 > 1e-20f);
        const float r2 = (r1 != r1) ? r1 : fmaxf(r1, 1e-20f);
        la = -logf(r2) * yv;
        lb = r2;
        dldr = live ? (1.0f - fast_div(yv, r2)) : 0.f;
      } else {                                           // rfest/loss.py:14-16
        la = err * err;
        lb = 0.f;
        dldr = -err * a.two_over_n;
      }
      float delta = (dldr == 0.f || !rowok) ? 0.f : dldr * dr;
      if (SPEC == 0) delta = softmax_delta(a, out_kind, delta, r, dldr, rowok);
      const bool sm_out = SPEC == 0 && out_kind == RFB_NL_SOFTMAX;
      const bool count_row = rowok && lane_counts && my_k == 0;
      if (count_row) {
        sc[S_LOSS_A] += la;
        sc[S_LOSS_B] += lb;
        sc[S_R] += r;
        sc[S_RR] += r * r;
        sc[S_YR] += yv * r;
        sc[S_SQERR] += err * err;
        if (!sm_out) sc[S_GGLOBAL] += delta;
        if (a.r_out != nullptr) a.r_out[myrow] = r;
      }
      const float delta_k = (delta == 0.f || !head_ok) ? 0.f : delta * dphi;
      const float delta_e = (delta == 0.f) ? 0.f : delta * dpe;
      if (lane_counts && !(sm_out && kind_lane == RFB_NL_NONE)) sc_head += delta_k;
      if (count_row && !(sm_out && kind_e == RFB_NL_NONE)) sc_e += delta_e;

      // ---- backward: G_k += delta_k x, one column chunk (all heads) at a time through 16 registers ----------
      asm volatile("" ::: "memory");                     // read the rows again, do not carry them in registers
      u64_t dd[K][NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float d = __shfl_sync(0xffffffffu, delta_k, (lane & ~(LPR - 1)) | (k + KP * p));
          dd[k][p] = pack2(d, d);
        }
        const float de = __shfl_sync(0xffffffffu, delta_e, (lane & ~(LPR - 1)) | (KP * p));
        ge.x = fmaf(de, xe[p].x, ge.x);
        ge.y = fmaf(de, xe[p].y, ge.y);
        ge.z = fmaf(de, xe[p].z, ge.z);
        ge.w = fmaf(de, xe[p].w, ge.w);
      }
      tmem_wait_st();                                    // the previous tile's stores have landed
      uint32_t rcur[16], rnext[16];
      tc::tmem_ld16(tacc, rcur);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        tmem_wait_ld();
        if (j + 1 < NJ) tc::tmem_ld16(tacc + 16u * (j + 1), rnext);   // in flight during this chunk's products
        u64_t glo[K], ghi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          glo[k] = pack_u32(rcur[4 * k], rcur[4 * k + 1]);
          ghi[k] = pack_u32(rcur[4 * k + 2], rcur[4 * k + 3]);
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          u64_t tl, th;
          ld2(so + xlane + (uint32_t)(p * NRG) * ldS + jstep * j, tl, th);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            glo[k] = ffma2(dd[k][p], tl, glo[k]);
            ghi[k] = ffma2(dd[k][p], th, ghi[k]);
          }
        }
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = 0u;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          unpack_u32(glo[k], w[4 * k], w[4 * k + 1]);
          unpack_u32(ghi[k], w[4 * k + 2], w[4 * k + 3]);
        }
        tmem_st16(tacc + 16u * j, w);
        if (j + 1 < NJ) {
#pragma unroll
          for (int q = 0; q < 16; ++q) rcur[q] = rnext[q];
        }
      }
      // both passes have read the stage: hand it back to the producer
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 16u * s + 8u) : "memory");
      if (++since_flush >= kSubFlushTiles) {
        flush();
        since_flush = 0;
      }
    }
    flush();
  }

  // ---- fixed-order fp64 reduction over the CTA's consumer warps ------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  const double* cta_rows = a.warp_acc + (size_t)blockIdx.x * kSub2Consumers * NE;
  double* out = a.cta_part + (size_t)blockIdx.x * NE;
  for (int e = threadIdx.x; e < NE; e += blockDim.x) {
    double t = 0.0;
    for (int wv = 0; wv < kSub2Consumers; ++wv) t += cta_rows[(size_t)wv * NE + e];
    out[e] = t;
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kSub2TmemCols) : "memory");
  }
  sweep_stamp(a, 1);
}

}  // namespace rfb
