// K2s: the streaming sweep for subunit models (LNLN: `num_subunits` filters that alias ONE projected
// design block, rfest/GLM/_glm.py:338-352, each with its own filter nonlinearity, plus at most one
// other head such as the history filter).
//
// The row-per-warp kernels of sweep.cuh keep every head's weights in registers and reduce each
// (head, row) dot product over all 32 lanes: with 4 subunits + history that is 250 warp instructions
// per row (62 forward + 62 backward FFMA, 31 for the transposed reductions, 30 for one epilogue chain per
// 4 rows ...) against a budget of ~200 issue slots per row at HBM speed -- instruction-bound at half the
// HBM rate (profiles/README.md).  This kernel changes the mapping instead of the arithmetic:
//   * a row is owned by LPR = 16 lanes (lane l holds the float4 chunks l, l + 16, ...), a warp works on
//     2 rows at a time and a tile is 4 such passes (8 rows): a (head, row) dot product is reduced over 16
//     lanes, the 4 heads x 4 passes of a lane's rows by ONE transposed butterfly of 15 shuffles;
//   * the head weights live in shared memory (one broadcast LDS.128 per head and chunk serves both row
//     groups and all passes) -- registers hold only the gradient accumulators, so no columns are padded
//     to the warp width (288 columns = 4.5 chunks per lane);
//   * products are packed (fma.rn.f32x2): the 8-byte halves of a chunk against the same halves of the
//     weights forward, (delta, delta) against the chunk backward;
//   * the rows are read from the shared-memory stage twice (forward, backward) instead of being held;
//   * the epilogue runs once per tile with lane = (row group, pass, head): all 32 lanes evaluate one
//     filter nonlinearity each.
// About 115 warp instructions per row for 4 subunits + history (was 250).  Same staging (per-warp ring of
// bulk-async copies), same accumulator layout, same fixed-order fp64 reductions as sweep.cuh:
// bit-reproducible.  Measurements and the mappings that lost (8 lanes per row, rows held in registers,
// a software-pipelined loop): profiles/sweep_sub_r1_ncu.txt, DESIGN.md section 3.2b.
#pragma once

#include "project_tc.cuh"
#include "sweep.cuh"

namespace rfb {

// Tensor memory as accumulator storage for ordinary warps (no MMA involved): tcgen05.ld / tcgen05.st move 16
// consecutive 32-bit columns of this lane between registers and tensor memory.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ u64_t pack_u32(uint32_t lo, uint32_t hi) {
  u64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack_u32(u64_t v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}

constexpr int kSubRows = 8;      // rows per tile: (32 / LPR) row groups x passes
constexpr int kSubMaxK = 4;      // subunit heads
constexpr int kSubMaxChunks = 80;   // chunks of the shared block (LPR x NJ <= 80: 320 columns)
constexpr int kSubFlushTiles = 64;

// Shared memory.  A stage = the 8-row tile of every design block, 32 B of responses and a zero pad: lane l
// (of a row's LPR lanes) reads the chunks l + LPR j, j < NJ, of its rows with compile-time offsets; the chunk
// slots past the block's C0 chunks have zero weights and read on into the next row / block / responses /
// pad -- finite data that the pad keeps inside the stage.  Then the mbarriers, then the weights
// [K][LPR NJ] float4 (zero past C0).  `cpl` = LPR x NJ: chunk slots per row.
__host__ __device__ inline size_t sweep_sub_stage_bytes(int ctot, int cpl, int c0) {
  return (size_t)kSubRows * ctot * sizeof(float) + 32 + (size_t)(cpl - c0) * 16;
}
__host__ __device__ inline size_t sweep_sub_w_offset(int warps, int stages, int ctot, int cpl, int c0) {
  return ((size_t)warps * stages * (sweep_sub_stage_bytes(ctot, cpl, c0) + 8) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t sweep_sub_smem_bytes(int warps, int stages, int ctot, int K, int cpl, int c0) {
  return sweep_sub_w_offset(warps, stages, ctot, cpl, c0) + (size_t)K * cpl * 16;
}

// p[i] = this lane's partial of value i; afterwards every lane of an LPR-lane group holds the group's
// sum of value (lane & (VN - 1)).
template <int VN, int LPR>
__device__ __forceinline__ float transpose_reduce_grp(float (&p)[VN], int lane) {
#pragma unroll
  for (int s = VN / 2; s >= 1; s >>= 1) {
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
  for (int m = VN; m < LPR; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// SPEC 0: run-time nonlinearities / distribution; 1: softplus subunits, softplus output, Poisson
// LPR lanes per row (8 or 16); HOLD: the rows stay in registers from the forward to the backward pass
// (the stage goes back to the copy engine after the forward loads) instead of being read twice.
// TM: the gradient accumulators (K x NJ float4 per lane, 80 registers for 4 subunits) live in tensor memory
// and pass through 16 registers per column chunk during the backward products -- which is what makes HOLD fit.
template <int K, int NJ, int LPR, bool HOLD, bool BWD, int SPEC, bool TM = false>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_subunit_kernel(const __grid_constant__ SweepArgs a) {
  sweep_stamp(a, 0);
  constexpr bool TMA = TM && BWD;        // a forward-only sweep has no accumulators
  static_assert(K >= 2 && K <= kSubMaxK && NJ >= 1 && LPR * NJ <= kSubMaxChunks, "shape");
  static_assert(LPR == 8 || LPR == 16, "lanes per row");
  constexpr int RT = kSubRows;
  constexpr int NRG = 32 / LPR;        // row groups of a warp
  constexpr int NP = RT / NRG;         // passes per tile
  constexpr int KP = K <= 2 ? 2 : 4;   // heads padded to a power of two
  constexpr int VN = NP * KP;          // (head, pass) values per row group
  static_assert(VN <= LPR, "one epilogue lane per (head, pass)");
  constexpr bool KEEP = HOLD && BWD;
  extern __shared__ __align__(128) unsigned char sweep_smem[];
  const int lane = threadIdx.x & 31, l8 = lane & (LPR - 1), rg = lane / LPR;
  const int wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpc + wic;
  const long long GW = (long long)gridDim.x * wpc;
  const int ctot = a.n_chunks * 4;
  const int NE = a.layout_heads * ctot + kNumScalars;
  const int NS = a.n_stages;
  const int S = a.sub_slot, cs0 = a.slot_chunk0[S], C0 = a.slot_chunk0[S + 1] - cs0;
  const uint32_t stage_bytes = (uint32_t)sweep_sub_stage_bytes(ctot, LPR * NJ, C0);
  const uint32_t y_off = (uint32_t)(RT * ctot * sizeof(float));
  const uint32_t ldS = (uint32_t)(a.slot_ld[S] * sizeof(float));

  const uint32_t sbase = smem_u32(sweep_smem);
  const uint32_t wbase = sbase + (uint32_t)wic * NS * stage_bytes;
  const uint32_t bbase = sbase + (uint32_t)wpc * NS * stage_bytes + (uint32_t)wic * NS * 8u;
  const uint32_t wsm = sbase + (uint32_t)sweep_sub_w_offset(wpc, NS, ctot, LPR * NJ, C0);
  constexpr uint32_t wstride = (uint32_t)LPR * NJ * 16u;
  constexpr uint32_t jstep = (uint32_t)LPR * 16u;
  const uint32_t wl8 = wsm + (uint32_t)l8 * 16u;

  // ---- head weights -> shared memory -----------------------------------------------------------------
  for (int idx = threadIdx.x; idx < K * LPR * NJ; idx += blockDim.x) {
    const int k = idx / (LPR * NJ), c = idx - k * (LPR * NJ);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C0) v = *reinterpret_cast<const float4*>(a.W + (size_t)a.sub_head[k] * ctot + 4 * (cs0 + c));
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wsm + (uint32_t)idx * 16u), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  }

  // the other head: lane l8 of a row's lanes owns chunk l8 of the remaining columns (at most 8 chunks)
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

  // ---- this lane's place in the epilogue: (row group, pass, head) -----------------------------------
  const int my_idx = l8 & (VN - 1);
  const int my_k = my_idx % KP, my_p = my_idx / KP;
  const bool head_ok = my_k < K;
  const bool lane_counts = l8 < VN;                 // K <= 2: the upper half of a group's lanes are duplicates
  const int my_head = head_ok ? a.sub_head[my_k] : 0;
  const float hic_lane = head_ok ? a.head_icpt[my_head] : 0.f;
  const int kind_lane = head_ok ? a.head_nl[my_head] : RFB_NL_NONE;
  const int kind_e = he >= 0 ? a.head_nl[he] : kNlAbsent;
  const float hic_e = he >= 0 ? a.head_icpt[he] : 0.f;
  const float gic = a.glob_icpt[0];
  const bool poisson = SPEC == 1 ? true : (a.distr == RFB_DISTR_POISSON);
  const int out_kind = SPEC == 1 ? RFB_NL_SOFTPLUS : a.out_nl;

  u64_t glo[TMA ? 1 : K][TMA ? 1 : NJ], ghi[TMA ? 1 : K][TMA ? 1 : NJ];
#pragma unroll
  for (int k = 0; k < (TMA ? 1 : K); ++k)
#pragma unroll
    for (int j = 0; j < (TMA ? 1 : NJ); ++j) glo[k][j] = ghi[k][j] = 0ull;
  // tensor memory: warp w touches lanes [32 (w % 4), +32); the warps of a lane quarter take 80 columns each:
  // column 16 j + 4 k + {0..3} = gradient of chunk slot j, head k
  __shared__ uint32_t tm_slot;
  uint32_t tacc = 0;
  constexpr uint32_t kTmCols = 256;
  if (TMA) {
    if (wic == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_slot)), "r"(kTmCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tacc = tm_slot + ((uint32_t)((wic & 3) * 32) << 16) + (uint32_t)((wic >> 2) * 80);
  }
  auto zero_tmem = [&]() {
    uint32_t z[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) z[q] = 0u;
#pragma unroll
    for (int j = 0; j < NJ; ++j) tmem_st16(tacc + 16u * j, z);
    tmem_wait_st();
  };
  if (TMA) zero_tmem();
  float sc[S_GHEAD];
#pragma unroll
  for (int k = 0; k < S_GHEAD; ++k) sc[k] = 0.f;
  float sc_head = 0.f, sc_e = 0.f;
  double* my_acc = a.warp_acc + (size_t)gw * NE;
  for (int e = lane; e < NE; e += 32) my_acc[e] = 0.0;

  // ---- the copies: lane s moves design block s, lane n_slots moves the responses --------------------
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
  const long long ntiles = (a.n_rows + RT - 1) / RT;

  auto issue = [&](long long tile, uint32_t stage_addr, uint32_t bar) {
    const long long row0 = tile * RT;
    const long long left = a.n_rows - row0;
    const uint32_t rows = left < RT ? (uint32_t)left : (uint32_t)RT;
    if (lane == 0) mbar_arrive_expect_tx_a(bar, rows * all_rowbytes + tile_y_bytes);
    if (my_rowbytes != 0) {
      const uint32_t bytes = y_lane ? 32u : rows * my_rowbytes;
      bulk_g2s_a(stage_addr + my_dst, my_src + row0 * my_rowbytes, bytes, bar, policy);
    }
  };

  for (uint32_t o = lane * 4u; o < (uint32_t)NS * stage_bytes; o += 128u)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(wbase + o), "r"(0u) : "memory");
  if (lane == 0)
    for (int s = 0; s < NS; ++s) mbar_init_a(bbase + 8u * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();                                   // weights visible to every warp
  for (int s = 0; s < NS; ++s)
    if (gw + (long long)s * GW < ntiles) issue(gw + (long long)s * GW, wbase + s * stage_bytes, bbase + 8u * s);

  auto flush = [&]() {
    if (BWD) {
      if (TMA) tmem_wait_st();
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        uint32_t tr[16];
        if (TMA) {
          tc::tmem_ld16(tacc + 16u * j, tr);
          tmem_wait_ld();
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float v[4];
          if (TMA) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = __uint_as_float(tr[4 * k + q]);
          } else {
            unpack2(glo[TMA ? 0 : k][TMA ? 0 : j], v[0], v[1]);
            unpack2(ghi[TMA ? 0 : k][TMA ? 0 : j], v[2], v[3]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {               // the row groups hold partial sums of the same columns
#pragma unroll
            for (int m = LPR; m < 32; m <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], m);
          }
          if (rg == (j % NRG) && l8 + LPR * j < C0) {
            double2* p2 = reinterpret_cast<double2*>(my_acc + (size_t)a.sub_head[k] * ctot + 4 * (cs0 + l8 + LPR * j));
            double2 lo = p2[0], hi = p2[1];
            lo.x += (double)v[0]; lo.y += (double)v[1];
            hi.x += (double)v[2]; hi.y += (double)v[3];
            p2[0] = lo;
            p2[1] = hi;
          }
          if (!TMA) glo[TMA ? 0 : k][TMA ? 0 : j] = ghi[TMA ? 0 : k][TMA ? 0 : j] = 0ull;
        }
      }
      if (TMA) zero_tmem();
      float v[4] = {ge.x, ge.y, ge.z, ge.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int m = LPR; m < 32; m <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], m);
      }
      if (rg == 0 && ecol >= 0) {
        double2* p2 = reinterpret_cast<double2*>(my_acc + (size_t)he * ctot + ecol);
        double2 lo = p2[0], hi = p2[1];
        lo.x += (double)v[0]; lo.y += (double)v[1];
        hi.x += (double)v[2]; hi.y += (double)v[3];
        p2[0] = lo;
        p2[1] = hi;
      }
      ge = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double* scal = my_acc + (size_t)a.layout_heads * ctot;
#pragma unroll
    for (int k = 0; k < S_GHEAD; ++k) {
      const double v = warp_sum_f64((double)sc[k]);
      if (lane == 0) scal[k] += v;
      sc[k] = 0.f;
    }
    {
      double v = (double)sc_head;                      // lanes with the same head: l8 bits >= KP, row groups
#pragma unroll
      for (int m = KP; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (lane < K) scal[S_GHEAD + a.sub_head[lane]] += v;
      sc_head = 0.f;
      const double ve = warp_sum_f64((double)sc_e);
      if (lane == 0 && he >= 0) scal[S_GHEAD + he] += ve;
      sc_e = 0.f;
    }
  };

  // Row and weight reads are ordinary shared-memory loads (not volatile asm): the compiler may move them
  // and the arithmetic between the shuffles (measured: 275 -> 288 it/s at config 4).  The barrier waits
  // and the copy issue are asm statements with memory clobbers, so no load crosses them.
  const unsigned char* const smem_gen = sweep_smem;
  auto ld2 = [&](uint32_t off, u64_t& lo, u64_t& hi) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(smem_gen + off);
    lo = v.x;
    hi = v.y;
  };
  const uint32_t xlane = (uint32_t)(RT * 16 * cs0) + (uint32_t)rg * ldS + (uint32_t)l8 * 16u;
  const uint32_t wl8o = wl8 - sbase;

  // forward products of the tile in the stage at byte offset `so`: per (head, pass) partial dot products of
  // this lane's chunks -> pv, the other head's -> ae; the response of this lane's epilogue row -> yv
  auto fwd_products = [&](uint32_t so, float (&pv)[VN], float (&ae)[NP], float4 (&xe)[NP], float& yv,
                          u64_t (&xl)[KEEP ? NP : 1][KEEP ? NJ : 1], u64_t (&xh)[KEEP ? NP : 1][KEEP ? NJ : 1]) {
    u64_t acc[K][NP];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int p = 0; p < NP; ++p) acc[k][p] = 0ull;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      u64_t tl[NP], th[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        ld2(so + xlane + (uint32_t)(p * NRG) * ldS + jstep * j, tl[p], th[p]);
        if (KEEP) {
          xl[p][j] = tl[p];
          xh[p][j] = th[p];
        }
      }
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
    for (int i = 0; i < VN; ++i) pv[i] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float lo, hi;
        unpack2(acc[k][p], lo, hi);
        pv[k + KP * p] = lo + hi;
      }
    }
  };

  // the cross-lane sums: this lane's (head, pass) dot product and the other head's of its pass
  auto fwd_reduce = [&](float (&pv)[VN], float (&ae)[NP], float& tot, float& aet) {
    tot = transpose_reduce_grp<VN, LPR>(pv, lane);
#pragma unroll
    for (int s = NP / 2; s >= 1; s >>= 1) {
      const bool up = (l8 & (s * KP)) != 0;
#pragma unroll
      for (int j = 0; j < s; ++j) {
        const float keep = up ? ae[j + s] : ae[j];
        const float send = up ? ae[j] : ae[j + s];
        ae[j] = keep + __shfl_xor_sync(0xffffffffu, send, s * KP);
      }
    }
    aet = ae[0];
#pragma unroll
    for (int m = 1; m < LPR; m <<= 1)
      if (m < KP || m >= NP * KP) aet += __shfl_xor_sync(0xffffffffu, aet, m);
  };

  // epilogue with lane = (row group, pass, head), then the gradient products of the tile
  auto epilogue_backward = [&](uint32_t so, long long tile, float tot, float aet, float yv, const float4 (&xe)[NP],
                               u64_t (&xl)[KEEP ? NP : 1][KEEP ? NJ : 1], u64_t (&xh)[KEEP ? NP : 1][KEEP ? NJ : 1]) {
    const long long myrow = tile * RT + my_p * NRG + rg;
    const bool rowok = myrow < a.n_rows;
    if (!rowok) yv = 0.f;
    float f, dphi;
    nl_eval(SPEC == 1 ? RFB_NL_SOFTPLUS : kind_lane, tot + hic_lane, f, dphi);
    float u = head_ok ? f : 0.f;
#pragma unroll
    for (int m = 1; m < KP; m <<= 1) u += __shfl_xor_sync(0xffffffffu, u, m);
    float fe, dpe;
    if (SPEC == 1) {                                   // the other head is linear (or absent): no branch
      fe = kind_e == RFB_NL_NONE ? aet + hic_e : 0.f;
      dpe = kind_e == RFB_NL_NONE ? 1.f : 0.f;
    } else {
      nl_eval(kind_e, aet + hic_e, fe, dpe);           // warp-uniform kind
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

    // backward: G_k += delta_k x, pass by pass (rows from registers, or read again from the stage)
    if (TMA) {
      // accumulators in tensor memory: one column chunk (all heads) at a time through 16 registers, the next
      // chunk's accumulators in flight meanwhile
      if (!KEEP) asm volatile("" ::: "memory");
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
      tmem_wait_st();                                  // the previous tile's stores have landed
      uint32_t rcur[16], rnext[16];
      tc::tmem_ld16(tacc, rcur);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        tmem_wait_ld();
        if (j + 1 < NJ) tc::tmem_ld16(tacc + 16u * (j + 1), rnext);
        u64_t gl[K], gh[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          gl[k] = pack_u32(rcur[4 * k], rcur[4 * k + 1]);
          gh[k] = pack_u32(rcur[4 * k + 2], rcur[4 * k + 3]);
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          u64_t tl, th;
          if (KEEP) {
            tl = xl[p][j];
            th = xh[p][j];
          } else {
            ld2(so + xlane + (uint32_t)(p * NRG) * ldS + jstep * j, tl, th);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            gl[k] = ffma2(dd[k][p], tl, gl[k]);
            gh[k] = ffma2(dd[k][p], th, gh[k]);
          }
        }
        uint32_t wv[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) wv[q] = 0u;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          unpack_u32(gl[k], wv[4 * k], wv[4 * k + 1]);
          unpack_u32(gh[k], wv[4 * k + 2], wv[4 * k + 3]);
        }
        tmem_st16(tacc + 16u * j, wv);
        if (j + 1 < NJ) {
#pragma unroll
          for (int q = 0; q < 16; ++q) rcur[q] = rnext[q];
        }
      }
    } else if (BWD) {
      if (!KEEP) asm volatile("" ::: "memory");        // read the rows again, do not carry them in registers
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        u64_t dd[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float d = __shfl_sync(0xffffffffu, delta_k, (lane & ~(LPR - 1)) | (k + KP * p));
          dd[k] = pack2(d, d);
        }
        const float de = __shfl_sync(0xffffffffu, delta_e, (lane & ~(LPR - 1)) | (KP * p));
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          u64_t tl, th;
          if (KEEP) {
            tl = xl[p][j];
            th = xh[p][j];
          } else {
            ld2(so + xlane + (uint32_t)(p * NRG) * ldS + jstep * j, tl, th);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            glo[TMA ? 0 : k][TMA ? 0 : j] = ffma2(dd[k], tl, glo[TMA ? 0 : k][TMA ? 0 : j]);
            ghi[TMA ? 0 : k][TMA ? 0 : j] = ffma2(dd[k], th, ghi[TMA ? 0 : k][TMA ? 0 : j]);
          }
        }
        ge.x = fmaf(de, xe[p].x, ge.x);
        ge.y = fmaf(de, xe[p].y, ge.y);
        ge.z = fmaf(de, xe[p].z, ge.z);
        ge.w = fmaf(de, xe[p].w, ge.w);
      }
    }
  };

  auto refill = [&](long long tile, int buf) {         // hand the stage to the copy engine for tile + NS * GW
    const long long nt = tile + (long long)NS * GW;
    __syncwarp();
    if (nt < ntiles) issue(nt, wbase + (uint32_t)buf * stage_bytes, bbase + 8u * (uint32_t)buf);
  };
  const uint32_t wbo = wbase - sbase;                  // this warp's ring, as an offset into shared memory

  int since_flush = 0;
  int buf = 0;
  uint32_t phase = 0;
  for (long long tile = gw; tile < ntiles; tile += GW) {
    const uint32_t so = wbo + (uint32_t)buf * stage_bytes;
    mbar_wait_a(bbase + 8u * (uint32_t)buf, phase);
    float pv[VN], ae[NP], yv, tot, aet;
    float4 xe[NP];
    u64_t xl[KEEP ? NP : 1][KEEP ? NJ : 1], xh[KEEP ? NP : 1][KEEP ? NJ : 1];
    fwd_products(so, pv, ae, xe, yv, xl, xh);
    if (KEEP || !BWD) refill(tile, buf);             // the rows are in registers
    fwd_reduce(pv, ae, tot, aet);
    epilogue_backward(so, tile, tot, aet, yv, xe, xl, xh);
    if (BWD && !KEEP) refill(tile, buf);             // both passes have read the stage
    if (++buf == NS) {
      buf = 0;
      phase ^= 1u;
    }
    if (++since_flush >= kSubFlushTiles) {
      flush();
      since_flush = 0;
    }
  }
  flush();

  // fixed-order fp64 reduction over the CTA's warps
  __syncthreads();
  const double* cta_rows = a.warp_acc + (size_t)blockIdx.x * wpc * NE;
  double* out = a.cta_part + (size_t)blockIdx.x * NE;
  for (int e = (BWD ? 0 : a.layout_heads * ctot) + threadIdx.x; e < NE; e += blockDim.x) {
    double s = 0.0;
    for (int wv = 0; wv < wpc; ++wv) s += cta_rows[(size_t)wv * NE + e];
    out[e] = s;
  }
  if (TMA) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wic == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_slot), "r"(kTmCols) : "memory");
    }
  }
  sweep_stamp(a, 1);
}

}  // namespace rfb
