// K1a on the 5th-generation tensor cores: Z[M x N] = stim[M x K] @ Sxy[K x N]   (M = frames)
//
// The spatial contraction of the projection is the one dense GEMM of the path
// (config 5: M = 2^24, K = 1024, N = 100; configs 2-4: K = 300, N = 36).  Inputs are fp32 and the
// result must keep fp32-level accuracy (parity: losses to 1e-5), which TF32 alone (10-bit mantissa)
// cannot give, so every operand is split  x = hi + lo  with  hi = tf32(x)  (truncation) and
// lo = x - hi  (exact in fp32), and three tcgen05.mma.kind::tf32 products are accumulated in fp32:
//     A.B  ~=  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        (the dropped lo.lo term is ~2^-22 relative)
// The tensor core adds into its fp32 accumulator with truncation (measured: a bias of about
// -0.5 ulp per tcgen05.mma), so the chain of accumulations is kept short: the two small cross terms
// go to their own accumulator, the hi.hi term is dealt round-robin over `n_acc` accumulators, and the
// epilogue adds the partial sums with round-to-nearest fp32.
//
// Structure (one CTA per SM, persistent over 128-row tiles of A; 11 warps):
//   warp 8      A producer: per 32-wide K tile one 2-D TMA load (cp.async.bulk.tensor.2d,
//               SWIZZLE_128B) of the raw 128 x 32 fp32 A tile into a deep ring of raw slots -- this ring
//               is what keeps enough HBM bytes in flight.
//   warp 10     B producer: TMA loads of the pre-split B_hi / B_lo tiles (N x 32, L2 resident) into
//               a ring of shared-memory stages.
//   warps 4-7   splitters: thread = row.  Read the row's 32 values of a landed raw slot (8 LDS.128
//               following the 128-byte swizzle, conflict-free), form A_hi = tf32(A) and
//               A_lo = A - A_hi in registers and write them straight into TENSOR MEMORY with
//               tcgen05.st -- the A operands never go back to shared memory.
//   warp 9      MMA issuer: one elected thread issues 12 tcgen05.mma.kind::tf32 per K tile (4 k-steps x
//               3 products), A from TMEM, B from shared-memory descriptors, D in TMEM;
//               tcgen05.commit releases the A / B stages; every `drain` K tiles it publishes the
//               accumulator and switches to the other one.
//   warps 0-3   accumulate + epilogue: tcgen05.ld the published partial sums (lane = row), add them
//               into fp32 registers with round-to-nearest, and store the finished 128 x N tile of Z.
// Keeping the TMEM accumulation chains short (12 * drain MMAs) is what holds fp32-level accuracy: the
// tensor core adds into its accumulator with truncation (measured bias ~ -0.5 ulp per MMA).
// Shared-memory traffic per K tile: 16 KB TMA write + 16 KB split read + 2 N x 128 B TMA write + 3 N x 128 B
// MMA reads (102 KB at N = 112) -- with SS-mode A operands it was 182 KB and bound the kernel.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "project.cuh"

namespace rfb {

constexpr int kTcM = 128;     // rows of A per tile = TMEM lanes
constexpr int kTcK = 32;      // K tile: 32 fp32 = 128 bytes = one swizzle atom
constexpr int kTcThreads = 352;   // warps 0-3 epilogue (TMEM lanes), 4-7 splitters, 8 A-TMA, 9 MMA, 10 B-TMA
constexpr int kTcMaxN = 128;
// FUSED variant (round 2): the temporal (Hankel) contraction runs in the same kernel.  The epilogue warps put the
// finished 128 x n_q tile of Z into a three-tile ring in shared memory instead of HBM, and four more warps
// (11-14) contract it over the lag window with the temporal basis and write the design block -- Z never leaves
// the SM, and the FMA-bound temporal step overlaps the HBM-bound spatial one.  Every CTA owns a contiguous range
// of tiles (the lag window of a tile reaches nlag - 1 rows back into the previous one; a CTA whose range does not
// start at tile 0 first recomputes the tile in front of it without emitting rows).
// Warp groups (fused): 0-3 epilogue, 4-7 splitters, 8-11 the three single-lane roles (+ one idle warp), 12-19 the
// temporal step.  One temporal warp per scheduler reached a quarter of the FMA rate (measured: 14.4 us per tile
// against 6.8 us of HBM time), so there are two; the kernel is compiled for 96 registers per thread and the
// producer group hands registers to the temporal groups (setmaxnreg 40 / 120).
constexpr int kTcFusedThreads = 640;
constexpr int kTcTemporalThreads = 256;
constexpr int kTcFusedMaxN = 64;
constexpr int kTcRingTiles = 3;
struct TcFuseArgs {
  long long off;        // Z row (relative to the slab's first raw row) that feeds (design row 0, lag 0); <= 0
  long long n_out;      // design rows of this slab
  float* out;           // block storage
  long long out_ld;
  long long out_row0;   // block row of design row 0 of this slab
  int col0;
  int nlag;
  int df_t;
  long long n_tiles;    // Z tiles to run: ceil((off + n_out + nlag - 1) / 128)
};

struct TcArgs {
  long long n_rows;   // M
  int n_k_tiles;      // ceil(K / 32)
  int n;              // real N (columns of Z)
  int n_pad;          // UMMA N (multiple of 16, <= 128)
  int n_stages;       // B ring slots (B_hi, B_lo)
  float* Z;           // [M x ldz]
  int ldz;
  int tmem_cols;      // power of two >= 2 * n_pad (two accumulators) + 128 (two A stages of hi|lo)
  int drain;          // K tiles accumulated in TMEM before the partial sums move to registers
  int products;       // 3 (default): A_lo.B_hi + A_hi.B_lo + A_hi.B_hi; 2 / 1: measurement only (drops the first / first two)
  int n_raw;          // raw A slots
};

constexpr size_t kTcRawBytes = (size_t)kTcM * kTcK * sizeof(float);   // 16 KB
__host__ __device__ inline size_t tc_stage_bytes(int n_pad) {
  return (size_t)2 * n_pad * kTcK * sizeof(float);
}
constexpr int kTcAStages = 2;        // A stages in tensor memory: 32 columns hi + 32 columns lo each
__host__ __device__ inline size_t tc_ring_bytes(int n_q) {   // fused: Z ring, rows of n_q floats
  return (((size_t)kTcRingTiles * kTcM * n_q * sizeof(float)) + 1023) & ~(size_t)1023;
}
__host__ __device__ inline size_t tc_smem_bytes(int n_pad, int stages, int n_raw, size_t ring_bytes = 0) {
  return 1024 /* alignment slack */ + (size_t)stages * tc_stage_bytes(n_pad) + (size_t)n_raw * kTcRawBytes +
         512 /* barriers */ + ring_bytes;
}

namespace tc {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The wait suspends the thread for up to kTcSuspendNs (it resumes at once when the phase completes): without the
// hint try_wait came back after ~4 cycles and the single-lane producer / issuer warps spun through 59 % of all
// issued instructions (ncu, fused variant), taking issue slots from the FMA-bound temporal warps.
constexpr uint32_t kTcSuspendNs = 20000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(kTcSuspendNs)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// low word of the shared-memory descriptor (start address in 16-byte units | leading-dim field); the
// high word (stride 1024 B, version 1, SWIZZLE_128B) is the constant kDescHi
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);
// A operand from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
// Called by every lane of a converged warp; `elected` is non-zero in the one lane that issues.
__device__ __forceinline__ void mma_tf32_ts(uint32_t elected, uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo,
                                            uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      ".reg .b64 bd;\n"
      "mov.b64 bd, {%2, %5};\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.ne.b32 e, %6, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(bdesc_lo), "r"(idesc), "r"(acc), "r"(kDescHi), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void commit_if(uint32_t elected, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "setp.ne.b32 e, %1, 0;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar), "r"(elected)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t e;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(e));
  return e;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

}  // namespace tc

template <bool FUSED, int DFT>
__global__ void __launch_bounds__(FUSED ? kTcFusedThreads : kTcThreads, 1)
spatial_project_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                          const __grid_constant__ CUtensorMap map_blo, const TcArgs a, const TcFuseArgs f) {
  constexpr int MAXN = FUSED ? kTcFusedMaxN : kTcMaxN;
  extern __shared__ unsigned char tc_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned tile area (SWIZZLE_128B atoms are 8 x 128 bytes)
  const uint32_t base = (tc::s32(tc_smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)kTcRawBytes;
  const uint32_t b_bytes = (uint32_t)a.n_pad * kTcK * sizeof(float);
  const uint32_t stage_bytes = 2 * b_bytes;
  const int NS = a.n_stages, NR = a.n_raw, NA = kTcAStages;
  auto st_bhi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
  auto st_blo = [&](int s) { return base + (uint32_t)s * stage_bytes + b_bytes; };
  const uint32_t raw0 = base + (uint32_t)NS * stage_bytes;
  auto st_raw = [&](int r) { return raw0 + (uint32_t)r * a_bytes; };
  const uint32_t bars = raw0 + (uint32_t)NR * a_bytes;              // 8-byte mbarriers
  auto bar_rfull = [&](int r) { return bars + 8u * r; };               // raw A landed              (tx)
  auto bar_rempty = [&](int r) { return bars + 8u * (NR + r); };       // raw slot read             (128)
  auto bar_bfull = [&](int s) { return bars + 8u * (2 * NR + s); };    // B tiles landed            (tx)
  auto bar_bempty = [&](int s) { return bars + 8u * (2 * NR + NS + s); };      // MMAs read B stage (1)
  auto bar_afull = [&](int s) { return bars + 8u * (2 * NR + 2 * NS + s); };   // A hi|lo in TMEM   (128)
  auto bar_aempty = [&](int s) { return bars + 8u * (2 * NR + 2 * NS + NA + s); };  // MMAs read A stage (1)
  auto bar_tfull = [&](int b) { return bars + 8u * (2 * NR + 2 * NS + 2 * NA + b); };      // partial sums published (1)
  auto bar_tempty = [&](int b) { return bars + 8u * (2 * NR + 2 * NS + 2 * NA + 2 + b); }; // ... drained       (128)
  const uint32_t tmem_slot = bars + 8u * (2 * NR + 2 * NS + 2 * NA + 4);
  auto bar_zfull = [&](int k) { return bars + 8u * (2 * NR + 2 * NS + 2 * NA + 6 + k); };                  // Z tile in the ring (128)
  auto bar_tdone = [&](int k) { return bars + 8u * (2 * NR + 2 * NS + 2 * NA + 6 + kTcRingTiles + k); };   // temporal step of a tile done (128)
  float* const zring = reinterpret_cast<float*>(tc_smem_raw + (bars + 512u - tc::s32(tc_smem_raw)));

  if (threadIdx.x == 0) {
    for (int r = 0; r < NR; ++r) {
      tc::mbar_init(bar_rfull(r), 1);
      tc::mbar_init(bar_rempty(r), 128);
    }
    for (int s = 0; s < NS; ++s) {
      tc::mbar_init(bar_bfull(s), 1);
      tc::mbar_init(bar_bempty(s), 1);
    }
    for (int s = 0; s < NA; ++s) {
      tc::mbar_init(bar_afull(s), 128);
      tc::mbar_init(bar_aempty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(bar_tfull(b), 1);
      tc::mbar_init(bar_tempty(b), 128);
    }
    if (FUSED) {
      for (int k = 0; k < kTcRingTiles; ++k) {
        tc::mbar_init(bar_zfull(k), 128);
        tc::mbar_init(bar_tdone(k), kTcTemporalThreads);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {   // TMEM allocation by one warp; it also owns the deallocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // tensor-memory map: [accumulator 0 | accumulator 1 | A stage 0: hi(32) lo(32) | A stage 1 ...]
  const uint32_t tm_acc = tmem_base;
  const uint32_t tm_a = tmem_base + (uint32_t)(2 * a.n_pad);

  const int KT = a.n_k_tiles;
  // tiles of this CTA: [t_begin, t_end) in steps of t_step.  Unfused: round-robin over the grid.  Fused: a
  // contiguous range, preceded by one warm-up tile unless it starts at tile 0.
  long long t_begin, t_end, t_step;
  bool warmup = false;
  if (FUSED) {
    const long long j0 = f.n_tiles * blockIdx.x / gridDim.x, j1 = f.n_tiles * (blockIdx.x + 1) / gridDim.x;
    warmup = j0 > 0;
    t_begin = warmup ? j0 - 1 : 0;
    t_end = j1;
    t_step = 1;
    // the lag window of tile 0 reaches into rows before the recording: zeros (ring tile 2 sits in front of tile 0)
    for (int i = threadIdx.x; i < kTcM * a.n; i += blockDim.x) zring[(size_t)(kTcRingTiles - 1) * kTcM * a.n + i] = 0.f;
    __syncthreads();
  } else {
    t_begin = blockIdx.x;
    t_end = (a.n_rows + kTcM - 1) / kTcM;
    t_step = gridDim.x;
  }

  if (FUSED && warp >= 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    // ================= temporal step (fused variant, warps 12-19) =================
    // Local tile l (global tile g = t_begin + l) completes the 128 design rows whose lag window ends inside it:
    //   k in [128 g - off - nlag + 1, 128 g - off - nlag + 128]  (clipped to [0, n_out)),
    // sixteen groups of 8 rows x n_q columns = tasks; a task's window is Z rows 128 g - nlag + 1 + 8 r + [0, nlag + 6],
    // i.e. ring rows starting nlag - 1 rows in front of ring tile l % 3 (the previous tile's tail).
    if constexpr (FUSED) {
      const int tt = threadIdx.x - 384;
      const int ring_rows = kTcRingTiles * kTcM;
      const int n_tasks = 16 * a.n;
      for (long long tile = t_begin; tile < t_end; ++tile) {
        const long long l = tile - t_begin;
        const int rt = (int)(l % kTcRingTiles);
        tc::mbar_wait(bar_zfull(rt), (uint32_t)((l / kTcRingTiles) & 1));
        if (!(warmup && l == 0)) {
          for (int task = tt; task < n_tasks; task += kTcTemporalThreads) {
            const int r = task / a.n, q = task - r * a.n;
            const long long k_base = tile * kTcM - f.off - f.nlag + 1 + 8 * r;
            if (k_base + 7 < 0 || k_base >= f.n_out) continue;
            int row0 = rt * kTcM - f.nlag + 1 + 8 * r;          // ring row of the window's first Z row
            if (row0 < 0) row0 += ring_rows;
            const float* zq = zring + q;
            const int nq = a.n;
            auto zload = [&](int k) -> float {
              int row = row0 + k;
              if (row >= ring_rows) row -= ring_rows;
              return zq[(size_t)row * nq];
            };
            float acc[8][DFT];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int d = 0; d < DFT; ++d) acc[i][d] = 0.f;
            temporal_accumulate_from<DFT>(f.nlag, zload, acc);
            float* dst = f.out + (f.out_row0 + k_base) * f.out_ld + f.col0 + q;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (k_base + i >= 0 && k_base + i < f.n_out) {
#pragma unroll
                for (int d = 0; d < DFT; ++d)
                  if (d < f.df_t) dst[(size_t)d * nq] = acc[i][d];
              }
              dst += f.out_ld;
            }
          }
        }
        tc::mbar_arrive(bar_tdone(rt));
      }
    }
  } else if (warp >= 8 && warp < 12) {
   if (FUSED) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 8) {
    // ================= A producer (raw ring) =================
    if (lane == 0) {
      int r = 0;
      uint32_t ph = 0;
      for (long long tile = t_begin; tile < t_end; tile += t_step) {
        for (int kt = 0; kt < KT; ++kt) {
          tc::mbar_wait(bar_rempty(r), ph ^ 1u);
          tc::mbar_expect_tx(bar_rfull(r), a_bytes);
          tc::tma_load_2d(st_raw(r), &map_a, kt * kTcK, (int)(tile * kTcM), bar_rfull(r));
          if (++r == NR) { r = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 10) {
    // ================= B producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long tile = t_begin; tile < t_end; tile += t_step) {
        for (int kt = 0; kt < KT; ++kt) {
          tc::mbar_wait(bar_bempty(s), ph ^ 1u);
          tc::mbar_expect_tx(bar_bfull(s), 2 * b_bytes);
          tc::tma_load_2d(st_bhi(s), &map_bhi, kt * kTcK, 0, bar_bfull(s));
          tc::tma_load_2d(st_blo(s), &map_blo, kt * kTcK, 0, bar_bfull(s));
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    // The whole warp walks the loop with warp-uniform control flow and one elected lane issues: written
    // as `if (lane == 0)` around the loop, every tcgen05 instruction got wrapped in a divergence-safe
    // elect/retry sequence and the issuing thread itself (about 190 dependent instructions per K tile)
    // paced the kernel.
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = n_pad, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.n_pad >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
    const uint32_t blo_off = b_bytes >> 4;             // B_lo sits b_bytes behind B_hi (descriptor units of 16 B)
    const uint32_t elected = tc::elect_one();
    int sb = 0, sa = 0, buf = 0, dk = 0;
    uint32_t phb = 0, pha = 0, tph = 0;
    for (long long tile = t_begin; tile < t_end; tile += t_step) {
      for (int kt = 0; kt < KT; ++kt) {
        const bool first = dk == 0;                      // first K tile into this accumulator
        if (first) tc::mbar_wait(bar_tempty(buf), tph ^ 1u);
        tc::mbar_wait(bar_afull(sa), pha);
        tc::mbar_wait(bar_bfull(sb), phb);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool last = (++dk == a.drain) || kt == KT - 1;
        {
          const uint32_t d = tm_acc + (uint32_t)(buf * a.n_pad);
          const uint32_t a_hi = tm_a + (uint32_t)(sa * 64), a_lo = a_hi + 32u;
          const uint32_t bhi = tc::smem_desc_lo(st_bhi(sb)), blo = bhi + blo_off;
#pragma unroll
          for (int ks = 0; ks < kTcK / 8; ++ks) {
            // 8 tf32 = 32 bytes along K inside the swizzle atom = 2 descriptor units, 8 TMEM columns
            const uint32_t ko = 2u * ks, kc = 8u * ks;
            uint32_t accum = (first && ks == 0) ? 0u : 1u;
            if (a.products >= 3) { tc::mma_tf32_ts(elected, d, a_lo + kc, bhi + ko, idesc, accum); accum = 1u; }
            if (a.products >= 2) { tc::mma_tf32_ts(elected, d, a_hi + kc, blo + ko, idesc, accum); accum = 1u; }
            tc::mma_tf32_ts(elected, d, a_hi + kc, bhi + ko, idesc, accum);
          }
          tc::commit_if(elected, bar_aempty(sa));        // stages reusable once these MMAs have read them
          tc::commit_if(elected, bar_bempty(sb));
          if (last) tc::commit_if(elected, bar_tfull(buf));          // partial sums complete
        }
        if (++sa == NA) { sa = 0; pha ^= 1u; }
        if (++sb == NS) { sb = 0; phb ^= 1u; }
        if (last) {
          dk = 0;
          if (++buf == 2) { buf = 0; tph ^= 1u; }
        }
      }
    }
   }
  } else if (warp >= 4 && warp < 8) {
    // ================= splitters (warps 4-7: thread = row of the tile = TMEM lane) =================
    const int t = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int r = 0, sa = 0;
    uint32_t rph = 0, pha = 0;
    for (long long tile = t_begin; tile < t_end; tile += t_step) {
      for (int kt = 0; kt < KT; ++kt) {
        tc::mbar_wait(bar_rfull(r), rph);
        const uint32_t row_addr = st_raw(r) + (uint32_t)t * 128u;
        // two halves of 16 columns: 32 registers of split operands in flight instead of 64
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {                    // chunk j of row t sits at j ^ (t & 7)
            const int j = 4 * half + jj;
            uint32_t x0, x1, x2, x3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                         : "r"(row_addr + (uint32_t)((j ^ (t & 7)) << 4)));
            hi[4 * jj + 0] = x0 & 0xFFFFE000u; hi[4 * jj + 1] = x1 & 0xFFFFE000u;
            hi[4 * jj + 2] = x2 & 0xFFFFE000u; hi[4 * jj + 3] = x3 & 0xFFFFE000u;
            lo[4 * jj + 0] = __float_as_uint(__uint_as_float(x0) - __uint_as_float(hi[4 * jj + 0]));
            lo[4 * jj + 1] = __float_as_uint(__uint_as_float(x1) - __uint_as_float(hi[4 * jj + 1]));
            lo[4 * jj + 2] = __float_as_uint(__uint_as_float(x2) - __uint_as_float(hi[4 * jj + 2]));
            lo[4 * jj + 3] = __float_as_uint(__uint_as_float(x3) - __uint_as_float(hi[4 * jj + 3]));
          }
          if (half == 0) {
            tc::mbar_wait(bar_aempty(sa), pha ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t ta = tm_a + (uint32_t)(sa * 64) + lane_base + 16u * half;
          tc::tmem_st16(ta, hi);
          tc::tmem_st16(ta + 32u, lo);
        }
        // the stores have consumed the registers the shared-memory loads filled, so the loads are
        // complete: only now may the raw slot be handed back to the TMA producer
        tc::mbar_arrive(bar_rempty(r));
        if (++r == NR) { r = 0; rph ^= 1u; }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc::mbar_arrive(bar_afull(sa));
        if (++sa == NA) { sa = 0; pha ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // ================= accumulate + epilogue (warps 0-3: thread = TMEM lane = row of the tile) =========
    const int t = threadIdx.x;
    int buf = 0;
    uint32_t tph = 0;
    const bool vec_ok = (a.ldz % 4) == 0;
    for (long long tile = t_begin; tile < t_end; tile += t_step) {
      float z[MAXN];
#pragma unroll
      for (int j = 0; j < MAXN; ++j) z[j] = 0.f;
      for (int k0 = 0; k0 < KT; k0 += a.drain) {
        tc::mbar_wait(bar_tfull(buf), tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tm_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * a.n_pad);
#pragma unroll
        for (int c0 = 0; c0 < MAXN; c0 += 16) {
          if (c0 < a.n_pad) {
            uint32_t v[16];
            tc::tmem_ld16(taddr + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) z[c0 + j] += __uint_as_float(v[j]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc::mbar_arrive(bar_tempty(buf));
        if (++buf == 2) { buf = 0; tph ^= 1u; }
      }
      if constexpr (FUSED) {
        // ring tile l % 3 was last read (as the lag halo) by the temporal step of local tile l - 2
        const long long l = tile - t_begin;
        const int rt = (int)(l % kTcRingTiles);
        if (l >= 2) tc::mbar_wait(bar_tdone((int)((l - 2) % kTcRingTiles)), (uint32_t)(((l - 2) / kTcRingTiles) & 1));
        float* zr = zring + ((size_t)rt * kTcM + t) * a.n;
#pragma unroll
        for (int c = 0; c < MAXN; c += 4) {
          if (c < a.n) {
            if ((a.n & 3) == 0) {
              *reinterpret_cast<float4*>(zr + c) = make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]);
            } else {
#pragma unroll
              for (int cc = c; cc < c + 4; ++cc)
                if (cc < a.n) zr[cc] = z[cc];
            }
          }
        }
        tc::mbar_arrive(bar_zfull(rt));
        continue;
      }
      const long long row = tile * kTcM + t;
      if (row < a.n_rows) {
        float* zr = a.Z + row * a.ldz;
#pragma unroll
        for (int c = 0; c < MAXN; c += 4) {
          if (c < a.n) {
            if (vec_ok && c + 3 < a.n) {
              *reinterpret_cast<float4*>(zr + c) = make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]);
            } else {
#pragma unroll
              for (int cc = c; cc < c + 4; ++cc)
                if (cc < a.n) zr[cc] = z[cc];
            }
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

}  // namespace rfb
