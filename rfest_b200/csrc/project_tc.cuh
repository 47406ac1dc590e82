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
// Structure (one CTA per SM, persistent over 128-row tiles of A; 10 warps):
//   warp 8      TMA producer: per 32-wide K tile one 2-D bulk-tensor copy of the raw A tile
//               (128 x 32 fp32, SWIZZLE_128B) and of the pre-split B_hi / B_lo tiles (N x 32), all
//               signalling the stage's mbarrier with their byte counts.
//   warps 4-7   splitters: turn the landed raw A tile into A_hi (in place) and A_lo with
//               conflict-free 16-byte LDS/STS (the split is element-wise, hence layout-agnostic),
//               fence the generic->async proxy and hand the stage to the MMA warp.
//   warps 0-3   epilogue: tcgen05.ld the partial 128 x N fp32 accumulators out of TMEM (lane = row),
//               add them and store Z with 16-byte stores, while the next tile is already running.
//   warp 9      MMA issuer: one elected thread issues 12 tcgen05.mma (4 k-steps x 3 products) per
//               stage from shared-memory descriptors into one of two TMEM accumulator buffers,
//               tcgen05.commit releases the stage / publishes the accumulator.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace rfb {

constexpr int kTcM = 128;     // rows of A per tile = TMEM lanes
constexpr int kTcK = 32;      // K tile: 32 fp32 = 128 bytes = one swizzle atom
constexpr int kTcThreads = 320;   // warps 0-3 epilogue (TMEM lanes), 4-7 splitters, 8 TMA, 9 MMA

struct TcArgs {
  long long n_rows;   // M
  int n_k_tiles;      // ceil(K / 32)
  int n;              // real N (columns of Z)
  int n_pad;          // UMMA N (multiple of 16, <= 128)
  int n_stages;
  float* Z;           // [M x ldz]
  int ldz;
  int tmem_cols;      // power of two >= n_bufs * (n_acc + 1) * n_pad
  int n_acc;          // independent accumulators of the hi.hi product (round-robin over k-steps)
  int n_bufs;         // 1 or 2 accumulator sets (2: the epilogue overlaps the next tile's MMAs)
};

__host__ __device__ inline size_t tc_stage_bytes(int n_pad) {
  return (size_t)2 * kTcM * kTcK * sizeof(float) + (size_t)2 * n_pad * kTcK * sizeof(float);
}
__host__ __device__ inline size_t tc_smem_bytes(int n_pad, int stages) {
  return 1024 /* alignment slack */ + (size_t)stages * tc_stage_bytes(n_pad) + 256 /* barriers */;
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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

__global__ void __launch_bounds__(kTcThreads, 1)
spatial_project_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                          const __grid_constant__ CUtensorMap map_blo, const TcArgs a) {
  extern __shared__ unsigned char tc_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned stage area (SWIZZLE_128B atoms are 8 x 128 bytes)
  const uint32_t base = (tc::s32(tc_smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = kTcM * kTcK * sizeof(float);            // 16 KB
  const uint32_t b_bytes = (uint32_t)a.n_pad * kTcK * sizeof(float);
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const int NS = a.n_stages;
  auto st_ahi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
  auto st_alo = [&](int s) { return base + (uint32_t)s * stage_bytes + a_bytes; };
  auto st_bhi = [&](int s) { return base + (uint32_t)s * stage_bytes + 2 * a_bytes; };
  auto st_blo = [&](int s) { return base + (uint32_t)s * stage_bytes + 2 * a_bytes + b_bytes; };
  const uint32_t bars = base + (uint32_t)NS * stage_bytes;        // 8-byte mbarriers
  auto bar_raw = [&](int s) { return bars + 8u * s; };              // TMA landed
  auto bar_split = [&](int s) { return bars + 8u * (NS + s); };     // hi/lo ready
  auto bar_empty = [&](int s) { return bars + 8u * (2 * NS + s); }; // MMAs done with the stage
  auto bar_tfull = [&](int b) { return bars + 8u * (3 * NS + b); };     // accumulator complete
  auto bar_tempty = [&](int b) { return bars + 8u * (3 * NS + 2 + b); }; // accumulator drained
  const uint32_t tmem_slot = bars + 8u * (3 * NS + 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      tc::mbar_init(bar_raw(s), 1);
      tc::mbar_init(bar_split(s), 128);
      tc::mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(bar_tfull(b), 1);
      tc::mbar_init(bar_tempty(b), 128);
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

  const long long n_tiles = (a.n_rows + kTcM - 1) / kTcM;
  const int KT = a.n_k_tiles;

  if (warp == 8) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = a_bytes + 2 * b_bytes;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kt = 0; kt < KT; ++kt) {
          tc::mbar_wait(bar_empty(s), ph ^ 1u);
          tc::mbar_expect_tx(bar_raw(s), tx);
          tc::tma_load_2d(st_ahi(s), &map_a, kt * kTcK, (int)(tile * kTcM), bar_raw(s));
          tc::tma_load_2d(st_bhi(s), &map_bhi, kt * kTcK, 0, bar_raw(s));
          tc::tma_load_2d(st_blo(s), &map_blo, kt * kTcK, 0, bar_raw(s));
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = n_pad, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.n_pad >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
      int s = 0, buf = 0;
      uint32_t ph = 0, tph = 0;
      const uint32_t set_cols = (uint32_t)((a.n_acc + 1) * a.n_pad);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        tc::mbar_wait(bar_tempty(buf), tph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d0 = tmem_base + (uint32_t)buf * set_cols;
        const uint32_t d_small = d0 + (uint32_t)(a.n_acc * a.n_pad);
        uint32_t small_acc = 0;
        int step = 0;
        for (int kt = 0; kt < KT; ++kt) {
          tc::mbar_wait(bar_split(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int ks = 0; ks < kTcK / 8; ++ks, ++step) {
            const uint32_t off = (uint32_t)ks * 32u;   // 8 tf32 = 32 bytes along K inside the swizzle atom
            const uint64_t ahi = tc::smem_desc(st_ahi(s) + off), alo = tc::smem_desc(st_alo(s) + off);
            const uint64_t bhi = tc::smem_desc(st_bhi(s) + off), blo = tc::smem_desc(st_blo(s) + off);
            tc::mma_tf32(d_small, alo, bhi, idesc, small_acc);
            small_acc = 1;
            tc::mma_tf32(d_small, ahi, blo, idesc, 1);
            tc::mma_tf32(d0 + (uint32_t)((step % a.n_acc) * a.n_pad), ahi, bhi, idesc, step >= a.n_acc ? 1u : 0u);
          }
          tc::commit(bar_empty(s));          // stage reusable once these MMAs have read it
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
        tc::commit(bar_tfull(buf));          // accumulators complete
        if (++buf == a.n_bufs) { buf = 0; tph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================= splitters (warps 4-7, 128 threads) =================
    const int t = threadIdx.x - 128;
    int s = 0;
    uint32_t ph = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kt = 0; kt < KT; ++kt) {
        tc::mbar_wait(bar_raw(s), ph);
        const uint32_t hi_base = st_ahi(s), lo_base = st_alo(s);
#pragma unroll
        for (int i = 0; i < (kTcM * kTcK / 4) / 128; ++i) {      // 1024 16-byte chunks / 128 threads
          const uint32_t off = (uint32_t)(t + 128 * i) * 16u;
          uint32_t x0, x1, x2, x3;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(hi_base + off));
          const uint32_t h0 = x0 & 0xFFFFE000u, h1 = x1 & 0xFFFFE000u, h2 = x2 & 0xFFFFE000u, h3 = x3 & 0xFFFFE000u;
          const float l0 = __uint_as_float(x0) - __uint_as_float(h0), l1 = __uint_as_float(x1) - __uint_as_float(h1);
          const float l2 = __uint_as_float(x2) - __uint_as_float(h2), l3 = __uint_as_float(x3) - __uint_as_float(h3);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(hi_base + off), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + off), "f"(l0), "f"(l1), "f"(l2), "f"(l3) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
        tc::mbar_arrive(bar_split(s));
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ================= epilogue (warps 0-3: thread = TMEM lane = row inside the tile) =================
    const int t = threadIdx.x;
    int buf = 0;
    uint32_t tph = 0;
    const bool vec_ok = (a.ldz % 4) == 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      tc::mbar_wait(bar_tfull(buf), tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long row = tile * kTcM + t;
      const uint32_t set_cols = (uint32_t)((a.n_acc + 1) * a.n_pad);
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * set_cols;
      // every accumulator needs at least one k-step to be defined: n_acc <= 4 * KT is enforced on the host
      for (int c0 = 0; c0 < a.n_pad; c0 += 16) {
        uint32_t v[16];
        float sum[16];
        tc::tmem_ld16(taddr + (uint32_t)(a.n_acc * a.n_pad + c0), v);   // the small cross terms first
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) sum[j] = __uint_as_float(v[j]);
        for (int k = 0; k < a.n_acc; ++k) {
          tc::tmem_ld16(taddr + (uint32_t)(k * a.n_pad + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
        }
        if (row < a.n_rows) {
          float* zr = a.Z + row * a.ldz + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (vec_ok && c0 + j + 3 < a.n) {
              *reinterpret_cast<float4*>(zr + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
            } else {
#pragma unroll
              for (int jj = j; jj < j + 4; ++jj)
                if (c0 + jj < a.n) zr[jj] = sum[jj];
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      tc::mbar_arrive(bar_tempty(buf));
      if (++buf == a.n_bufs) { buf = 0; tph ^= 1u; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

}  // namespace rfb
