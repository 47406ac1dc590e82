// Shared definitions of the rfest_b200 CUDA engine (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rfest_b200.h"

namespace rfb {

constexpr int kMaxSlots = 8;    // design blocks bound to one data set
constexpr int kMaxHeads = 16;   // dot products per row (merged 'none' filters count once); the
                                // register-resident kernels take up to 5, the generic one up to 16
constexpr int kNumScalars = 32; // per-sweep scalar sums (8 fixed + one per head, padded)
constexpr int kNlAbsent = 100;  // internal nonlinearity code: head without data, contributes 0

// layout of the scalar part of a sweep's accumulator vector
enum ScalarSlot {
  S_LOSS_A = 0,  // poisson: sum(-log(r') * y)   gaussian: sum((y - r)^2)
  S_LOSS_B = 1,  // poisson: sum(r')
  S_R = 2,       // sum r
  S_RR = 3,      // sum r^2
  S_YR = 4,      // sum y r
  S_SQERR = 5,   // sum (y - r)^2
  S_GGLOBAL = 6, // d loss / d intercept['global'] = sum delta
  S_BAD = 7,     // rows whose prediction is NaN
  S_GHEAD = 8    // + h: d loss / d (head intercept) = sum delta * phi_h'(a_h)
};

struct SweepArgs {
  const float* slot_ptr[kMaxSlots]; // block base pointers (row-major)
  long long slot_ld[kMaxSlots];     // row strides in floats (multiples of 4)
  int slot_chunk0[kMaxSlots + 1];   // prefix sums of float4 chunks per slot
  int n_slots;
  int n_chunks;                     // total float4 chunks per row = Ctot / 4
  long long n_rows;
  const float* y;                   // [n_rows] or nullptr (prediction only)
  const float* W;                   // [H][Ctot] dense head weights
  const float* head_icpt;           // [H]
  const float* glob_icpt;           // [1]
  int head_nl[kMaxHeads];
  unsigned head_slotmask[kMaxHeads];  // bit i: head has columns among chunks [32 i, 32 i + 32)
  int out_nl;
  int distr;
  float two_over_n;                 // gaussian: 2 / n_global
  double* warp_acc;                 // [total warps][NE] scratch (L2 resident)
  double* cta_part;                 // [grid][NE] per-CTA partial sums (output)
  float* r_out;                     // optional [n_rows]
  int flush_tiles;                  // fp32 -> fp64 flush cadence, in tiles
  int n_stages;                     // staged variant: depth of each warp's copy ring
  const float* softmax;             // out_nl == RFB_NL_SOFTMAX: {1 / Z, D} on the device (see softmax_scale)
  // 'softmax' as a FILTER nonlinearity (head_softmax_fix in sweep.cuh): head_scale[h] = 1 / Z_h on the device;
  // prep != 0 marks the preparatory sweep that produces Z_h = sum exp(a_h) and E_h = sum exp(a_h) x
  const float* head_scale;
  int prep;
  // subunit sweep (sweep_sub.cuh): K heads share the block of slot `sub_slot`; at most one other head
  int sub_slot;
  int sub_head[4];                  // head index (accumulator layout) of every subunit
  int sub_extra_head;               // head over the remaining (<= 8) chunks, -1: none
  int layout_heads;                 // heads in the accumulator layout (the plan's template head count)
  // in-graph timing of the fit loop: when `stamp` is set every CTA folds %globaltimer into
  // stamp[2 * (*stamp_it)] (min at entry) and stamp[2 * (*stamp_it) + 1] (max at exit)
  unsigned long long* stamp;
  const int* stamp_it;
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void sweep_stamp(const SweepArgs& a, int end) {
  if (a.stamp != nullptr && threadIdx.x == 0) {
    unsigned long long* s = a.stamp + 2 * (size_t)(*a.stamp_it) + end;
    if (end) atomicMax(s, global_ns());
    else atomicMin(s, global_ns());
  }
}

// One-shot exchange over NVLink peer memory (multi-GPU): every rank owns a buffer
//   [flags: 2 parities x kMaxRanks uint64][data: 2 parities x world x cap doubles]
// mapped into all peers (CUDA IPC).  The reduction kernel of rank r stores its reduced vector
// into slot r of EVERY rank's buffer and then publishes the iteration's sequence number in the
// flag; the update kernel waits for all ranks' flags and adds the slots in rank order (so all
// ranks compute bit-identical sums).  Two parities: a rank can run at most one iteration ahead.
constexpr int kMaxRanks = 8;
constexpr int kXchgDataOffset = 256;   // bytes: flags first
struct XchgArgs {
  unsigned char* peer[kMaxRanks];      // every rank's buffer (peer[rank] is the local one)
  int world;                           // 1: exchange disabled
  int rank;
  long long cap;                       // doubles per slot
  unsigned long long* seq;             // local device counter: completed exchanges
  unsigned int* done;                  // local device counter for "last CTA" detection
};

// Packed fp32 arithmetic: two IEEE fused multiply-adds on a 64-bit register pair for one issue slot
// (fma.rn.f32x2, SASS FFMA2; bit-identical to the scalar ones).
typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t pack2(float lo, float hi) {
  u64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64_t ffma2(u64_t a, u64_t b, u64_t c) {
  u64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__host__ __device__ inline int sweep_num_entries(int H, int n_chunks) {
  return H * n_chunks * 4 + kNumScalars;
}

}  // namespace rfb
