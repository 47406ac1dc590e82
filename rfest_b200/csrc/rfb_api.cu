// C ABI of the rfest_b200 engine (see include/rfest_b200.h).
//
// Host-side runtime: engine (device / stream / NCCL), design blocks, model plan (filters ->
// heads), data sets, the CUDA-graph fit loop.  The kernels live in sweep.cuh (K2/K4),
// update.cuh (K3) and project.cuh (K1).
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gram.cuh"
#include "band.cuh"
#include "project.cuh"
#include "project_tc.cuh"
#include "sweep.cuh"
#include "sweep_sub.cuh"
#include "sweep_sub2.cuh"
#include "softmax.cuh"
#include "update.cuh"
#include "fit_small.cuh"

using namespace rfb;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(RFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),       \
                  __FILE__, __LINE__);                                                        \
  } while (0)

#define TRY(call)              \
  do {                         \
    int rc_ = (call);          \
    if (rc_ != RFB_OK) return rc_; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, loaded at run time (the copy torch already mapped, or the system one)
// ------------------------------------------------------------------------------------------------
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

static int nccl_load() {
  if (g_nccl.lib) return RFB_OK;
  const char* names[] = {getenv("RFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  }
  if (!lib) return fail(RFB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  NcclApi a;
  a.lib = lib;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
  a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString)
    return fail(RFB_ERR_NCCL, "libnccl is missing a required symbol");
  g_nccl = a;
  return RFB_OK;
}

#define NC(call)                                                                              \
  do {                                                                                        \
    int r_ = (call);                                                                          \
    if (r_ != 0)                                                                              \
      return fail(RFB_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_));           \
  } while (0)

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
struct Scratch {
  void* p = nullptr;
  size_t bytes = 0;
};

struct rfb_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  int64_t launches = 0;
  void* comm = nullptr;
  int rank = 0, world = 1;
  Scratch stage, zbuf, stbuf, sxybuf, small, bsplit, gram_part;
  // host -> device staging ring of rfb_block_project: kRingSlots pinned bounce buffers + device slabs on a copy
  // stream, so the host memcpy of slab i+1, the DMA of slab i and the projection kernels of slab i-1 overlap
  static constexpr int kRingSlots = 3;
  cudaStream_t copy_stream = nullptr;
  void* ring_pinned[kRingSlots] = {};
  void* ring_dev[kRingSlots] = {};
  size_t ring_bytes = 0;
  cudaEvent_t ring_copied[kRingSlots] = {}, ring_consumed[kRingSlots] = {};
  int h2d_threads = 8;    // host threads filling a pinned slot from pageable memory
  int h2d_overlap = 1;    // 0: the round-1 path (one staging buffer, synchronous per slab) for comparison
  double last_project_h2d_bytes = 0.0;
  int sweep_pool_ctas = 0;     // CTAs of the pooled subunit sweep (0: one per SM); tests shrink it to force stage reuse
  int fit_small_ctas = 0;      // CTAs of the persistent small-model kernel (0: sized from the row count)
  int project_fused = 0;       // 1: spatial + temporal projection of a slab in ONE kernel where the shape allows (measured slower: DESIGN 3.1)
  int gram_mma = 1;            // Gram matrices on the FP64 tensor-core path (gram_mma_kernel); 0: register-tile DFMA kernel
  int fit_persistent = 1;      // small models: one persistent cooperative kernel for the whole loop (fit_small.cuh)
  int project_slab_rows = 0;   // rows per projection slab (0: sized from the frame size); tests shrink it
  int project_tc = 1;     // spatial projection on tensor cores (tcgen05, 3xTF32) when the shape allows
  // one-shot NVLink exchange (multi-GPU): local buffer, peers' mappings, device counters
  int use_xchg = 1;
  bool xchg_ready = false;
  unsigned char* xchg_local = nullptr;
  unsigned char* xchg_peer[kMaxRanks] = {};
  unsigned long long* xchg_seq = nullptr;
  unsigned int* xchg_done = nullptr;
  long long xchg_cap = 8192;   // doubles per slot
  // sweep launch shape (tunable: rfb_engine_set_option / RFB_SWEEP_* environment variables)
  int sweep_staged = 1;   // 1: bulk-async staged rows (cp.async.bulk ring), 0: direct streaming loads
  int sweep_stages = 0;   // depth of each warp's copy ring (0: sized for ~112 KB in flight per SM)
  int sweep_warps = 4;    // warps per CTA
  int sweep_sub = 3;      // subunit models (several heads on one block) on the subunit sweep (sweep_sub.cuh):
                          // 0 = off, 3 = default mapping, 1 / 2 = the mappings it was measured against
  int sweep_ctas = -1;    // CTAs per SM (0: as many as fit; -1: auto = 1 for single-head models, where
                          // 4 warps x 3 stages measured best, as many as fit for the compute-heavier multi-head ones)
};

static int scratch_reserve(rfb_engine* e, Scratch& s, size_t bytes) {
  if (s.bytes >= bytes) return RFB_OK;
  if (s.p) {
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaFree(s.p));
    s.p = nullptr;
    s.bytes = 0;
  }
  CU(cudaMalloc(&s.p, bytes));
  s.bytes = bytes;
  return RFB_OK;
}

static int use_device(const rfb_engine* e) {
  CU(cudaSetDevice(e->device));
  return RFB_OK;
}

struct rfb_block {
  rfb_engine* e = nullptr;
  float* d = nullptr;
  int64_t n_rows = 0;
  int n_cols = 0;
  int64_t ld = 0;
};

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct Filter {
  int slot, nl, n_coef;
};

struct DataSet {
  bool bound = false;
  int n_slots = 0;
  rfb_block* blocks[kMaxSlots] = {};
  float* y = nullptr;      // device copy, owned
  bool has_y = false;
  int64_t n = 0;
  double n_total = 0, sy = 0, syy = 0;
  float* r_buf = nullptr;  // predictions of the last forward pass that asked for them
};


// floats in a data set's prediction buffer: n + 4 rounded up to whole float4 rows (a [rows x 4] block of the
// same capacity can trade storage with it, rfb_fit_predictions_swap)
static inline size_t r_buf_floats(int64_t n) { return ((size_t)n + 4 + 3) / 4 * 4; }

struct SweepBuffers {
  double* warp_acc = nullptr;
  double* cta_part = nullptr;
  size_t warp_acc_bytes = 0, cta_part_bytes = 0;
  int grid = 0;
  int threads = 0;
  int staged = 0, stages = 0;
  size_t smem = 0;
  int sub = 0;        // this data set runs the subunit sweep: variant (see sub_instance), 0 = no
  int sub_nj = 0;
  // gradient sweeps of subunit models on the stage-pool / tensor-memory kernel (sweep_sub2.cuh)
  bool pool = false;
  int pool_ncw = 0;   // consumer warps per CTA
  int pool_grid = 0, pool_stages = 0;
  size_t pool_smem = 0;
  int cur_grid = 0;   // CTAs (= partial rows) of the sweep launched last: what the reduction adds up
};

// subunit pattern: K heads whose columns are ONE shared block, at most one other head on <= 8 chunks
struct SubPlan {
  bool ok = false;
  int slot = 0, K = 0, c0 = 0;
  int heads[kSubMaxK] = {};
  int extra_head = -1;
};

struct Plan {
  bool ok = false;
  int F = 0, P = 0;
  int H = 0;   // heads in use
  int Ht = 0;  // template head count (>= H)
  int NCH = 0; // template chunks per lane
  int RB = 0;
  int n_smx = 0;          // heads whose filter nonlinearity is 'softmax' (a normalisation over all samples)
  bool generic = false;   // shape outside the register-resident kernels: generic fallback sweep
  int n_slots = 0;
  int slot_chunk0[kMaxSlots + 1] = {};
  int n_chunks = 0, ctot = 0, NE = 0;
  std::vector<int> head_of;    // [F]
  std::vector<int> head_nl;    // [Ht]
  std::vector<unsigned> head_slotmask;  // [Ht] chunk slots (of 32 chunks) each head has columns in
  std::vector<int> entry_of;   // [P]
  std::vector<int> coef0;      // [F] offset of the filter's coefficients in b
  SubPlan sub;
};

struct ParamSet {  // device-side parameter image + what the sweep reads
  float* theta = nullptr;
  double* ic = nullptr;
  float* W = nullptr;
  float* head_icpt = nullptr;
  float* glob_icpt = nullptr;
  int* entry_of = nullptr;
  int* head_of = nullptr;
};

struct Fit {
  bool active = false;
  int max_iters = 0, executed = 0, metric = 0;
  bool has_dev = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int kernels_per_iter = 0;
  double* step = nullptr;
  double *cost_train = nullptr, *cost_dev = nullptr, *metric_train = nullptr, *metric_dev = nullptr;
  float* ring_b = nullptr;
  double* ring_ic = nullptr;
  int* it = nullptr;
  float *m = nullptr, *v = nullptr;
  double *mi = nullptr, *vi = nullptr;
  double* fin = nullptr;  // [NE + kNumScalars]: train entries, then dev scalars
  UpdateArgs ua{};
  bool finished = false;
  // %globaltimer stamps of the sweeps, written inside the replayed graph: [2 kinds][max_iters + 1][start, end]
  float* bc_f = nullptr;        // Adam bias corrections per iteration (see UpdateArgs)
  double* bc_d = nullptr;
  unsigned long long* stamps = nullptr;
  bool stamp_now = false;       // set while enqueue_iteration builds / launches an iteration
  // small models: the whole loop as one persistent cooperative kernel (fit_small.cuh)
  bool small = false;
  SmallArgs sa{};
  double* small_part = nullptr;
  unsigned long long* small_flags = nullptr;
  unsigned long long small_epoch = 0;
  int small_grid = 0;
  size_t small_smem = 0;
  const void* small_fn = nullptr;
};

struct rfb_model {
  rfb_engine* e = nullptr;
  int distr = 0, out_nl = 0;
  std::vector<Filter> filters;
  int slot_cols[kMaxSlots] = {};  // padded widths (ld) once known, 0 = unknown
  DataSet data[3];
  Plan plan;
  ParamSet cur, tmp;  // current parameters; scratch image for rfb_model_eval
  std::vector<float> host_b;
  std::vector<double> host_ic;
  bool have_params = false;
  SweepBuffers sb[3];
  double* fin_eval = nullptr;
  // output softmax: per data set {Z, sum (dL/dr) r, sum r} (double), block partials, {1 / Z, D} (float)
  double* sm_scal[3] = {};
  double* sm_part[3] = {};
  float* sm_f[3] = {};
  // softmax filter heads: per data set the summed vector of the preparatory sweep and 1 / Z_h
  double* smx_fin[3] = {};
  int smx_fin_n[3] = {};
  float* smx_scale[3] = {};
  Fit fit;
};

// ------------------------------------------------------------------------------------------------
// sweep dispatch
// ------------------------------------------------------------------------------------------------
typedef void (*SweepFn)(const SweepArgs);

template <int H, int NCH, int RB, int SPEC>
static SweepFn pick_staged(bool bwd) {
  return bwd ? (SweepFn)sweep_staged_kernel<H, NCH, RB, true, SPEC> : (SweepFn)sweep_staged_kernel<H, NCH, RB, false, SPEC>;
}

template <int H, int NCH, int RB>
static SweepFn pick_variant(bool bwd, bool staged, int spec) {
  if (!staged)
    return bwd ? (SweepFn)sweep_direct_kernel<H, NCH, RB, true> : (SweepFn)sweep_direct_kernel<H, NCH, RB, false>;
  if constexpr (H == 1) {   // the compile-time model specialisations exist for single-head models
    if (spec == 1) return pick_staged<H, NCH, RB, 1>(bwd);
    if (spec == 2) return pick_staged<H, NCH, RB, 2>(bwd);
    if (spec == 3) return pick_staged<H, NCH, RB, 3>(bwd);
  }
  return pick_staged<H, NCH, RB, 0>(bwd);
}

template <int K, int NJ, int LPR, bool HOLD, bool TM = false>
static SweepFn pick_sub(bool bwd, int spec) {
  if (spec == 1)
    return bwd ? (SweepFn)sweep_subunit_kernel<K, NJ, LPR, HOLD, true, 1, TM>
               : (SweepFn)sweep_subunit_kernel<K, NJ, LPR, false, false, 1, false>;
  return bwd ? (SweepFn)sweep_subunit_kernel<K, NJ, LPR, HOLD, true, 0, TM>
             : (SweepFn)sweep_subunit_kernel<K, NJ, LPR, false, false, 0, false>;
}
// variant 3 (default, all shapes): 16 lanes per row, rows read twice from the stage.  For 4 subunits the two
// mappings it was measured against are kept: 1 = 8 lanes per row, 2 = 16 lanes per row with the rows held in
// registers between the passes (profiles/README.md).
static int sub_lpr(int variant) { return variant == 1 ? 8 : 16; }
static int sub_template_nj(int variant, int c0) {
  if (variant == 1) return c0 <= 24 ? 3 : 9;
  return c0 <= 16 ? 1 : 5;
}
static SweepFn sub_instance(int variant, int K, int NJ, bool bwd, int spec) {
#define INST(v, k, nj, lpr, hold) \
  if (variant == v && K == k && NJ == nj) return pick_sub<k, nj, lpr, hold>(bwd, spec);
  INST(1, 4, 3, 8, false) INST(1, 4, 9, 8, false)
  INST(2, 4, 1, 16, true) INST(2, 4, 5, 16, true)
  INST(3, 2, 1, 16, false) INST(3, 3, 1, 16, false) INST(3, 4, 1, 16, false)
  INST(3, 2, 5, 16, false) INST(3, 3, 5, 16, false) INST(3, 4, 5, 16, false)
#undef INST
  // 6: accumulators in tensor memory + rows held in registers between the passes; 7: tensor memory only
#define INSTT(v, k, nj, hold) \
  if (variant == v && K == k && NJ == nj) return pick_sub<k, nj, 16, hold, true>(bwd, spec);
  INSTT(6, 2, 1, true) INSTT(6, 3, 1, true) INSTT(6, 4, 1, true) INSTT(6, 2, 5, true) INSTT(6, 3, 5, true) INSTT(6, 4, 5, true)
  INSTT(7, 2, 1, false) INSTT(7, 3, 1, false) INSTT(7, 4, 1, false) INSTT(7, 2, 5, false) INSTT(7, 3, 5, false) INSTT(7, 4, 5, false)
#undef INSTT
  return nullptr;
}

static SweepFn pool_instance(int K, int NJ, int spec, int ncw) {
#define INST(k, nj)                                                                                              \
  if (K == k && NJ == nj) {                                                                                      \
    if (ncw == 15) return spec == 1 ? (SweepFn)sweep_subunit_pool_kernel<k, nj, 1, 15> : (SweepFn)sweep_subunit_pool_kernel<k, nj, 0, 15>; \
    return spec == 1 ? (SweepFn)sweep_subunit_pool_kernel<k, nj, 1, 12> : (SweepFn)sweep_subunit_pool_kernel<k, nj, 0, 12>;              \
  }
  INST(2, 1) INST(3, 1) INST(4, 1) INST(2, 5) INST(3, 5) INST(4, 5)
#undef INST
  return nullptr;
}

static int template_heads(int H) { return H <= 1 ? 1 : H <= 2 ? 2 : H <= 3 ? 3 : 5; }
static int template_chunks(int nch) { return nch <= 1 ? 1 : nch <= 3 ? 3 : nch <= 5 ? 5 : 10; }

// rows per tile for each (heads, chunks-per-lane) instance: bounded by registers
static SweepFn sweep_instance(int Ht, int NCH, bool bwd, bool staged, int spec, int* RB) {
#define INST(h, c, rb)                                \
  if (Ht == h && NCH == c) {                          \
    *RB = rb;                                         \
    return pick_variant<h, c, rb>(bwd, staged, spec); \
  }
  INST(1, 1, 8) INST(1, 3, 8) INST(1, 5, 4) INST(1, 10, 2)
  INST(2, 1, 8) INST(2, 3, 8) INST(2, 5, 4) INST(2, 10, 1)
  INST(3, 1, 8) INST(3, 3, 8) INST(3, 5, 4)
  INST(5, 1, 4) INST(5, 3, 4)
#undef INST
  return nullptr;
}

static int build_plan(rfb_model* m) {
  Plan& p = m->plan;
  p = Plan();
  p.F = (int)m->filters.size();
  if (p.F == 0) return fail(RFB_ERR_STATE, "model has no filters");
  if (p.F + 1 > kUpdateThreads) return fail(RFB_ERR_UNSUPPORTED, "too many filters (%d)", p.F);
  int n_slots = 0;
  for (auto& f : m->filters) n_slots = std::max(n_slots, f.slot + 1);
  p.n_slots = n_slots;
  for (int s = 0; s < n_slots; ++s) {
    if (m->slot_cols[s] == 0)
      return fail(RFB_ERR_STATE, "slot %d has no design block bound in any data set", s);
    p.slot_chunk0[s + 1] = p.slot_chunk0[s] + m->slot_cols[s] / 4;
  }
  p.n_chunks = p.slot_chunk0[n_slots];
  p.ctot = p.n_chunks * 4;

  // heads: every filter with a nonlinearity is its own dot product; 'none' filters on
  // different slots add up linearly and share one (their intercepts add, too).
  struct Head { int nl; unsigned slots; };
  std::vector<Head> heads;
  p.head_of.resize(p.F);
  for (int f = 0; f < p.F; ++f) {
    const Filter& fl = m->filters[f];
    int h = -1;
    if (fl.nl == RFB_NL_NONE) {
      for (int k = 0; k < (int)heads.size(); ++k)
        if (heads[k].nl == RFB_NL_NONE && !(heads[k].slots & (1u << fl.slot))) { h = k; break; }
    }
    if (h < 0) {
      heads.push_back({fl.nl, 0u});
      h = (int)heads.size() - 1;
    }
    heads[h].slots |= 1u << fl.slot;
    p.head_of[f] = h;
  }
  p.H = (int)heads.size();
  if (p.H > kMaxHeads)
    return fail(RFB_ERR_UNSUPPORTED, "%d dot products per sample (max %d)", p.H, kMaxHeads);
  const int nch = (p.n_chunks + 31) / 32;
  p.Ht = p.H <= 5 ? template_heads(p.H) : p.H;
  p.NCH = nch <= 10 ? template_chunks(nch) : nch;
  if (p.H > 5 || nch > 10 || !sweep_instance(p.Ht, p.NCH, true, false, 0, &p.RB)) {
    // too wide / too many heads for the register-resident kernels: generic fallback
    p.generic = true;
    p.Ht = p.H;
    p.RB = kGenRB;
  }
  p.head_nl.assign(p.Ht, RFB_NL_NONE);
  for (int h = 0; h < p.H; ++h) {
    p.head_nl[h] = heads[h].nl;
    if (heads[h].nl == RFB_NL_SOFTMAX) ++p.n_smx;
  }
  p.NE = sweep_num_entries(p.Ht, p.n_chunks);
  if (!p.generic && p.H >= 2 && p.n_smx == 0) {   // (softmax heads need the run-time-kind row-per-warp kernels)
    for (int s = 0; s < n_slots && !p.sub.ok; ++s) {
      SubPlan sp;
      int others = 0;
      bool clean = true;
      for (int h = 0; h < p.H; ++h) {
        if (heads[h].slots == (1u << s)) {
          if (sp.K < kSubMaxK) sp.heads[sp.K] = h;
          ++sp.K;
        } else {
          if (heads[h].slots & (1u << s)) clean = false;   // a merged head also reads the shared block
          sp.extra_head = h;
          ++others;
        }
      }
      sp.slot = s;
      sp.c0 = p.slot_chunk0[s + 1] - p.slot_chunk0[s];
      sp.ok = clean && sp.K >= 2 && sp.K <= kSubMaxK && others <= 1 && sp.c0 <= 72 &&
              p.n_chunks - sp.c0 <= 8 && (others == 1 || p.n_chunks == sp.c0);
      if (sp.ok) p.sub = sp;
    }
  }

  p.coef0.resize(p.F);
  p.P = 0;
  for (int f = 0; f < p.F; ++f) {
    p.coef0[f] = p.P;
    p.P += m->filters[f].n_coef;
  }
  p.entry_of.resize(p.P);
  p.head_slotmask.assign(p.Ht, 0u);
  for (int f = 0; f < p.F; ++f) {
    const Filter& fl = m->filters[f];
    const int c_lo = p.slot_chunk0[fl.slot], c_hi = c_lo + (fl.n_coef + 3) / 4;   // chunks this filter reads
    for (int c = c_lo; c < c_hi; ++c) p.head_slotmask[p.head_of[f]] |= 1u << (c / 32);
    if (fl.n_coef > m->slot_cols[fl.slot])
      return fail(RFB_ERR_INVALID, "filter %d has %d coefficients but its block has %d columns", f,
                  fl.n_coef, m->slot_cols[fl.slot]);
    for (int j = 0; j < fl.n_coef; ++j)
      p.entry_of[p.coef0[f] + j] = p.head_of[f] * p.ctot + 4 * p.slot_chunk0[fl.slot] + j;
  }
  p.ok = true;
  return RFB_OK;
}

static int alloc_paramset(rfb_model* m, ParamSet& ps) {
  const Plan& p = m->plan;
  CU(cudaMalloc(&ps.theta, sizeof(float) * std::max(p.P, 1)));
  CU(cudaMalloc(&ps.ic, sizeof(double) * (p.F + 1)));
  CU(cudaMalloc(&ps.W, sizeof(float) * ((size_t)p.Ht * p.ctot + 4)));
  CU(cudaMalloc(&ps.head_icpt, sizeof(float) * p.Ht));
  CU(cudaMalloc(&ps.glob_icpt, sizeof(float)));
  CU(cudaMalloc(&ps.entry_of, sizeof(int) * std::max(p.P, 1)));
  CU(cudaMalloc(&ps.head_of, sizeof(int) * p.F));
  CU(cudaMemsetAsync(ps.W, 0, sizeof(float) * ((size_t)p.Ht * p.ctot + 4), m->e->stream));
  CU(cudaMemsetAsync(ps.head_icpt, 0, sizeof(float) * p.Ht, m->e->stream));
  CU(cudaMemcpyAsync(ps.entry_of, p.entry_of.data(), sizeof(int) * p.P, cudaMemcpyHostToDevice,
                     m->e->stream));
  CU(cudaMemcpyAsync(ps.head_of, p.head_of.data(), sizeof(int) * p.F, cudaMemcpyHostToDevice,
                     m->e->stream));
  CU(cudaStreamSynchronize(m->e->stream));
  return RFB_OK;
}

static void free_paramset(ParamSet& ps) {
  cudaFree(ps.theta); cudaFree(ps.ic); cudaFree(ps.W); cudaFree(ps.head_icpt);
  cudaFree(ps.glob_icpt); cudaFree(ps.entry_of); cudaFree(ps.head_of);
  ps = ParamSet();
}

static UpdateArgs publish_args(const rfb_model* m, const ParamSet& ps) {
  UpdateArgs a{};
  a.P = m->plan.P;
  a.F = m->plan.F;
  a.H = m->plan.Ht;
  a.ctot = m->plan.ctot;
  a.entry_of = ps.entry_of;
  a.head_of = ps.head_of;
  a.theta = ps.theta;
  a.ic = ps.ic;
  a.W = ps.W;
  a.head_icpt = ps.head_icpt;
  a.glob_icpt = ps.glob_icpt;
  return a;
}

static int fit_release(rfb_model* m);

static int ensure_plan(rfb_model* m) {
  if (m->plan.ok) return RFB_OK;
  TRY(use_device(m->e));
  TRY(build_plan(m));
  free_paramset(m->cur);
  free_paramset(m->tmp);
  TRY(alloc_paramset(m, m->cur));
  TRY(alloc_paramset(m, m->tmp));
  cudaFree(m->fin_eval);
  CU(cudaMalloc(&m->fin_eval, sizeof(double) * m->plan.NE));
  if (m->have_params) {
    if ((int)m->host_b.size() != m->plan.P || (int)m->host_ic.size() != m->plan.F + 1)
      return fail(RFB_ERR_STATE, "stored parameters do not match the filters");
    CU(cudaMemcpyAsync(m->cur.theta, m->host_b.data(), sizeof(float) * m->plan.P,
                       cudaMemcpyHostToDevice, m->e->stream));
    CU(cudaMemcpyAsync(m->cur.ic, m->host_ic.data(), sizeof(double) * (m->plan.F + 1),
                       cudaMemcpyHostToDevice, m->e->stream));
    publish_params_kernel<<<1, kUpdateThreads, 0, m->e->stream>>>(publish_args(m, m->cur));
    m->e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(m->e->stream));
  }
  return RFB_OK;
}

static void invalidate_plan(rfb_model* m) {
  m->plan.ok = false;
}

// compile-time model specialisation of the staged sweep (see row_epilogue), 0 = none
static int sweep_spec(const rfb_model* m, const int* head_nl) {
  const Plan& p = m->plan;
  if (p.Ht != 1) return 0;
  for (int h = 0; h < p.Ht; ++h)
    if (head_nl[h] != RFB_NL_NONE) return 0;
  if (m->distr == RFB_DISTR_POISSON && m->out_nl == RFB_NL_SOFTPLUS) return 1;
  if (m->distr == RFB_DISTR_POISSON && m->out_nl == RFB_NL_EXPONENTIAL) return 2;
  if (m->distr == RFB_DISTR_GAUSSIAN && m->out_nl == RFB_NL_NONE) return 3;
  return 0;
}

// Size the launch and allocate the scratch of the sweeps over data set `kind` (not capturable,
// so it happens before any graph capture).
static int ensure_sweep_buffers(rfb_model* m, int kind) {
  const Plan& p = m->plan;
  DataSet& ds = m->data[kind];
  rfb_engine* e = m->e;
  SweepBuffers& sb = m->sb[kind];
  if (sb.grid != 0) return RFB_OK;
  int RB = 0;
  const int warps = std::min(std::max(e->sweep_warps, 1), kSweepThreads / 32);
  const int threads = warps * 32;
  if (p.generic) {
    int o = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sweep_generic_kernel<true>, threads, 0));
    const int64_t ntiles = (ds.n + kGenRB - 1) / kGenRB;
    int64_t grid = (int64_t)e->num_sms * std::max(o, 1);
    grid = std::max<int64_t>(1, std::min<int64_t>(grid, (ntiles + warps - 1) / warps));
    sb.warp_acc_bytes = sizeof(double) * (size_t)grid * warps * p.NE;
    sb.cta_part_bytes = sizeof(double) * (size_t)grid * p.NE;
    CU(cudaMalloc(&sb.warp_acc, sb.warp_acc_bytes));
    CU(cudaMalloc(&sb.cta_part, sb.cta_part_bytes));
    sb.grid = (int)grid;
    sb.threads = threads;
    sb.staged = 0;
    sb.stages = 0;
    sb.smem = 0;
    return RFB_OK;
  }
  bool all_slots = true;
  for (int s = 0; s < p.n_slots; ++s) all_slots = all_slots && s < ds.n_slots && ds.blocks[s] != nullptr;
  if (p.sub.ok && e->sweep_sub && e->sweep_staged && all_slots) {
    // subunit sweep (sweep_sub.cuh): 8-row tiles, rows stay in the stage through both passes
    const SubPlan& sp = p.sub;
    int variant = (e->sweep_sub == 6 || e->sweep_sub == 7) ? e->sweep_sub : std::min(std::max(e->sweep_sub, 1), 3);
    if (variant <= 2 && sp.K != 4) variant = 3;        // the comparison variants exist for 4 subunits only
    const int NJ = sub_template_nj(variant, sp.c0), cpl = sub_lpr(variant) * NJ;
    int st = e->sweep_stages > 0 ? e->sweep_stages : 2;
    int occ = 0;
    size_t smem = 0;
    bool fits = false;
    for (; st >= 2 && !fits; --st) {
      smem = sweep_sub_smem_bytes(warps, st, p.ctot, sp.K, cpl, sp.c0);
      fits = true;
      occ = 1 << 30;
      for (int bwd = 0; bwd < 2 && fits; ++bwd) {
        for (int spec = 0; spec < 2 && fits; ++spec) {
          SweepFn fn = sub_instance(variant, sp.K, NJ, bwd != 0, spec);
          if (!fn) return fail(RFB_ERR_UNSUPPORTED, "no subunit sweep instance");
          if (smem > 48 * 1024 &&
              cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            fits = false;
            break;
          }
          int o = 0;
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, threads, smem));
          occ = std::min(occ, o);
        }
      }
      if (fits && occ < 1) fits = false;
      if (fits) break;
    }
    if (fits) {
      if (e->sweep_ctas > 0) occ = std::min(occ, e->sweep_ctas);
      const int64_t ntiles = (ds.n + kSubRows - 1) / kSubRows;
      int64_t grid = (int64_t)e->num_sms * occ;
      const int64_t need = (ntiles + warps - 1) / warps;
      if (grid > need) grid = std::max<int64_t>(need, 1);
      sb.warp_acc_bytes = sizeof(double) * (size_t)grid * warps * p.NE;
      sb.cta_part_bytes = sizeof(double) * (size_t)grid * p.NE;
      CU(cudaMalloc(&sb.warp_acc, sb.warp_acc_bytes));
      CU(cudaMalloc(&sb.cta_part, sb.cta_part_bytes));
      sb.grid = (int)grid;
      sb.threads = threads;
      sb.staged = 1;
      sb.stages = st;
      sb.smem = smem;
      sb.sub = variant;
      sb.sub_nj = NJ;
      if (e->sweep_sub >= 4 && variant == 3) {
        // gradient sweeps: 15 (sweep_sub = 4) or 12 (= 5) consumer warps + 1 producer per SM, as many pooled
        // stages as shared memory holds
        const int ncw = e->sweep_sub == 5 ? 12 : 15;
        const int nthreads = sub2_threads(ncw);
        int ps = e->sweep_stages > 0 ? e->sweep_stages : 64;
        size_t psmem = 0;
        bool pfits = false;
        for (; ps >= ncw + 2 && !pfits; --ps) {
          psmem = sweep_sub2_smem_bytes(ps, p.ctot, sp.K, cpl, sp.c0);
          if (psmem > 227 * 1024) continue;
          pfits = true;
          for (int spec = 0; spec < 2 && pfits; ++spec) {
            SweepFn fn = pool_instance(sp.K, NJ, spec, ncw);
            if (!fn || cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem) != cudaSuccess) {
              cudaGetLastError();
              pfits = false;
              break;
            }
            int o = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, nthreads, psmem) != cudaSuccess || o < 1) {
              cudaGetLastError();
              pfits = false;
            }
          }
          if (pfits) break;
        }
        if (pfits) {
          int64_t pgrid = std::max<int64_t>(1, std::min<int64_t>(e->num_sms, (ntiles + ncw - 1) / ncw));
          if (e->sweep_pool_ctas > 0) pgrid = std::min<int64_t>(pgrid, e->sweep_pool_ctas);
          const size_t wbytes = sizeof(double) * (size_t)pgrid * ncw * p.NE;
          const size_t cbytes = sizeof(double) * (size_t)pgrid * p.NE;
          if (wbytes > sb.warp_acc_bytes) {
            cudaFree(sb.warp_acc);
            sb.warp_acc = nullptr;
            CU(cudaMalloc(&sb.warp_acc, wbytes));
            sb.warp_acc_bytes = wbytes;
          }
          if (cbytes > sb.cta_part_bytes) {
            cudaFree(sb.cta_part);
            sb.cta_part = nullptr;
            CU(cudaMalloc(&sb.cta_part, cbytes));
            sb.cta_part_bytes = cbytes;
          }
          sb.pool = true;
          sb.pool_ncw = ncw;
          sb.pool_grid = (int)pgrid;
          sb.pool_stages = ps;
          sb.pool_smem = psmem;
        }
      }
      return RFB_OK;
    }
  }
  int staged = e->sweep_staged ? 1 : 0;
  int stages = e->sweep_stages;
  if (stages <= 0) {
    // Measured on B200 (tools/tune_sweep.py, 2^25 x 296): bandwidth peaks with ~112 KB of rows in
    // flight per SM (4 warps x 3 stages x 9.5 KB); both fewer and more bytes in flight are slower.
    const int ctas = e->sweep_ctas > 0 ? e->sweep_ctas : (p.Ht == 1 ? 1 : 2);
    const double per_stage = (double)warps * ctas * sweep_stage_bytes(p.RB, p.ctot);
    stages = (int)std::lround(114688.0 / per_stage);
    stages = std::min(std::max(stages, 2), 8);
  }
  int occ = 0;
  size_t smem = 0;
  for (;;) {
    smem = staged ? sweep_smem_bytes(warps, stages, p.RB, p.ctot) : 0;
    occ = 0;
    bool ok = true;
    for (int bwd = 0; bwd < 2 && ok; ++bwd) {
      for (int spec = 0; spec < kNumSpecs && ok; ++spec) {
      SweepFn fn = sweep_instance(p.Ht, p.NCH, bwd != 0, staged != 0, spec, &RB);
      if (!fn) return fail(RFB_ERR_UNSUPPORTED, "no sweep kernel instance");
      if (smem > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) {
          cudaGetLastError();
          ok = false;
          break;
        }
      }
      int o = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, threads, smem));
      if (bwd) occ = (spec == 0) ? o : std::min(occ, o);   // the gradient sweeps decide the launch shape
      }
    }
    if (ok && occ >= 1) break;
    if (staged && stages > 2) { --stages; continue; }       // shallower ring
    if (staged) { staged = 0; continue; }                   // rows too wide to stage: stream directly
    return fail(RFB_ERR_CUDA, "sweep kernel does not fit on an SM");
  }
  const int want_ctas = e->sweep_ctas >= 0 ? e->sweep_ctas : (p.Ht == 1 ? 1 : 0);
  if (want_ctas > 0) occ = std::min(occ, want_ctas);
  const int64_t ntiles = (ds.n + RB - 1) / RB;
  int64_t grid = (int64_t)e->num_sms * occ;
  const int64_t need = (ntiles + warps - 1) / warps;
  if (grid > need) grid = std::max<int64_t>(need, 1);
  sb.warp_acc_bytes = sizeof(double) * (size_t)grid * warps * p.NE;
  sb.cta_part_bytes = sizeof(double) * (size_t)grid * p.NE;
  CU(cudaMalloc(&sb.warp_acc, sb.warp_acc_bytes));
  CU(cudaMalloc(&sb.cta_part, sb.cta_part_bytes));
  sb.grid = (int)grid;
  sb.threads = threads;
  sb.staged = staged;
  sb.stages = stages;
  sb.smem = smem;
  return RFB_OK;
}

// Launch one sweep over data set `kind` with the parameter image `ps`.  Leaves the per-CTA
// partial sums in m->sb[kind].cta_part (grid rows).
// The kernel-independent part of a sweep's arguments: where the rows, the response and the parameters are.
static SweepArgs make_sweep_args(const rfb_model* m, int kind, const ParamSet& ps, const int* head_nl,
                                 int out_nl_override, const float* softmax) {
  const Plan& p = m->plan;
  const DataSet& ds = m->data[kind];
  SweepArgs a{};
  a.n_slots = p.n_slots;
  for (int s = 0; s < p.n_slots; ++s) {
    rfb_block* b = s < ds.n_slots ? ds.blocks[s] : nullptr;
    a.slot_ptr[s] = b ? b->d : nullptr;
    a.slot_ld[s] = b ? b->ld : 0;
  }
  for (int s = 0; s <= p.n_slots; ++s) a.slot_chunk0[s] = p.slot_chunk0[s];
  a.n_chunks = p.n_chunks;
  a.n_rows = ds.n;
  a.y = ds.has_y ? ds.y : nullptr;
  a.W = ps.W;
  a.head_icpt = ps.head_icpt;
  a.glob_icpt = ps.glob_icpt;
  for (int h = 0; h < kMaxHeads; ++h) {
    a.head_nl[h] = h < p.Ht ? head_nl[h] : RFB_NL_NONE;
    a.head_slotmask[h] = h < p.Ht ? p.head_slotmask[h] : 0u;
  }
  a.out_nl = out_nl_override >= 0 ? out_nl_override : m->out_nl;
  a.softmax = softmax;
  a.distr = m->distr;
  a.two_over_n = ds.n_total > 0 ? (float)(2.0 / ds.n_total) : 0.f;
  a.layout_heads = p.Ht;
  a.sub_slot = p.sub.slot;
  for (int k = 0; k < kSubMaxK; ++k) a.sub_head[k] = p.sub.heads[k];
  a.sub_extra_head = p.sub.extra_head;
  return a;
}

static int launch_sweep(rfb_model* m, int kind, const ParamSet& ps, const int* head_nl, bool bwd,
                        float* r_out, int out_nl_override = -1, const float* softmax = nullptr, int prep = 0) {
  const Plan& p = m->plan;
  rfb_engine* e = m->e;
  int RB = p.RB;
  TRY(ensure_sweep_buffers(m, kind));
  SweepBuffers& sb = m->sb[kind];
  SweepFn fn = nullptr;
  const int eff_out_nl = out_nl_override >= 0 ? out_nl_override : m->out_nl;
  if (sb.sub) {
    int spec = (m->distr == RFB_DISTR_POISSON && eff_out_nl == RFB_NL_SOFTPLUS) ? 1 : 0;
    for (int k = 0; k < p.sub.K; ++k)
      if (head_nl[p.sub.heads[k]] != RFB_NL_SOFTPLUS) spec = 0;
    fn = sub_instance(sb.sub, p.sub.K, sb.sub_nj, bwd, spec);
    if (!fn) return fail(RFB_ERR_UNSUPPORTED, "no subunit sweep instance");
  } else if (!p.generic) {
    fn = sweep_instance(p.Ht, p.NCH, bwd, sb.staged != 0, sweep_spec(m, head_nl), &RB);
    if (!fn) return fail(RFB_ERR_UNSUPPORTED, "no sweep kernel instance");
  }

  SweepArgs a = make_sweep_args(m, kind, ps, head_nl, out_nl_override, softmax);
  a.warp_acc = sb.warp_acc;
  a.cta_part = sb.cta_part;
  a.r_out = r_out;
  a.head_scale = p.n_smx ? m->smx_scale[kind] : nullptr;
  a.prep = prep;
  if (p.n_smx && !a.head_scale) return fail(RFB_ERR_STATE, "softmax filter heads without their normalisers");
  a.flush_tiles = std::max(1, 256 / RB);
  a.n_stages = sb.stages;
  if (m->fit.stamp_now && m->fit.stamps && kind <= RFB_KIND_DEV) {
    a.stamp = m->fit.stamps + (size_t)kind * 2 * ((size_t)m->fit.max_iters + 1);
    a.stamp_it = m->fit.it;
  }
  sb.cur_grid = sb.grid;
  if (sb.sub && sb.pool && bwd) {
    int spec = (m->distr == RFB_DISTR_POISSON && eff_out_nl == RFB_NL_SOFTPLUS) ? 1 : 0;
    for (int k = 0; k < p.sub.K; ++k)
      if (head_nl[p.sub.heads[k]] != RFB_NL_SOFTPLUS) spec = 0;
    SweepFn pfn = pool_instance(p.sub.K, sb.sub_nj, spec, sb.pool_ncw);
    if (!pfn) return fail(RFB_ERR_UNSUPPORTED, "no pooled subunit sweep instance");
    a.n_stages = sb.pool_stages;
    pfn<<<sb.pool_grid, sub2_threads(sb.pool_ncw), sb.pool_smem, e->stream>>>(a);
    sb.cur_grid = sb.pool_grid;
  } else if (sb.sub) {
    fn<<<sb.grid, sb.threads, sb.smem, e->stream>>>(a);
  } else if (p.generic) {
    if (bwd) sweep_generic_kernel<true><<<sb.grid, sb.threads, 0, e->stream>>>(a, p.Ht);
    else sweep_generic_kernel<false><<<sb.grid, sb.threads, 0, e->stream>>>(a, p.Ht);
  } else {
    fn<<<sb.grid, sb.threads, sb.smem, e->stream>>>(a);
  }
  e->launches++;
  CU(cudaGetLastError());
  return RFB_OK;
}

// Reduce per-CTA partial rows in a fixed order.  Segment 0: entries [e0, e0 + n0) of the
// partial rows of `kind0` -> out0; optional segment 1 likewise (kind1 < 0: none).
static int launch_reduce(rfb_model* m, int kind0, int e0, int n0, double* out0, int kind1, int e1,
                         int n1, double* out1) {
  const Plan& p = m->plan;
  ReduceSeg s0{m->sb[kind0].cta_part + e0, m->sb[kind0].cur_grid, p.NE, n0, out0};
  const int blocks0 = (n0 + kReduceEx - 1) / kReduceEx;
  ReduceSeg s1{nullptr, 0, 0, 0, nullptr};
  int blocks1 = 0;
  if (kind1 >= 0) {
    s1 = ReduceSeg{m->sb[kind1].cta_part + e1, m->sb[kind1].cur_grid, p.NE, n1, out1};
    blocks1 = (n1 + kReduceEx - 1) / kReduceEx;
  }
  reduce_partials_kernel<<<blocks0 + blocks1, dim3(kReduceEx, kReduceEy), 0, m->e->stream>>>(s0, s1, blocks0);
  m->e->launches++;
  CU(cudaGetLastError());
  return RFB_OK;
}

// ---- output softmax (see softmax.cuh) ------------------------------------------------------------
static int allreduce_device_f64(rfb_engine* e, double* d, size_t count);
// Scratch of the softmax evaluation of data set `kind`; allocation is not capturable, so the fit
// calls this before it captures the iteration.
static int ensure_softmax_buffers(rfb_model* m, int kind) {
  DataSet& ds = m->data[kind];
  if (!ds.r_buf) CU(cudaMalloc(&ds.r_buf, sizeof(float) * r_buf_floats(ds.n)));
  if (!m->sm_scal[kind]) {
    CU(cudaMalloc(&m->sm_scal[kind], sizeof(double) * 3));
    CU(cudaMalloc(&m->sm_part[kind], sizeof(double) * 2 * kSoftmaxBlocks));
    CU(cudaMalloc(&m->sm_f[kind], sizeof(float) * 2));
  }
  return RFB_OK;
}

// Enqueue steps 1-2 of the softmax evaluation at parameters `ps`; afterwards m->sm_f[kind] holds
// {1 / Z, D} for the final sweep.  `want_d`: the gradient follows (needs the response).
static int softmax_prepare(rfb_model* m, int kind, const ParamSet& ps, const int* head_nl, bool want_d) {
  DataSet& ds = m->data[kind];
  rfb_engine* e = m->e;
  const Plan& p = m->plan;
  double* Z = m->sm_scal[kind];
  double* D = m->sm_scal[kind] + 1;
  TRY(launch_sweep(m, kind, ps, head_nl, false, ds.r_buf, RFB_NL_EXPONENTIAL));
  TRY(launch_reduce(m, kind, p.Ht * p.ctot + S_R, 1, Z, -1, 0, 0, nullptr));
  TRY(allreduce_device_f64(e, Z, 1));
  if (want_d) {
    softmax_stats_kernel<<<kSoftmaxBlocks, 256, 0, e->stream>>>(
        ds.r_buf, ds.y, ds.n, Z, m->distr, ds.n_total > 0 ? (float)(2.0 / ds.n_total) : 0.f, m->sm_part[kind]);
    softmax_sum_kernel<<<1, 32, 0, e->stream>>>(m->sm_part[kind], kSoftmaxBlocks, D);
    e->launches += 2;
    CU(cudaGetLastError());
    TRY(allreduce_device_f64(e, D, 2));
  }
  softmax_finish_kernel<<<1, 32, 0, e->stream>>>(Z, want_d ? D : nullptr, m->sm_f[kind]);
  e->launches++;
  CU(cudaGetLastError());
  return RFB_OK;
}

// ---- 'softmax' as a filter nonlinearity (head_softmax_fix in sweep.cuh, smx_*_kernel in softmax.cuh) ---------
static SmxHeads smx_heads(const Plan& p, const int* head_nl) {
  SmxHeads hs{};
  hs.n_heads = p.Ht;
  for (int h = 0; h < p.Ht; ++h) hs.is_smx[h] = head_nl[h] == RFB_NL_SOFTMAX;
  return hs;
}
// not capturable (allocation): the fit calls this before it captures the iteration
static int ensure_smx_buffers(rfb_model* m, int kind) {
  const Plan& p = m->plan;
  if (!p.n_smx) return RFB_OK;
  if (m->smx_fin[kind] && m->smx_fin_n[kind] != p.NE) {
    CU(cudaFree(m->smx_fin[kind]));
    m->smx_fin[kind] = nullptr;
  }
  if (!m->smx_fin[kind]) {
    CU(cudaMalloc(&m->smx_fin[kind], sizeof(double) * p.NE));
    m->smx_fin_n[kind] = p.NE;
  }
  if (!m->smx_scale[kind]) CU(cudaMalloc(&m->smx_scale[kind], sizeof(float) * kMaxHeads));
  return RFB_OK;
}
// Enqueue the preparatory sweep at parameters `ps`: afterwards m->smx_scale[kind] holds 1 / Z_h and (want_grad)
// m->smx_fin[kind] the vectors E_h that smx_fix needs.
static int smx_prepare(rfb_model* m, int kind, const ParamSet& ps, const int* head_nl, bool want_grad) {
  const Plan& p = m->plan;
  if (!p.n_smx) return RFB_OK;
  bool any = false;
  for (int h = 0; h < p.Ht; ++h) any = any || head_nl[h] == RFB_NL_SOFTMAX;
  const int tail = p.Ht * p.ctot;
  if (any) {
    // (the output stage is irrelevant here -- delta is forced to 1 -- so it runs as the identity)
    TRY(launch_sweep(m, kind, ps, head_nl, want_grad, nullptr, RFB_NL_NONE, nullptr, 1));
    const int lo = want_grad ? 0 : tail;
    TRY(launch_reduce(m, kind, lo, p.NE - lo, m->smx_fin[kind] + lo, -1, 0, 0, nullptr));
    TRY(allreduce_device_f64(m->e, m->smx_fin[kind] + lo, (size_t)(p.NE - lo)));
  }
  smx_finish_kernel<<<1, 32, 0, m->e->stream>>>(m->smx_fin[kind], tail, smx_heads(p, head_nl), m->smx_scale[kind]);
  m->e->launches++;
  CU(cudaGetLastError());
  return RFB_OK;
}
// after the ordinary sweep's sums are complete (all ranks): the cross-sample term of the softmax Jacobian
static int smx_fix(rfb_model* m, int kind, const int* head_nl, double* fin) {
  const Plan& p = m->plan;
  if (!p.n_smx) return RFB_OK;
  smx_fix_kernel<<<1, 256, 0, m->e->stream>>>(fin, m->smx_fin[kind], p.ctot, p.Ht * p.ctot, smx_heads(p, head_nl));
  m->e->launches++;
  CU(cudaGetLastError());
  return RFB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: misc / engine
// ------------------------------------------------------------------------------------------------
extern "C" const char* rfb_last_error(void) { return g_err.c_str(); }
extern "C" int rfb_abi_version(void) { return RFB_ABI_VERSION; }

extern "C" int rfb_engine_create(int device, rfb_engine** out) {
  if (!out) return fail(RFB_ERR_INVALID, "out is NULL");
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count)
    return fail(RFB_ERR_INVALID, "device %d out of range (%d visible)", device, count);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(RFB_ERR_UNSUPPORTED, "rfest_b200 is built for sm_100a only; device %d is sm_%d%d",
                device, prop.major, prop.minor);
  rfb_engine* e = new rfb_engine();
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  if (const char* v = getenv("RFB_SWEEP_STAGED")) e->sweep_staged = atoi(v);
  if (const char* v = getenv("RFB_SWEEP_STAGES")) e->sweep_stages = atoi(v);
  if (const char* v = getenv("RFB_SWEEP_WARPS")) e->sweep_warps = atoi(v);
  if (const char* v = getenv("RFB_SWEEP_CTAS")) e->sweep_ctas = atoi(v);
  if (const char* v = getenv("RFB_SWEEP_SUB")) e->sweep_sub = atoi(v);
  if (const char* v = getenv("RFB_PROJECT_TC")) e->project_tc = atoi(v);
  if (const char* v = getenv("RFB_XCHG")) e->use_xchg = atoi(v);
  if (const char* v = getenv("RFB_H2D_THREADS")) e->h2d_threads = std::max(1, atoi(v));
  if (const char* v = getenv("RFB_H2D_OVERLAP")) e->h2d_overlap = atoi(v);
  if (const char* v = getenv("RFB_FIT_PERSISTENT")) e->fit_persistent = atoi(v);
  if (const char* v = getenv("RFB_GRAM_MMA")) e->gram_mma = atoi(v);
  if (const char* v = getenv("RFB_PROJECT_FUSED")) e->project_fused = atoi(v);
  if (const char* v = getenv("RFB_FIT_SMALL_CTAS")) e->fit_small_ctas = std::max(0, atoi(v));
  *out = e;
  return RFB_OK;
}

extern "C" int rfb_engine_set_option(rfb_engine* e, const char* key, int value) {
  if (!e || !key) return fail(RFB_ERR_INVALID, "NULL argument");
  const std::string k(key);
  if (k == "sweep_staged") e->sweep_staged = value;
  else if (k == "sweep_stages") e->sweep_stages = value;
  else if (k == "sweep_warps") e->sweep_warps = value;
  else if (k == "sweep_ctas") e->sweep_ctas = value;
  else if (k == "sweep_sub") e->sweep_sub = value;
  else if (k == "project_tc") e->project_tc = value;
  else if (k == "xchg") e->use_xchg = value;
  else if (k == "h2d_threads") e->h2d_threads = std::max(1, value);
  else if (k == "h2d_overlap") e->h2d_overlap = value;
  else if (k == "fit_persistent") e->fit_persistent = value;
  else if (k == "gram_mma") e->gram_mma = value;
  else if (k == "project_fused") e->project_fused = value;
  else if (k == "sweep_pool_ctas") e->sweep_pool_ctas = std::max(0, value);
  else if (k == "fit_small_ctas") e->fit_small_ctas = std::max(0, value);
  else if (k == "project_slab_rows") e->project_slab_rows = std::max(0, value);
  else return fail(RFB_ERR_INVALID, "unknown option '%s'", key);
  return RFB_OK;
}

extern "C" int rfb_engine_destroy(rfb_engine* e) {
  if (!e) return RFB_OK;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (Scratch* s : {&e->stage, &e->zbuf, &e->stbuf, &e->sxybuf, &e->small, &e->bsplit, &e->gram_part}) cudaFree(s->p);
  for (int k = 0; k < rfb_engine::kRingSlots; ++k) {
    if (e->ring_pinned[k]) cudaFreeHost(e->ring_pinned[k]);
    if (e->ring_dev[k]) cudaFree(e->ring_dev[k]);
    if (e->ring_copied[k]) cudaEventDestroy(e->ring_copied[k]);
    if (e->ring_consumed[k]) cudaEventDestroy(e->ring_consumed[k]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  cudaStreamDestroy(e->stream);
  delete e;
  return RFB_OK;
}

extern "C" int rfb_engine_synchronize(rfb_engine* e) {
  if (!e) return fail(RFB_ERR_INVALID, "engine is NULL");
  TRY(use_device(e));
  CU(cudaStreamSynchronize(e->stream));
  return RFB_OK;
}

extern "C" int rfb_engine_launch_count(rfb_engine* e, int64_t* out) {
  if (!e || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  *out = e->launches;
  return RFB_OK;
}

extern "C" int rfb_engine_stream(rfb_engine* e, uint64_t* out) {
  if (!e || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  *out = (uint64_t)(uintptr_t)e->stream;
  return RFB_OK;
}

extern "C" int rfb_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(RFB_ERR_INVALID, "out is NULL");
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return RFB_OK;
}

extern "C" int rfb_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return RFB_OK;
}

extern "C" int rfb_comm_unique_id(void* out128) {
  if (!out128) return fail(RFB_ERR_INVALID, "out128 is NULL");
  TRY(nccl_load());
  NcclUid id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return RFB_OK;
}

extern "C" int rfb_engine_comm_init(rfb_engine* e, const void* id128, int rank, int world) {
  if (!e || !id128) return fail(RFB_ERR_INVALID, "NULL argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(RFB_ERR_INVALID, "bad rank/world");
  if (e->comm) return fail(RFB_ERR_STATE, "communicator already initialised");
  if (world == 1) {
    e->rank = 0;
    e->world = 1;
    return RFB_OK;
  }
  TRY(nccl_load());
  TRY(use_device(e));
  NcclUid id;
  memcpy(&id, id128, sizeof(id));
  NC(g_nccl.CommInitRank(&e->comm, world, id, rank));
  e->rank = rank;
  e->world = world;
  return RFB_OK;
}

extern "C" int rfb_engine_comm_info(rfb_engine* e, int* rank, int* world) {
  if (!e) return fail(RFB_ERR_INVALID, "engine is NULL");
  if (rank) *rank = e->rank;
  if (world) *world = e->world;
  return RFB_OK;
}

static size_t xchg_bytes(const rfb_engine* e) {
  return (size_t)kXchgDataOffset + sizeof(double) * 2 * kMaxRanks * (size_t)e->xchg_cap;
}

extern "C" int rfb_engine_xchg_create(rfb_engine* e, void* handle64) {
  if (!e || !handle64) return fail(RFB_ERR_INVALID, "NULL argument");
  TRY(use_device(e));
  if (!e->xchg_local) {
    CU(cudaMalloc(&e->xchg_local, xchg_bytes(e)));
    CU(cudaMemset(e->xchg_local, 0, xchg_bytes(e)));
    CU(cudaMalloc(&e->xchg_seq, sizeof(unsigned long long)));
    CU(cudaMalloc(&e->xchg_done, sizeof(unsigned int)));
    CU(cudaMemset(e->xchg_seq, 0, sizeof(unsigned long long)));
    CU(cudaMemset(e->xchg_done, 0, sizeof(unsigned int)));
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, e->xchg_local));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, sizeof(h));
  return RFB_OK;
}

extern "C" int rfb_engine_xchg_connect(rfb_engine* e, const void* handles) {
  if (!e || !handles) return fail(RFB_ERR_INVALID, "NULL argument");
  if (!e->xchg_local) return fail(RFB_ERR_STATE, "rfb_engine_xchg_create has not been called");
  if (e->world < 2 || e->world > kMaxRanks) return fail(RFB_ERR_UNSUPPORTED, "exchange needs 2..%d ranks", kMaxRanks);
  TRY(use_device(e));
  for (int q = 0; q < e->world; ++q) {
    if (q == e->rank) {
      e->xchg_peer[q] = e->xchg_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * q, sizeof(h));
    void* p = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      return fail(RFB_ERR_CUDA, "cannot map rank %d's exchange buffer: %s", q, cudaGetErrorString(ce));
    }
    e->xchg_peer[q] = (unsigned char*)p;
  }
  e->xchg_ready = true;
  return RFB_OK;
}

static XchgArgs xchg_args(const rfb_engine* e, bool on) {
  XchgArgs x{};
  x.world = 1;
  if (on) {
    for (int q = 0; q < e->world; ++q) x.peer[q] = e->xchg_peer[q];
    x.world = e->world;
    x.rank = e->rank;
    x.cap = e->xchg_cap;
    x.seq = e->xchg_seq;
    x.done = e->xchg_done;
  }
  return x;
}

static int allreduce_device_f64(rfb_engine* e, double* d, size_t count) {
  if (!e->comm) return RFB_OK;
  NC(g_nccl.AllReduce(d, d, count, kNcclFloat64, kNcclSum, e->comm, e->stream));
  return RFB_OK;
}

extern "C" int rfb_engine_allreduce_f64(rfb_engine* e, double* host, int count) {
  if (!e || !host || count < 0) return fail(RFB_ERR_INVALID, "bad argument");
  if (!e->comm || count == 0) return RFB_OK;
  TRY(use_device(e));
  TRY(scratch_reserve(e, e->small, sizeof(double) * count));
  double* d = (double*)e->small.p;
  CU(cudaMemcpyAsync(d, host, sizeof(double) * count, cudaMemcpyHostToDevice, e->stream));
  TRY(allreduce_device_f64(e, d, count));
  CU(cudaMemcpyAsync(host, d, sizeof(double) * count, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return RFB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: blocks and projection
// ------------------------------------------------------------------------------------------------
extern "C" int rfb_block_create(rfb_engine* e, int64_t n_rows, int n_cols, rfb_block** out) {
  if (!e || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  if (n_rows <= 0 || n_cols <= 0) return fail(RFB_ERR_INVALID, "empty block %lld x %d", (long long)n_rows, n_cols);
  TRY(use_device(e));
  rfb_block* b = new rfb_block();
  b->e = e;
  b->n_rows = n_rows;
  b->n_cols = n_cols;
  b->ld = ((int64_t)n_cols + 3) / 4 * 4;
  const size_t bytes = sizeof(float) * (size_t)b->ld * (size_t)n_rows;
  cudaError_t err = cudaMalloc(&b->d, bytes);
  if (err != cudaSuccess) {
    delete b;
    return fail(RFB_ERR_CUDA, "cudaMalloc of a %lld x %d block (%.2f GB) failed: %s",
                (long long)n_rows, n_cols, bytes / 1e9, cudaGetErrorString(err));
  }
  CU(cudaMemsetAsync(b->d, 0, bytes, e->stream));
  *out = b;
  return RFB_OK;
}

extern "C" int rfb_block_destroy(rfb_block* b) {
  if (!b) return RFB_OK;
  cudaSetDevice(b->e->device);
  cudaStreamSynchronize(b->e->stream);
  cudaFree(b->d);
  delete b;
  return RFB_OK;
}

extern "C" int rfb_block_shape(const rfb_block* b, int64_t* n_rows, int* n_cols, int64_t* ld) {
  if (!b) return fail(RFB_ERR_INVALID, "block is NULL");
  if (n_rows) *n_rows = b->n_rows;
  if (n_cols) *n_cols = b->n_cols;
  if (ld) *ld = b->ld;
  return RFB_OK;
}

extern "C" int rfb_block_device_ptr(const rfb_block* b, uint64_t* out) {
  if (!b || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  *out = (uint64_t)(uintptr_t)b->d;
  return RFB_OK;
}

extern "C" int rfb_block_read(rfb_block* b, int64_t row0, int64_t n, float* host_out) {
  if (!b || !host_out) return fail(RFB_ERR_INVALID, "NULL argument");
  if (row0 < 0 || n < 0 || row0 + n > b->n_rows) return fail(RFB_ERR_INVALID, "row range out of bounds");
  if (n == 0) return RFB_OK;
  TRY(use_device(b->e));
  CU(cudaMemcpy2DAsync(host_out, sizeof(float) * b->n_cols, b->d + row0 * b->ld, sizeof(float) * b->ld,
                       sizeof(float) * b->n_cols, (size_t)n, cudaMemcpyDeviceToHost, b->e->stream));
  CU(cudaStreamSynchronize(b->e->stream));
  return RFB_OK;
}

extern "C" int rfb_block_write(rfb_block* b, int64_t row0, int64_t n, const float* src, int on_device) {
  if (!b || !src) return fail(RFB_ERR_INVALID, "NULL argument");
  if (row0 < 0 || n < 0 || row0 + n > b->n_rows) return fail(RFB_ERR_INVALID, "row range out of bounds");
  if (n == 0) return RFB_OK;
  TRY(use_device(b->e));
  CU(cudaMemcpy2DAsync(b->d + row0 * b->ld, sizeof(float) * b->ld, src, sizeof(float) * b->n_cols,
                       sizeof(float) * b->n_cols, (size_t)n,
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, b->e->stream));
  CU(cudaStreamSynchronize(b->e->stream));
  return RFB_OK;
}

// ---- tensor-core spatial projection (project_tc.cuh) ---------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// fp32 row-major [rows x cols] (row pitch `pitch_floats`), box = 32 columns x `box_rows` rows, 128-byte swizzle
static bool make_map_2d(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t pitch_floats,
                        uint32_t box_rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch_floats * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kTcK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TcPlan {
  bool ok = false;
  int n_pad = 0, k_tiles = 0, stages = 0, tmem_cols = 0, drain = 4, n_raw = 0;
  size_t smem = 0;
  // fused spatial + temporal variant (Z ring in shared memory): fewer raw A slots fit
  bool fused_ok = false;
  int fused_n_raw = 0;
  size_t fused_smem = 0;
  const float *bhi = nullptr, *blo = nullptr;   // device, [n_pad x k_tiles*32], K contiguous
  CUtensorMap map_bhi, map_blo;
};

// Decide whether Z = stim @ Sxy can run on the tensor cores and stage the split, transposed B.
static int tc_prepare(rfb_engine* e, const float* Sxy, int n_pix, int n_q, TcPlan* tp) {
  tp->ok = false;
  if (!e->project_tc || !Sxy) return RFB_OK;
  if (n_pix < 64 || (n_pix % 4) != 0 || n_q < 8 || n_q > 128) return RFB_OK;
  if (!tensor_map_encoder()) return RFB_OK;
  tp->n_pad = (n_q + 15) / 16 * 16;
  tp->k_tiles = (n_pix + kTcK - 1) / kTcK;
  const int kpad = tp->k_tiles * kTcK;
  // TMEM: two accumulators (drained into registers every `drain` K tiles so that no accumulation
  // chain exceeds 12 * drain truncating adds) + two A stages of 32 hi + 32 lo columns
  tp->drain = 2;
  tp->tmem_cols = 32;
  while (tp->tmem_cols < 2 * tp->n_pad + 64 * kTcAStages) tp->tmem_cols *= 2;
  if (tp->tmem_cols > 512) return RFB_OK;
  // shared memory: B stages + as many raw A slots (HBM bytes in flight) as fit
  tp->stages = 3;
  const size_t budget = 225 * 1024;
  const size_t fixed = 1024 + 512 + (size_t)tp->stages * tc_stage_bytes(tp->n_pad);
  if (fixed + 2 * kTcRawBytes > budget) return RFB_OK;
  tp->n_raw = (int)std::min<size_t>(8, (budget - fixed) / kTcRawBytes);
  if (const char* v = getenv("RFB_TC_NRAW")) tp->n_raw = std::max(2, std::min(tp->n_raw, atoi(v)));
  if (const char* v = getenv("RFB_TC_DRAIN")) tp->drain = std::max(1, atoi(v));
  if (const char* v = getenv("RFB_TC_STAGES")) tp->stages = std::max(2, std::min(tp->stages, atoi(v)));
  tp->smem = tc_smem_bytes(tp->n_pad, tp->stages, tp->n_raw);
  // hi = tf32 truncation of b, lo = b - hi (exact); transposed so that K is contiguous (K-major B)
  std::vector<float> h((size_t)2 * tp->n_pad * kpad, 0.f);
  float* hi = h.data();
  float* lo = h.data() + (size_t)tp->n_pad * kpad;
  for (int k = 0; k < n_pix; ++k)
    for (int q = 0; q < n_q; ++q) {
      const float b = Sxy[(size_t)k * n_q + q];
      uint32_t u;
      memcpy(&u, &b, 4);
      u &= 0xFFFFE000u;
      float bh;
      memcpy(&bh, &u, 4);
      hi[(size_t)q * kpad + k] = bh;
      lo[(size_t)q * kpad + k] = b - bh;
    }
  TRY(scratch_reserve(e, e->bsplit, sizeof(float) * h.size()));
  CU(cudaMemcpyAsync(e->bsplit.p, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  tp->bhi = (const float*)e->bsplit.p;
  tp->blo = tp->bhi + (size_t)tp->n_pad * kpad;
  if (!make_map_2d(&tp->map_bhi, tp->bhi, tp->n_pad, kpad, kpad, tp->n_pad)) return RFB_OK;
  if (!make_map_2d(&tp->map_blo, tp->blo, tp->n_pad, kpad, kpad, tp->n_pad)) return RFB_OK;
  if (cudaFuncSetAttribute(spatial_project_tc_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)tp->smem) != cudaSuccess) {
    cudaGetLastError();
    return RFB_OK;
  }
  tp->ok = true;
  if (e->project_fused && n_q <= kTcFusedMaxN) {
    const size_t ring = tc_ring_bytes(n_q);
    if (fixed + ring + 3 * kTcRawBytes <= budget) {
      tp->fused_n_raw = (int)std::min<size_t>(8, (budget - fixed - ring) / kTcRawBytes);
      if (const char* v = getenv("RFB_TC_NRAW")) tp->fused_n_raw = std::max(2, std::min(tp->fused_n_raw, atoi(v)));
      tp->fused_smem = tc_smem_bytes(tp->n_pad, tp->stages, tp->fused_n_raw, ring);
      tp->fused_ok =
          cudaFuncSetAttribute(spatial_project_tc_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)tp->fused_smem) == cudaSuccess &&
          cudaFuncSetAttribute(spatial_project_tc_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)tp->fused_smem) == cudaSuccess;
      if (!tp->fused_ok) cudaGetLastError();
    }
  }
  return RFB_OK;
}

// returns false if this chunk cannot use the tensor-core path (misaligned pointer)
static bool tc_launch(rfb_engine* e, const TcPlan& tp, const float* d_stim, int64_t nz, int n_pix, int n_q,
                      float* Z) {
  if (((uintptr_t)d_stim & 15u) != 0) return false;
  CUtensorMap map_a;
  if (!make_map_2d(&map_a, d_stim, (uint64_t)nz, (uint64_t)n_pix, (uint64_t)n_pix, kTcM)) return false;
  TcArgs a{};
  a.n_rows = nz;
  a.n_k_tiles = tp.k_tiles;
  a.n = n_q;
  a.n_pad = tp.n_pad;
  a.n_stages = tp.stages;
  a.Z = Z;
  a.ldz = n_q;
  a.tmem_cols = tp.tmem_cols;
  a.drain = tp.drain;
  a.products = getenv("RFB_TC_PRODUCTS") ? std::max(1, std::min(3, atoi(getenv("RFB_TC_PRODUCTS")))) : 3;
  a.n_raw = tp.n_raw;
  const int64_t tiles = (nz + kTcM - 1) / kTcM;
  const int grid = (int)std::min<int64_t>(tiles, e->num_sms);
  spatial_project_tc_kernel<false, 8><<<grid, kTcThreads, tp.smem, e->stream>>>(map_a, tp.map_bhi, tp.map_blo, a,
                                                                                TcFuseArgs{});
  e->launches++;
  return true;
}

// Spatial + temporal contraction of one slab in ONE kernel (project_tc.cuh, FUSED): returns false if the slab
// cannot take this path.  `off`: Z row feeding (design row 0 of the slab, lag 0), <= 0.
static bool tc_launch_fused(rfb_engine* e, const TcPlan& tp, const float* d_stim, int64_t nz, int n_pix, int n_q,
                            int DFT, const TemporalArgs& t, int64_t off) {
  if (!tp.fused_ok || ((uintptr_t)d_stim & 15u) != 0 || (DFT != 4 && DFT != 8)) return false;
  if (off > 0 || off < -(int64_t)(t.nlag - 1) || t.nlag > kTempMaxLag || nz < 1) return false;
  CUtensorMap map_a;
  if (!make_map_2d(&map_a, d_stim, (uint64_t)nz, (uint64_t)n_pix, (uint64_t)n_pix, kTcM)) return false;
  TcArgs a{};
  a.n_rows = nz;
  a.n_k_tiles = tp.k_tiles;
  a.n = n_q;
  a.n_pad = tp.n_pad;
  a.n_stages = tp.stages;
  a.Z = nullptr;
  a.ldz = n_q;
  a.tmem_cols = tp.tmem_cols;
  a.drain = tp.drain;
  a.products = getenv("RFB_TC_PRODUCTS") ? std::max(1, std::min(3, atoi(getenv("RFB_TC_PRODUCTS")))) : 3;
  a.n_raw = tp.fused_n_raw;
  TcFuseArgs f{};
  f.off = off;
  f.n_out = t.n_out;
  f.out = t.out;
  f.out_ld = t.out_ld;
  f.out_row0 = t.out_row0;
  f.col0 = t.col0;
  f.nlag = t.nlag;
  f.df_t = t.df_t;
  f.n_tiles = (off + t.n_out + t.nlag - 1 + kTcM - 1) / kTcM;
  if (f.n_tiles < 1) return false;
  const int grid = (int)std::min<int64_t>(f.n_tiles, e->num_sms);
  if (DFT == 8)
    spatial_project_tc_kernel<true, 8><<<grid, kTcFusedThreads, tp.fused_smem, e->stream>>>(map_a, tp.map_bhi, tp.map_blo, a, f);
  else
    spatial_project_tc_kernel<true, 4><<<grid, kTcFusedThreads, tp.fused_smem, e->stream>>>(map_a, tp.map_bhi, tp.map_blo, a, f);
  e->launches++;
  return true;
}

template <int DFT>
static void launch_temporal(const TemporalArgs& a, cudaStream_t s, bool use_const) {
  const int threads = 128;
  if (use_const) {   // temporal basis in constant memory, register window over the lags
    const long long items = ((a.n_out + 7) / 8) * a.n_q;
    temporal_project_const_kernel<DFT><<<(unsigned)((items + threads - 1) / threads), threads, 0, s>>>(a);
    return;
  }
  constexpr int TT = DFT <= 8 ? 8 : 4;
  const long long groups = (a.n_out + TT - 1) / TT;
  const long long items = groups * a.n_q;
  const long long blocks = (items + threads - 1) / threads;
  temporal_project_kernel<DFT, TT><<<(unsigned)blocks, threads, 0, s>>>(a);
}

extern "C" int rfb_block_matmul(rfb_block* b, const float* B, int k, float* out) {
  if (!b || !B || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  if (k < 1) return fail(RFB_ERR_INVALID, "k must be >= 1");
  if (b->n_rows == 0) return RFB_OK;
  rfb_engine* e = b->e;
  TRY(use_device(e));
  float *d_B = nullptr, *d_out = nullptr;
  CU(cudaMalloc(&d_B, sizeof(float) * (size_t)b->n_cols * k));
  if (cudaMalloc(&d_out, sizeof(float) * (size_t)b->n_rows * k) != cudaSuccess) {
    cudaFree(d_B);
    return fail(RFB_ERR_CUDA, "out of device memory for a %lld x %d product", (long long)b->n_rows, k);
  }
  cudaMemcpyAsync(d_B, B, sizeof(float) * (size_t)b->n_cols * k, cudaMemcpyHostToDevice, e->stream);
  const unsigned grid = (unsigned)std::min<long long>((b->n_rows + 7) / 8, (long long)e->num_sms * 8);
  for (int j0 = 0; j0 < k; j0 += 8) {
    block_matmul_kernel<8><<<grid, 256, 0, e->stream>>>(b->d, b->n_rows, b->n_cols, b->ld, d_B, k, j0,
                                                         std::min(8, k - j0), d_out, k);
    e->launches++;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)b->n_rows * k, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(d_B);
  cudaFree(d_out);
  if (ce != cudaSuccess) return fail(RFB_ERR_CUDA, "block matmul: %s", cudaGetErrorString(ce));
  return RFB_OK;
}

// Host memory -> pinned slot with a few threads (one memcpy thread moves ~10 GB/s, a PCIe 5 x16 link ~55 GB/s).
static void parallel_memcpy(void* dst, const void* src, size_t bytes, int threads) {
  threads = (int)std::min<size_t>((size_t)std::max(threads, 1), std::max<size_t>(bytes >> 22, 1));
  if (threads <= 1) { memcpy(dst, src, bytes); return; }
  const size_t per = ((bytes / threads) + 4095) & ~(size_t)4095;
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) {
    const size_t lo = std::min(bytes, per * t), hi = std::min(bytes, per * (t + 1));
    if (hi > lo) pool.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
  }
  memcpy(dst, src, std::min(bytes, per));
  for (auto& th : pool) th.join();
}

static int ring_reserve(rfb_engine* e, size_t bytes, bool need_pinned) {
  if (!e->copy_stream) {
    CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < rfb_engine::kRingSlots; ++k) {
      CU(cudaEventCreateWithFlags(&e->ring_copied[k], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e->ring_consumed[k], cudaEventDisableTiming));
    }
  }
  if (e->ring_bytes < bytes) {
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaStreamSynchronize(e->copy_stream));
    for (int k = 0; k < rfb_engine::kRingSlots; ++k) {
      if (e->ring_pinned[k]) { cudaFreeHost(e->ring_pinned[k]); e->ring_pinned[k] = nullptr; }
      if (e->ring_dev[k]) { cudaFree(e->ring_dev[k]); e->ring_dev[k] = nullptr; }
    }
    e->ring_bytes = 0;
    for (int k = 0; k < rfb_engine::kRingSlots; ++k) CU(cudaMalloc(&e->ring_dev[k], bytes));
    e->ring_bytes = bytes;
  }
  if (need_pinned)
    for (int k = 0; k < rfb_engine::kRingSlots; ++k)
      if (!e->ring_pinned[k]) CU(cudaHostAlloc(&e->ring_pinned[k], e->ring_bytes, cudaHostAllocDefault));
  return RFB_OK;
}

static bool is_pinned_host(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeHost;
}

extern "C" int rfb_block_project(rfb_block* dst, int64_t dst_row0, int64_t n_out, int col0,
                                 const float* stim, int on_device, int64_t src0, int64_t n_src,
                                 int64_t n_raw_total, int n_pix, int64_t first_t, int nlag, int shift,
                                 const float* St, int df_t, const float* Sxy, int n_q) {
  if (!dst || !stim || !St) return fail(RFB_ERR_INVALID, "NULL argument");
  if (nlag < 1 || df_t < 1 || n_pix < 1 || n_q < 1) return fail(RFB_ERR_INVALID, "bad dims");
  if (!Sxy && (n_pix != 1 || n_q != 1)) return fail(RFB_ERR_INVALID, "Sxy is NULL but n_pix/n_q != 1");
  if (dst_row0 < 0 || n_out < 0 || dst_row0 + n_out > dst->n_rows)
    return fail(RFB_ERR_INVALID, "destination rows out of bounds");
  if (col0 < 0 || (int64_t)col0 + (int64_t)df_t * n_q > dst->n_cols) return fail(RFB_ERR_INVALID, "destination columns out of bounds");
  if (n_src < 0 || src0 < 0 || src0 + n_src > n_raw_total) return fail(RFB_ERR_INVALID, "bad source range");
  if (n_out == 0) return RFB_OK;
  rfb_engine* e = dst->e;
  TRY(use_device(e));

  const int64_t front = (nlag + shift > 0) ? (int64_t)nlag + shift - 1 : 0;   // rfest/utils.py:54-57
  // every raw row this call needs must have been supplied
  {
    int64_t lo = first_t - front, hi = first_t + n_out - 1 + nlag - 1 - front;
    lo = std::max<int64_t>(lo, 0);
    hi = std::min<int64_t>(hi, n_raw_total - 1);
    if (lo <= hi && (lo < src0 || hi >= src0 + n_src))
      return fail(RFB_ERR_INVALID,
                  "rows [%lld, %lld] of the input are needed but [%lld, %lld) were supplied (halo missing)",
                  (long long)lo, (long long)hi, (long long)src0, (long long)(src0 + n_src));
  }

  // The temporal kernels hold up to 16 basis columns per thread; a wider time basis (a filter without a
  // spline basis hands over St = I[nlag], a spline with df[0] > 16) is applied in groups of <= 16 columns,
  // each group writing its own column range of the block.
  struct Group { int d0, dw, DFT; size_t off; };
  std::vector<Group> groups;
  std::vector<float> stp;                                   // all groups' [nlag x DFT] padded bases, back to back
  for (int d0 = 0; d0 < df_t; d0 += 16) {
    const int dw = std::min(16, df_t - d0);
    const int DFT = dw <= 1 ? 1 : dw <= 2 ? 2 : dw <= 4 ? 4 : dw <= 8 ? 8 : dw <= 12 ? 12 : 16;
    groups.push_back(Group{d0, dw, DFT, stp.size()});
    stp.resize(stp.size() + (size_t)nlag * DFT, 0.f);
    for (int i = 0; i < nlag; ++i)
      for (int d = 0; d < dw; ++d) stp[groups.back().off + (size_t)i * DFT + d] = St[(size_t)i * df_t + d0 + d];
  }
  const bool temporal_const = nlag <= kTempMaxLag && e->project_tc != 2;
  std::vector<float> cst((size_t)kTempMaxLag * 16, 0.f);
  auto upload_const = [&](const Group& g) -> int {          // zero-padded copy in constant memory (stream ordered)
    std::fill(cst.begin(), cst.end(), 0.f);
    std::copy(stp.begin() + g.off, stp.begin() + g.off + (size_t)nlag * g.DFT, cst.begin());
    CU(cudaMemcpyToSymbolAsync(c_st, cst.data(), sizeof(float) * cst.size(), 0, cudaMemcpyHostToDevice, e->stream));
    return RFB_OK;
  };
  TRY(scratch_reserve(e, e->stbuf, sizeof(float) * stp.size()));
  CU(cudaMemcpyAsync(e->stbuf.p, stp.data(), sizeof(float) * stp.size(), cudaMemcpyHostToDevice, e->stream));
  if (temporal_const && groups.size() == 1) TRY(upload_const(groups[0]));
  CU(cudaStreamSynchronize(e->stream));

  TcPlan tcp;
  TRY(tc_prepare(e, Sxy, n_pix, n_q, &tcp));
  const float* d_sxy = nullptr;
  if (Sxy) {
    TRY(scratch_reserve(e, e->sxybuf, sizeof(float) * (size_t)n_pix * n_q));
    CU(cudaMemcpyAsync(e->sxybuf.p, Sxy, sizeof(float) * (size_t)n_pix * n_q, cudaMemcpyHostToDevice, e->stream));
    d_sxy = (const float*)e->sxybuf.p;
  }

  // rows per slab.  Device input only needs the Z scratch (<= 1 GB).  Host input goes through the ring:
  // slabs of <= 64 MB (<= 256 MB on the synchronous comparison path).
  const bool ring = !on_device && e->h2d_overlap != 0;
  int64_t chunk = on_device ? (int64_t)(256u << 20) / n_q : (int64_t)((ring ? 16u : 64u) << 20) / n_pix;
  chunk = std::min<int64_t>(std::max<int64_t>(chunk, 4096), on_device ? (1 << 23) : (1 << 20));
  if (e->project_slab_rows > 0) chunk = e->project_slab_rows;
  const bool src_pinned = ring && is_pinned_host(stim);
  if (ring) TRY(ring_reserve(e, sizeof(float) * (size_t)(chunk + nlag) * n_pix, !src_pinned));
  e->last_project_h2d_bytes = 0.0;
  int64_t slab = 0;
  for (int64_t k0 = 0; k0 < n_out; k0 += chunk, ++slab) {
    const int64_t kc = std::min<int64_t>(chunk, n_out - k0);
    int64_t lo = first_t + k0 - front, hi = first_t + k0 + kc - 1 + nlag - 1 - front;
    lo = std::max<int64_t>(lo, 0);
    hi = std::min<int64_t>(hi, n_raw_total - 1);
    const int64_t nz = hi >= lo ? hi - lo + 1 : 0;
    const int slot = (int)(slab % rfb_engine::kRingSlots);

    const float* d_z = nullptr;
    int ldz = n_q;
    if (nz > 0) {
      const float* d_stim = nullptr;
      const float* h_src = stim + (lo - src0) * (int64_t)n_pix;
      const size_t bytes = sizeof(float) * (size_t)nz * n_pix;
      if (on_device) {
        d_stim = h_src;
      } else if (ring) {
        const void* from = h_src;
        if (!src_pinned) {
          // the pinned slot is free once the DMA that last read it has finished
          if (slab >= rfb_engine::kRingSlots) CU(cudaEventSynchronize(e->ring_copied[slot]));
          parallel_memcpy(e->ring_pinned[slot], h_src, bytes, e->h2d_threads);
          from = e->ring_pinned[slot];
        }
        // the device slab is free once the kernels of the slab that last used it have run
        if (slab >= rfb_engine::kRingSlots) CU(cudaStreamWaitEvent(e->copy_stream, e->ring_consumed[slot], 0));
        CU(cudaMemcpyAsync(e->ring_dev[slot], from, bytes, cudaMemcpyHostToDevice, e->copy_stream));
        CU(cudaEventRecord(e->ring_copied[slot], e->copy_stream));
        CU(cudaStreamWaitEvent(e->stream, e->ring_copied[slot], 0));
        d_stim = (const float*)e->ring_dev[slot];
        e->last_project_h2d_bytes += (double)bytes;
      } else {
        TRY(scratch_reserve(e, e->stage, bytes));
        CU(cudaMemcpyAsync(e->stage.p, h_src, bytes, cudaMemcpyHostToDevice, e->stream));
        d_stim = (const float*)e->stage.p;
        e->last_project_h2d_bytes += (double)bytes;
      }
      if (Sxy && tcp.fused_ok && groups.size() == 1 && temporal_const) {
        // one kernel for the slab: Z stays in shared memory
        TemporalArgs t{};
        t.out = dst->d;
        t.out_ld = dst->ld;
        t.out_row0 = dst_row0 + k0;
        t.n_out = kc;
        t.col0 = col0;
        t.nlag = nlag;
        t.df_t = groups[0].dw;
        if (tc_launch_fused(e, tcp, d_stim, nz, n_pix, n_q, groups[0].DFT, t, first_t + k0 - front - lo)) {
          CU(cudaGetLastError());
          if (ring) CU(cudaEventRecord(e->ring_consumed[slot], e->stream));
          if (!on_device && !ring) CU(cudaStreamSynchronize(e->stream));
          continue;
        }
      }
      if (Sxy) {
        TRY(scratch_reserve(e, e->zbuf, sizeof(float) * (size_t)nz * n_q));
        if (!(tcp.ok && tc_launch(e, tcp, d_stim, nz, n_pix, n_q, (float*)e->zbuf.p))) {
          dim3 grid((unsigned)((nz + kSpTM - 1) / kSpTM), (unsigned)((n_q + kSpTN - 1) / kSpTN));
          spatial_project_kernel<<<grid, 512, 0, e->stream>>>(d_stim, nz, n_pix, d_sxy, n_q, (float*)e->zbuf.p, n_q);
          e->launches++;
        }
        CU(cudaGetLastError());
        d_z = (const float*)e->zbuf.p;
      } else {
        d_z = d_stim;
        ldz = 1;
      }
    }
    for (const Group& g : groups) {
      if (temporal_const && groups.size() > 1) TRY(upload_const(g));
      TemporalArgs a{};
      a.Z = d_z;
      a.z_src0 = lo;
      a.n_z = nz;
      a.ldz = ldz;
      a.n_raw_total = n_raw_total;
      a.out = dst->d;
      a.out_ld = dst->ld;
      a.out_row0 = dst_row0 + k0;
      a.n_out = kc;
      a.col0 = col0 + g.d0 * n_q;
      a.first_t = first_t + k0;
      a.front = front;
      a.nlag = nlag;
      a.df_t = g.dw;
      a.n_q = n_q;
      a.St = (const float*)e->stbuf.p + g.off;
      switch (g.DFT) {
        case 1: launch_temporal<1>(a, e->stream, temporal_const); break;
        case 2: launch_temporal<2>(a, e->stream, temporal_const); break;
        case 4: launch_temporal<4>(a, e->stream, temporal_const); break;
        case 8: launch_temporal<8>(a, e->stream, temporal_const); break;
        case 12: launch_temporal<12>(a, e->stream, temporal_const); break;
        default: launch_temporal<16>(a, e->stream, temporal_const); break;
      }
      e->launches++;
      CU(cudaGetLastError());
    }
    if (ring && nz > 0) CU(cudaEventRecord(e->ring_consumed[slot], e->stream));
    if (!on_device && !ring) CU(cudaStreamSynchronize(e->stream));  // the single staging buffer is reused
  }
  CU(cudaStreamSynchronize(e->stream));
  return RFB_OK;
}

extern "C" int rfb_engine_last_project_h2d_bytes(rfb_engine* e, double* out) {
  if (!e || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  *out = e->last_project_h2d_bytes;
  return RFB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: model
// ------------------------------------------------------------------------------------------------
static bool valid_nl(int nl) { return nl >= RFB_NL_NONE && nl <= RFB_NL_LEAKY_RELU; }

extern "C" int rfb_model_create(rfb_engine* e, int distr, int output_nl, rfb_model** out) {
  if (!e || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  if (distr != RFB_DISTR_GAUSSIAN && distr != RFB_DISTR_POISSON)
    return fail(RFB_ERR_UNSUPPORTED, "unknown distribution %d", distr);
  if (!valid_nl(output_nl) && output_nl != RFB_NL_SOFTMAX)
    return fail(RFB_ERR_UNSUPPORTED, "unknown output nonlinearity %d", output_nl);
  rfb_model* m = new rfb_model();
  m->e = e;
  m->distr = distr;
  m->out_nl = output_nl;
  *out = m;
  return RFB_OK;
}

static void free_sweep_buffers(rfb_model* m) {
  for (auto& sb : m->sb) {
    cudaFree(sb.warp_acc);
    cudaFree(sb.cta_part);
    sb = SweepBuffers();
  }
}

extern "C" int rfb_model_destroy(rfb_model* m) {
  if (!m) return RFB_OK;
  cudaSetDevice(m->e->device);
  cudaStreamSynchronize(m->e->stream);
  fit_release(m);
  free_paramset(m->cur);
  free_paramset(m->tmp);
  free_sweep_buffers(m);
  cudaFree(m->fin_eval);
  for (auto& ds : m->data) {
    cudaFree(ds.y);
    cudaFree(ds.r_buf);
  }
  for (int k = 0; k < 3; ++k) {
    cudaFree(m->sm_scal[k]);
    cudaFree(m->sm_part[k]);
    cudaFree(m->sm_f[k]);
    cudaFree(m->smx_fin[k]);
    cudaFree(m->smx_scale[k]);
  }
  delete m;
  return RFB_OK;
}

extern "C" int rfb_model_clear_filters(rfb_model* m) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  TRY(use_device(m->e));
  CU(cudaStreamSynchronize(m->e->stream));
  fit_release(m);
  m->filters.clear();
  m->have_params = false;
  invalidate_plan(m);
  free_sweep_buffers(m);
  return RFB_OK;
}

extern "C" int rfb_model_add_filter(rfb_model* m, int slot, int filter_nl, int n_coef) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  if (slot < 0 || slot >= kMaxSlots) return fail(RFB_ERR_UNSUPPORTED, "slot %d out of range (max %d)", slot, kMaxSlots);
  if (!valid_nl(filter_nl) && filter_nl != RFB_NL_SOFTMAX)
    return fail(RFB_ERR_UNSUPPORTED, "unknown filter nonlinearity %d", filter_nl);
  if (n_coef < 1) return fail(RFB_ERR_INVALID, "filter needs at least one coefficient");
  if (m->fit.active) return fail(RFB_ERR_STATE, "cannot add filters during a fit");
  m->filters.push_back({slot, filter_nl, n_coef});
  m->have_params = false;
  invalidate_plan(m);
  free_sweep_buffers(m);
  return (int)m->filters.size() - 1;
}

// Upload the response of a bound data set (padded by 8 zero floats: the staged sweeps copy the responses of
// a tile as one 32-byte chunk) and reduce n, sum y, sum y^2 over the communicator.  Reuses the device copy.
static int upload_y(rfb_model* m, DataSet& ds, const float* y, int y_on_device) {
  rfb_engine* e = m->e;
  const int64_t n = ds.n;
  double mom[3] = {(double)n, 0.0, 0.0};
  if (y) {
    if (!ds.y) {
      CU(cudaMalloc(&ds.y, sizeof(float) * ((size_t)n + 8)));
      CU(cudaMemsetAsync(ds.y + n, 0, sizeof(float) * 8, e->stream));
    }
    CU(cudaMemcpyAsync(ds.y, y, sizeof(float) * (size_t)n,
                       y_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    ds.has_y = true;
    const int blocks_y = (int)std::min<int64_t>(512, (n + 255) / 256);
    TRY(scratch_reserve(e, e->small, sizeof(double) * 2 * 512));
    y_moments_kernel<<<blocks_y, 256, 0, e->stream>>>(ds.y, n, (double*)e->small.p);
    e->launches++;
    CU(cudaGetLastError());
    std::vector<double> part(2 * blocks_y);
    CU(cudaMemcpyAsync(part.data(), e->small.p, sizeof(double) * 2 * blocks_y, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < blocks_y; ++k) {
      mom[1] += part[2 * k];
      mom[2] += part[2 * k + 1];
    }
  }
  TRY(rfb_engine_allreduce_f64(e, mom, 3));
  ds.n_total = mom[0];
  ds.sy = mom[1];
  ds.syy = mom[2];
  return RFB_OK;
}

extern "C" int rfb_model_set_data(rfb_model* m, int kind, int n_slots, rfb_block* const* blocks,
                                  const float* y, int y_on_device, int64_t n) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  if (n_slots < 1 || n_slots > kMaxSlots || !blocks) return fail(RFB_ERR_INVALID, "bad slots");
  if (n <= 0) return fail(RFB_ERR_INVALID, "data set is empty");
  if (m->fit.active) return fail(RFB_ERR_STATE, "cannot rebind data during a fit");
  rfb_engine* e = m->e;
  TRY(use_device(e));
  for (int s = 0; s < n_slots; ++s) {
    rfb_block* b = blocks[s];
    if (!b) continue;  // filter without data in this set (allowed for prediction)
    if (b->e != e) return fail(RFB_ERR_INVALID, "block %d belongs to another engine", s);
    if (b->n_rows != n) return fail(RFB_ERR_INVALID, "block %d has %lld rows, expected %lld", s,
                                    (long long)b->n_rows, (long long)n);
    if (m->slot_cols[s] != 0 && m->slot_cols[s] != (int)b->ld) {
      if (kind != RFB_KIND_TRAIN)
        return fail(RFB_ERR_INVALID, "block %d has %d columns but the slot was registered with %d", s,
                    (int)b->ld, m->slot_cols[s]);
      invalidate_plan(m);
    }
    if (m->slot_cols[s] != (int)b->ld) {
      m->slot_cols[s] = (int)b->ld;
      invalidate_plan(m);
    }
  }
  DataSet& ds = m->data[kind];
  CU(cudaStreamSynchronize(e->stream));
  cudaFree(ds.y);
  cudaFree(ds.r_buf);
  ds = DataSet();
  ds.bound = true;
  ds.n_slots = n_slots;
  for (int s = 0; s < n_slots; ++s) ds.blocks[s] = blocks[s];
  ds.n = n;
  cudaFree(m->sb[kind].warp_acc);
  cudaFree(m->sb[kind].cta_part);
  m->sb[kind] = SweepBuffers();
  return upload_y(m, ds, y, y_on_device);
}

extern "C" int rfb_model_set_y(rfb_model* m, int kind, const float* y, int y_on_device, int64_t n) {
  if (!m || !y) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  if (m->fit.active) return fail(RFB_ERR_STATE, "cannot rebind data during a fit");
  DataSet& ds = m->data[kind];
  if (!ds.bound) return fail(RFB_ERR_STATE, "no data bound for kind %d (rfb_model_set_data first)", kind);
  if (ds.n != n) return fail(RFB_ERR_INVALID, "y has %lld samples, the bound design %lld", (long long)n, (long long)ds.n);
  TRY(use_device(m->e));
  CU(cudaStreamSynchronize(m->e->stream));
  return upload_y(m, ds, y, y_on_device);
}

extern "C" int rfb_model_n_params(rfb_model* m, int* n_coef_total, int* n_intercepts) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  int P = 0;
  for (auto& f : m->filters) P += f.n_coef;
  if (n_coef_total) *n_coef_total = P;
  if (n_intercepts) *n_intercepts = (int)m->filters.size() + 1;
  return RFB_OK;
}

extern "C" int rfb_model_set_params(rfb_model* m, const float* b, const double* intercepts) {
  if (!m || !b || !intercepts) return fail(RFB_ERR_INVALID, "NULL argument");
  if (m->fit.active) return fail(RFB_ERR_STATE, "cannot set parameters during a fit");
  int P = 0, K = 0;
  rfb_model_n_params(m, &P, &K);
  m->host_b.assign(b, b + P);
  m->host_ic.assign(intercepts, intercepts + K);
  m->have_params = true;
  if (m->plan.ok) {
    rfb_engine* e = m->e;
    TRY(use_device(e));
    CU(cudaMemcpyAsync(m->cur.theta, m->host_b.data(), sizeof(float) * P, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(m->cur.ic, m->host_ic.data(), sizeof(double) * K, cudaMemcpyHostToDevice, e->stream));
    publish_params_kernel<<<1, kUpdateThreads, 0, e->stream>>>(publish_args(m, m->cur));
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
  }
  return RFB_OK;
}

extern "C" int rfb_model_get_params(rfb_model* m, float* b, double* intercepts) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  if (!m->have_params) return fail(RFB_ERR_STATE, "no parameters set");
  if (m->plan.ok) {
    rfb_engine* e = m->e;
    TRY(use_device(e));
    CU(cudaMemcpyAsync(m->host_b.data(), m->cur.theta, sizeof(float) * m->plan.P, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(m->host_ic.data(), m->cur.ic, sizeof(double) * (m->plan.F + 1), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  if (b) memcpy(b, m->host_b.data(), sizeof(float) * m->host_b.size());
  if (intercepts) memcpy(intercepts, m->host_ic.data(), sizeof(double) * m->host_ic.size());
  return RFB_OK;
}

// only_filter >= 0: evaluate that filter alone with all intercepts zero (the `u` of
// _get_filter_variance, _glm.py:1125-1131); keep_r: leave the predictions in ds.r_buf on the device.
static int eval_impl(rfb_model* m, int kind, const float* b, const double* intercepts, float* r_out,
                     double* sums, double* grad_b, double* grad_ic, int only_filter = -1, bool keep_r = false) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  DataSet& ds = m->data[kind];
  if (!ds.bound) return fail(RFB_ERR_STATE, "no data bound for kind %d", kind);
  if ((b == nullptr) != (intercepts == nullptr)) return fail(RFB_ERR_INVALID, "pass both b and intercepts or neither");
  if (!b && !m->have_params) return fail(RFB_ERR_STATE, "no parameters set");
  TRY(ensure_plan(m));
  rfb_engine* e = m->e;
  const Plan& p = m->plan;
  TRY(use_device(e));

  // parameter image for this evaluation; filters without data in this set drop out
  // (forwardpass only loops over self.X[kind], _glm.py:555)
  std::vector<float> hb(p.P);
  std::vector<double> hic(p.F + 1);
  if (b) {
    std::copy(b, b + p.P, hb.begin());
    std::copy(intercepts, intercepts + p.F + 1, hic.begin());
  } else {
    TRY(rfb_model_get_params(m, hb.data(), hic.data()));
  }
  std::vector<int> head_of(p.head_of), head_nl(p.head_nl);
  std::vector<char> head_live(p.Ht, 0);
  for (int f = 0; f < p.F; ++f) {
    const int s = m->filters[f].slot;
    const bool present = s < ds.n_slots && ds.blocks[s] != nullptr && (only_filter < 0 || only_filter == f);
    if (present) {
      head_live[p.head_of[f]] = 1;
    } else {
      for (int j = 0; j < m->filters[f].n_coef; ++j) hb[p.coef0[f] + j] = 0.f;
      hic[f] = 0.0;
    }
  }
  for (int h = 0; h < p.H; ++h)
    if (!head_live[h]) head_nl[h] = kNlAbsent;
  if (only_filter >= 0)
    for (auto& c : hic) c = 0.0;
  CU(cudaMemcpyAsync(m->tmp.theta, hb.data(), sizeof(float) * p.P, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(m->tmp.ic, hic.data(), sizeof(double) * (p.F + 1), cudaMemcpyHostToDevice, e->stream));
  publish_params_kernel<<<1, kUpdateThreads, 0, e->stream>>>(publish_args(m, m->tmp));
  e->launches++;
  CU(cudaGetLastError());

  float* d_r = nullptr;
  if (r_out || keep_r) {
    if (!ds.r_buf) CU(cudaMalloc(&ds.r_buf, sizeof(float) * r_buf_floats(ds.n)));
    d_r = ds.r_buf;
  }
  const bool want_grad = grad_b != nullptr || grad_ic != nullptr;
  if (want_grad && !ds.has_y) return fail(RFB_ERR_STATE, "a gradient needs the response of kind %d", kind);
  const float* d_sm = nullptr;
  TRY(ensure_smx_buffers(m, kind));
  TRY(smx_prepare(m, kind, m->tmp, head_nl.data(), want_grad));
  if (m->out_nl == RFB_NL_SOFTMAX) {
    TRY(ensure_softmax_buffers(m, kind));
    if (d_r) d_r = ds.r_buf;
    TRY(softmax_prepare(m, kind, m->tmp, head_nl.data(), want_grad));
    d_sm = m->sm_f[kind];
  }
  TRY(launch_sweep(m, kind, m->tmp, head_nl.data(), want_grad, d_r, -1, d_sm));
  const int tail = p.Ht * p.ctot;
  double* d_sc = m->fin_eval + (want_grad ? tail : 0);
  std::vector<double> fin;
  if (want_grad) {
    TRY(launch_reduce(m, kind, 0, p.NE, m->fin_eval, -1, 0, 0, nullptr));
    TRY(allreduce_device_f64(e, m->fin_eval, p.NE));
    TRY(smx_fix(m, kind, head_nl.data(), m->fin_eval));
    fin.resize(p.NE);
    CU(cudaMemcpyAsync(fin.data(), m->fin_eval, sizeof(double) * p.NE, cudaMemcpyDeviceToHost, e->stream));
  } else {
    // only the scalar tail matters for a forward pass
    TRY(launch_reduce(m, kind, tail, kNumScalars, d_sc, -1, 0, 0, nullptr));
    TRY(allreduce_device_f64(e, d_sc, kNumScalars));
  }
  double sc[kNumScalars];
  CU(cudaMemcpyAsync(sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, e->stream));
  if (r_out) CU(cudaMemcpyAsync(r_out, d_r, sizeof(float) * (size_t)ds.n, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (sums) {
    sums[RFB_SUM_LOSS_A] = sc[S_LOSS_A];
    sums[RFB_SUM_LOSS_B] = sc[S_LOSS_B];
    sums[RFB_SUM_R] = sc[S_R];
    sums[RFB_SUM_RR] = sc[S_RR];
    sums[RFB_SUM_YR] = sc[S_YR];
    sums[RFB_SUM_SQERR] = sc[S_SQERR];
    sums[RFB_SUM_Y] = ds.sy;
    sums[RFB_SUM_YY] = ds.syy;
  }
  if (grad_b)
    for (int j = 0; j < p.P; ++j) grad_b[j] = fin[p.entry_of[j]];
  if (grad_ic) {
    for (int f = 0; f < p.F; ++f) grad_ic[f] = fin[tail + S_GHEAD + p.head_of[f]];
    grad_ic[p.F] = fin[tail + S_GGLOBAL];
  }
  return RFB_OK;
}

extern "C" int rfb_model_eval(rfb_model* m, int kind, const float* b, const double* intercepts,
                              float* r_out, double* sums) {
  return eval_impl(m, kind, b, intercepts, r_out, sums, nullptr, nullptr);
}

extern "C" int rfb_model_loss_of_predictions(rfb_model* m, int kind, const float* r, int r_on_device,
                                            double* loss_ab) {
  if (!m || !r || !loss_ab) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  DataSet& ds = m->data[kind];
  if (!ds.bound || !ds.has_y) return fail(RFB_ERR_STATE, "no response bound for kind %d", kind);
  rfb_engine* e = m->e;
  TRY(use_device(e));
  const int64_t n = ds.n;
  const float* d_r = r;
  if (!r_on_device) {
    TRY(scratch_reserve(e, e->stage, sizeof(float) * (size_t)n));
    CU(cudaMemcpyAsync(e->stage.p, r, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    d_r = (const float*)e->stage.p;
  }
  const int blocks = (int)std::min<int64_t>(512, (n + 255) / 256);
  TRY(scratch_reserve(e, e->small, sizeof(double) * 2 * 512));
  loss_terms_kernel<<<blocks, 256, 0, e->stream>>>(d_r, ds.y, n, m->distr, (double*)e->small.p);
  e->launches++;
  CU(cudaGetLastError());
  std::vector<double> part(2 * blocks);
  CU(cudaMemcpyAsync(part.data(), e->small.p, sizeof(double) * 2 * blocks, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  loss_ab[0] = loss_ab[1] = 0.0;
  for (int k = 0; k < blocks; ++k) {
    loss_ab[0] += part[2 * k];
    loss_ab[1] += part[2 * k + 1];
  }
  return rfb_engine_allreduce_f64(e, loss_ab, 2);
}

extern "C" int rfb_model_value_and_grad(rfb_model* m, int kind, const float* b, const double* intercepts,
                                        double* sums, double* grad_b, double* grad_icpt) {
  if (!grad_b || !grad_icpt) return fail(RFB_ERR_INVALID, "NULL gradient buffer");
  return eval_impl(m, kind, b, intercepts, nullptr, sums, grad_b, grad_icpt);
}

// Gram pieces over the column window [col_lo, col_lo + ncols) of the concatenated design, optionally
// row-weighted; results (float64, all-reduced over ranks) are copied to the host.
static int gram_impl(rfb_model* m, int kind, int col_lo, int ncols, const float* d_wgt, double* gram,
                     double* colsum, double* xty) {
  DataSet& ds = m->data[kind];
  const Plan& p = m->plan;
  rfb_engine* e = m->e;
  GramArgs a{};
  a.n_slots = p.n_slots;
  for (int s = 0; s < p.n_slots; ++s) {
    rfb_block* b = s < ds.n_slots ? ds.blocks[s] : nullptr;
    a.slot_ptr[s] = b ? b->d : nullptr;
    a.slot_ld[s] = b ? b->ld : 0;
  }
  for (int s = 0; s <= p.n_slots; ++s) a.slot_col0[s] = 4 * p.slot_chunk0[s];
  a.ctot = p.ctot;
  a.n_rows = ds.n;
  a.y = ds.has_y ? ds.y : nullptr;
  a.wgt = d_wgt;
  a.col_lo = col_lo;
  a.ncols = ncols;
  a.n_tiles = (ncols + kGramTile - 1) / kGramTile;
  a.n_pairs = a.n_tiles * (a.n_tiles + 1) / 2;
  // row slabs: about four balanced waves of CTAs (4 resident per SM: 206 registers x 64 threads), never shorter than 4096 rows
  // (tile pairs differ in work -- diagonal ones do 36 of 64 blocks -- so the tensor-core kernel takes twelve short
  // waves rather than four long ones: the ragged end of the launch shrinks with the CTA duration)
  const int resident = 4 * e->num_sms;
  int slabs = std::max(1, ((e->gram_mma ? 12 : 4) * resident) / a.n_pairs);
  slabs = (int)std::min<int64_t>(slabs, std::max<int64_t>(1, ds.n / 4096));
  const size_t n_entries = (size_t)ncols * ncols + 2 * (size_t)ncols;
  slabs = (int)std::min<size_t>(slabs, std::max<size_t>(1, ((size_t)256 << 20) / (sizeof(double) * n_entries)));   // <= 256 MB of partial sums
  a.n_slabs = slabs;
  static bool gram_attr = false;
  if (!gram_attr) {
    CU(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem));
    CU(cudaFuncSetAttribute(gram_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramMmaSmem));
    gram_attr = true;
  }
  // partial sums + result live in a scratch buffer the engine keeps (a cudaMalloc / cudaFree pair of a few hundred
  // MB per call cost up to tens of milliseconds on some boxes)
  if (scratch_reserve(e, e->gram_part, sizeof(double) * n_entries * ((size_t)slabs + 1)) != RFB_OK) {
    cudaGetLastError();
    return fail(RFB_ERR_CUDA, "out of device memory for the Gram partial sums (%zu bytes)",
                sizeof(double) * n_entries * (slabs + 1));
  }
  double* d_part = (double*)e->gram_part.p;
  double* d_out = d_part + n_entries * (size_t)slabs;
  a.part = d_part;
  if (e->gram_mma) gram_mma_kernel<<<dim3(a.n_pairs, slabs), kGramMmaThreads, kGramMmaSmem, e->stream>>>(a);
  else gram_kernel<<<dim3(a.n_pairs, slabs), kGramThreads, kGramSmem, e->stream>>>(a);
  gram_reduce_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, e->stream>>>(d_part, slabs, (long long)n_entries, d_out);
  e->launches += 2;
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  int rc = RFB_OK;
  if (ce != cudaSuccess) rc = fail(RFB_ERR_CUDA, "gram kernel failed: %s", cudaGetErrorString(ce));
  if (rc == RFB_OK) rc = allreduce_device_f64(e, d_out, n_entries);
  if (rc == RFB_OK) {
    const size_t cc = (size_t)ncols * ncols;
    if (gram) cudaMemcpyAsync(gram, d_out, sizeof(double) * cc, cudaMemcpyDeviceToHost, e->stream);
    if (colsum) cudaMemcpyAsync(colsum, d_out + cc, sizeof(double) * ncols, cudaMemcpyDeviceToHost, e->stream);
    if (xty) cudaMemcpyAsync(xty, d_out + cc + ncols, sizeof(double) * ncols, cudaMemcpyDeviceToHost, e->stream);
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) rc = fail(RFB_ERR_CUDA, "gram copy-back failed");
  }
  return rc;
}

extern "C" int rfb_model_gram(rfb_model* m, int kind, double* gram, double* colsum, double* xty, int* ctot_out) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  if (!m->data[kind].bound) return fail(RFB_ERR_STATE, "no data bound for kind %d", kind);
  TRY(ensure_plan(m));
  if (ctot_out) *ctot_out = m->plan.ctot;
  if (!gram && !colsum && !xty) return RFB_OK;     // size query
  TRY(use_device(m->e));
  return gram_impl(m, kind, 0, m->plan.ctot, nullptr, gram, colsum, xty);
}

extern "C" int rfb_model_filter_gram(rfb_model* m, int kind, int filter, const float* b, int weighted, double* gram) {
  if (!m || !gram) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  DataSet& ds = m->data[kind];
  if (!ds.bound) return fail(RFB_ERR_STATE, "no data bound for kind %d", kind);
  TRY(ensure_plan(m));
  const Plan& p = m->plan;
  if (filter < 0 || filter >= p.F) return fail(RFB_ERR_INVALID, "filter %d out of range", filter);
  rfb_engine* e = m->e;
  TRY(use_device(e));
  const Filter& fl = m->filters[filter];
  const float* d_w = nullptr;
  if (weighted) {
    if (!b) return fail(RFB_ERR_INVALID, "the weighted Gram needs the coefficients");
    std::vector<double> ic(p.F + 1, 0.0);
    TRY(eval_impl(m, kind, b, ic.data(), nullptr, nullptr, nullptr, nullptr, filter, true));
    inv_square_kernel<<<(unsigned)((ds.n + 255) / 256), 256, 0, e->stream>>>(ds.r_buf, ds.n);
    e->launches++;
    CU(cudaGetLastError());
    d_w = ds.r_buf;
  }
  return gram_impl(m, kind, 4 * p.slot_chunk0[fl.slot], fl.n_coef, d_w, gram, nullptr, nullptr);
}

extern "C" int rfb_model_response_band(rfb_model* m, int kind, const float* b, const double* intercepts,
                                       const float* const* V, float* pred, float* upper, float* lower,
                                       int on_device) {
  if (!m || !b || !intercepts || !V || !pred || !upper || !lower) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind < 0 || kind > 2) return fail(RFB_ERR_INVALID, "bad kind %d", kind);
  DataSet& ds = m->data[kind];
  if (!ds.bound) return fail(RFB_ERR_STATE, "no data bound for kind %d", kind);
  TRY(ensure_plan(m));
  const Plan& p = m->plan;
  rfb_engine* e = m->e;
  TRY(use_device(e));
  if (p.F > 2 * kMaxHeads) return fail(RFB_ERR_UNSUPPORTED, "response band: at most %d filters", 2 * kMaxHeads);
  const size_t n = (size_t)ds.n;
  // scratch: per filter lin + se, the three outputs when they go to the host, V and b images
  size_t v_floats = 0;
  for (int f = 0; f < p.F; ++f) v_floats += (size_t)m->filters[f].n_coef * m->filters[f].n_coef;
  float *d_rows = nullptr, *d_small = nullptr, *d_out = nullptr;
  void* d_smx = nullptr;
  auto cleanup = [&]() { cudaFree(d_rows); cudaFree(d_small); cudaFree(d_out); cudaFree(d_smx); };
  if (cudaMalloc(&d_rows, sizeof(float) * std::max<size_t>(n, 1) * 2 * p.F) != cudaSuccess) {
    cudaGetLastError();
    return fail(RFB_ERR_CUDA, "out of device memory");
  }
  if (cudaMalloc(&d_small, sizeof(float) * (v_floats + p.P + 4)) != cudaSuccess) { cleanup(); return fail(RFB_ERR_CUDA, "out of device memory"); }
  if (!on_device && cudaMalloc(&d_out, sizeof(float) * n * 3) != cudaSuccess) { cleanup(); return fail(RFB_ERR_CUDA, "out of device memory"); }
  float* d_b = d_small + v_floats;
  cudaMemcpyAsync(d_b, b, sizeof(float) * p.P, cudaMemcpyHostToDevice, e->stream);
  BandArgs ba{};
  ba.n_rows = ds.n;
  size_t v_off = 0;
  for (int f = 0; f < p.F; ++f) {
    const Filter& fl = m->filters[f];
    const size_t pp = (size_t)fl.n_coef * fl.n_coef;
    rfb_block* blk = fl.slot < ds.n_slots ? ds.blocks[fl.slot] : nullptr;
    if (blk) {                                  // forwardpass only loops over self.X[kind]
      const int i = ba.n_filters++;
      QuadformArgs qa{};
      qa.X = blk->d;
      qa.ld = blk->ld;
      qa.n_rows = ds.n;
      qa.p = fl.n_coef;
      qa.b = d_b + p.coef0[f];
      qa.lin = d_rows + (size_t)(2 * i) * n;
      qa.se = d_rows + (size_t)(2 * i + 1) * n;
      const bool has_v = V[f] != nullptr;
      if (has_v) cudaMemcpyAsync(d_small + v_off, V[f], sizeof(float) * pp, cudaMemcpyHostToDevice, e->stream);
      else cudaMemsetAsync(d_small + v_off, 0, sizeof(float) * pp, e->stream);
      qa.V = d_small + v_off;
      if (ds.n > 0) {
        quadform_kernel<<<(unsigned)((ds.n + kQfRows - 1) / kQfRows), 256, 0, e->stream>>>(qa);
        e->launches++;
      }
      ba.lin[i] = qa.lin;
      ba.se[i] = has_v ? qa.se : nullptr;
      ba.icpt[i] = (float)intercepts[f];
      ba.nl[i] = fl.nl;
    }
    v_off += pp;
  }
  ba.glob_icpt = (float)intercepts[p.F];
  ba.out_nl = m->out_nl;
  ba.pred = on_device ? pred : d_out;
  ba.upper = on_device ? upper : d_out + n;
  ba.lower = on_device ? lower : d_out + 2 * n;
  const int band_grid = (int)std::max<int64_t>(1, std::min<int64_t>(kBandBlocks, (ds.n + 255) / 256));
  // softmax lines are normalised over all samples (and ranks): sums first, one filter / the output at a time
  bool any_smx = m->out_nl == RFB_NL_SOFTMAX;
  for (int i = 0; i < ba.n_filters; ++i) any_smx = any_smx || ba.nl[i] == RFB_NL_SOFTMAX;
  if (any_smx) {
    if (cudaMalloc(&d_smx, sizeof(double) * (3 * kBandBlocks + 4) + sizeof(float) * 3 * (ba.n_filters + 1)) != cudaSuccess) {
      cleanup();
      return fail(RFB_ERR_CUDA, "out of device memory");
    }
    double* d_part = (double*)d_smx;
    double* d_sums = d_part + 3 * kBandBlocks;
    float* d_scale = (float*)(d_sums + 4);
    ba.part = d_part;
    ba.fscale = d_scale;
    ba.oscale = d_scale + 3 * ba.n_filters;
    auto normaliser = [&](int mode, int which, float* out) -> int {
      ba.mode = mode;
      ba.which = which;
      band_combine_kernel<<<band_grid, 256, 0, e->stream>>>(ba);
      band_sum_kernel<<<1, 32, 0, e->stream>>>(d_part, band_grid, d_sums);
      int rc = allreduce_device_f64(e, d_sums, 3);
      band_scale_kernel<<<1, 32, 0, e->stream>>>(d_sums, out);
      e->launches += 3;
      return rc;
    };
    int rc = RFB_OK;
    for (int i = 0; i < ba.n_filters && rc == RFB_OK; ++i)
      if (ba.nl[i] == RFB_NL_SOFTMAX) rc = normaliser(1, i, d_scale + 3 * i);
    if (rc == RFB_OK && m->out_nl == RFB_NL_SOFTMAX) rc = normaliser(2, 0, d_scale + 3 * ba.n_filters);
    if (rc != RFB_OK) { cudaStreamSynchronize(e->stream); cleanup(); return rc; }
    ba.mode = 0;
  }
  band_combine_kernel<<<band_grid, 256, 0, e->stream>>>(ba);
  e->launches++;
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess && !on_device) {
    cudaMemcpyAsync(pred, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream);
    cudaMemcpyAsync(upper, d_out + n, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream);
    cudaMemcpyAsync(lower, d_out + 2 * n, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cleanup();
  if (ce != cudaSuccess) return fail(RFB_ERR_CUDA, "response band failed: %s", cudaGetErrorString(ce));
  return RFB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: fit loop
// ------------------------------------------------------------------------------------------------
static int fit_release(rfb_model* m) {
  Fit& f = m->fit;
  if (f.exec) cudaGraphExecDestroy(f.exec);
  if (f.graph) cudaGraphDestroy(f.graph);
  cudaFree(f.step); cudaFree(f.cost_train); cudaFree(f.cost_dev); cudaFree(f.metric_train);
  cudaFree(f.metric_dev); cudaFree(f.ring_b); cudaFree(f.ring_ic); cudaFree(f.it);
  cudaFree(f.m); cudaFree(f.v); cudaFree(f.mi); cudaFree(f.vi); cudaFree(f.fin); cudaFree(f.stamps); cudaFree(f.small_part); cudaFree(f.small_flags); cudaFree(f.bc_f); cudaFree(f.bc_d);
  f = Fit();
  return RFB_OK;
}

static int fill_nan(double* d, int n, cudaStream_t s) {
  std::vector<double> h(n, std::nan(""));
  CU(cudaMemcpyAsync(d, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));
  return RFB_OK;
}


// One iteration's kernels in stream order: train sweep, dev sweep, fixed-order reduction,
// all-reduce over ranks, bookkeeping + Adam.  Captured into the graph by rfb_fit_begin and
// run eagerly (forward-only) by rfb_fit_finish.
static int enqueue_iteration(rfb_model* m, bool final_only, float* r_train, float* r_dev,
                             cudaEvent_t* ev = nullptr) {
  Fit& f = m->fit;
  const Plan& p = m->plan;
  rfb_engine* e = m->e;
  const bool softmax = m->out_nl == RFB_NL_SOFTMAX;
  if (ev) CU(cudaEventRecord(ev[0], e->stream));
  TRY(smx_prepare(m, RFB_KIND_TRAIN, m->cur, p.head_nl.data(), !final_only));
  if (softmax) TRY(softmax_prepare(m, RFB_KIND_TRAIN, m->cur, p.head_nl.data(), !final_only));
  f.stamp_now = true;
  int rc_sweep = launch_sweep(m, RFB_KIND_TRAIN, m->cur, p.head_nl.data(), !final_only, r_train, -1,
                              softmax ? m->sm_f[RFB_KIND_TRAIN] : nullptr);
  f.stamp_now = false;
  TRY(rc_sweep);
  if (ev) CU(cudaEventRecord(ev[1], e->stream));
  if (f.has_dev) {
    TRY(smx_prepare(m, RFB_KIND_DEV, m->cur, p.head_nl.data(), false));
    if (softmax) TRY(softmax_prepare(m, RFB_KIND_DEV, m->cur, p.head_nl.data(), false));
    f.stamp_now = true;
    rc_sweep = launch_sweep(m, RFB_KIND_DEV, m->cur, p.head_nl.data(), false, r_dev, -1,
                            softmax ? m->sm_f[RFB_KIND_DEV] : nullptr);
    f.stamp_now = false;
    TRY(rc_sweep);
  }
  if (ev) CU(cudaEventRecord(ev[2], e->stream));
  const int tail = p.Ht * p.ctot;
  // (softmax filter heads correct the summed vector between the exchange and the update: unfused path)
  const bool xchg = e->comm && e->use_xchg && e->xchg_ready && (long long)p.NE + kNumScalars <= e->xchg_cap &&
                    p.n_smx == 0;
  const int lo = final_only ? tail : 0;
  const int n_train = p.NE - lo;
  long long sum_lo = lo, sum_hi = (long long)p.NE + (f.has_dev ? kNumScalars : 0);
  if (xchg) {
    // reduction fused with the exchange: sums go straight into every rank's buffer over NVLink
    ReduceSeg s0{m->sb[RFB_KIND_TRAIN].cta_part + lo, m->sb[RFB_KIND_TRAIN].cur_grid, p.NE, n_train, nullptr};
    const int blocks0 = (n_train + kReduceEx - 1) / kReduceEx;
    ReduceSeg s1{nullptr, 0, 0, 0, nullptr};
    int blocks1 = 0;
    if (f.has_dev) {
      s1 = ReduceSeg{m->sb[RFB_KIND_DEV].cta_part + tail, m->sb[RFB_KIND_DEV].cur_grid, p.NE, kNumScalars, nullptr};
      blocks1 = (kNumScalars + kReduceEx - 1) / kReduceEx;
    }
    if (e->use_xchg == 3)   // one-CTA form (round 2): measured slower (43 vs 20 us at 2 GPUs), kept for comparison
      reduce_exchange1_kernel<<<1, kReduce1Threads, 0, e->stream>>>(s0, s1, (long long)lo, (long long)p.NE,
                                                                    xchg_args(e, true));
    else
      reduce_exchange_kernel<<<blocks0 + blocks1, dim3(kReduceEx, kReduceEy), 0, e->stream>>>(
          s0, s1, blocks0, (long long)lo, (long long)p.NE, xchg_args(e, true));
    e->launches++;
    CU(cudaGetLastError());
  } else {
    TRY(launch_reduce(m, RFB_KIND_TRAIN, lo, n_train, f.fin + lo, f.has_dev ? RFB_KIND_DEV : -1, tail,
                      kNumScalars, f.fin + p.NE));
    TRY(allreduce_device_f64(e, f.fin + lo, (size_t)(sum_hi - lo)));
    if (!final_only) TRY(smx_fix(m, RFB_KIND_TRAIN, p.head_nl.data(), f.fin));
  }
  if (ev) CU(cudaEventRecord(ev[3], e->stream));
  UpdateArgs ua = f.ua;
  ua.final_only = final_only ? 1 : 0;
  ua.x = xchg_args(e, xchg);
  ua.fin_local = f.fin;
  ua.sum_lo = sum_lo;
  ua.sum_hi = sum_hi;
  update_kernel<<<1, kUpdateThreads, 0, e->stream>>>(ua);
  e->launches++;
  CU(cudaGetLastError());
  if (ev) CU(cudaEventRecord(ev[4], e->stream));
  return RFB_OK;
}

// Decide whether this fit runs as the persistent small-model kernel and size it.  Eligible: one GPU, no
// output softmax, a register-resident head count, <= 64 design columns, and both data sets small enough
// to live in the shared memory of one CTA per SM (the regime where the graph path is launch-bound).
static int plan_small_fit(rfb_model* m) {
  Fit& f = m->fit;
  const Plan& p = m->plan;
  rfb_engine* e = m->e;
  f.small = false;
  if (!e->fit_persistent || e->world > 1 || m->out_nl == RFB_NL_SOFTMAX || p.generic || p.n_smx) return RFB_OK;
  if (p.F > 100 || p.n_chunks > kSmallMaxChunks || !(p.Ht == 1 || p.Ht == 2 || p.Ht == 3 || p.Ht == 5)) return RFB_OK;
  const DataSet& tr = m->data[RFB_KIND_TRAIN];
  const DataSet& dev = m->data[RFB_KIND_DEV];
  const int64_t n_dev = f.has_dev ? dev.n : 0;
  const double bytes = 4.0 * (double)(tr.n + n_dev) * (p.ctot + 1);
  if (bytes > 8.0 * 1024 * 1024 && e->fit_persistent < 2) return RFB_OK;     // 2, 3: whenever the data fit (tests)
  const void* fn = p.Ht == 1 ? (const void*)fit_small_kernel<1> : p.Ht == 2 ? (const void*)fit_small_kernel<2>
                 : p.Ht == 3 ? (const void*)fit_small_kernel<3> : (const void*)fit_small_kernel<5>;
    int lpr = 1;
  while (lpr < p.n_chunks) lpr <<= 1;
  const int rows_per_pass = (kSmallThreads / 32) * (32 / lpr);
  // measured at config 1 (16k rows): 64 CTAs x 4 passes beat 32 x 8 and 125 x 2 (the barrier and the CTA-ordered
  // totals grow with the grid, the row loop shrinks with it)
  int64_t grid = std::min<int64_t>(e->num_sms, std::max<int64_t>(1, (tr.n + 4 * rows_per_pass - 1) / (4 * rows_per_pass)));
  if (e->fit_small_ctas > 0) grid = std::min<int64_t>(e->num_sms, e->fit_small_ctas);
  const int rows_tr = (int)((tr.n + grid - 1) / grid), rows_dv = (int)((n_dev + grid - 1) / grid);
  const size_t smem = fit_small_smem_bytes(rows_tr, rows_dv, p.ctot, p.Ht, p.P, p.F);
  if (smem > 200 * 1024) return RFB_OK;
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kSmallThreads, smem));
  if (occ < 1) return RFB_OK;
  const size_t row = (size_t)p.NE + kNumScalars;
  CU(cudaMalloc(&f.small_part, sizeof(double) * 2 * row * (size_t)grid));
  CU(cudaMemsetAsync(f.small_part, 0, sizeof(double) * 2 * row * (size_t)grid, e->stream));
  CU(cudaMalloc(&f.small_flags, sizeof(unsigned long long) * (size_t)grid));
  CU(cudaMemsetAsync(f.small_flags, 0, sizeof(unsigned long long) * (size_t)grid, e->stream));
  f.small_epoch = 0;
  SmallArgs& sa = f.sa;
  sa = SmallArgs{};
  sa.tr = make_sweep_args(m, RFB_KIND_TRAIN, m->cur, p.head_nl.data(), -1, nullptr);
  if (f.has_dev) sa.dv = make_sweep_args(m, RFB_KIND_DEV, m->cur, p.head_nl.data(), -1, nullptr);
  else sa.dv.n_rows = 0;
  sa.ua = f.ua;
  sa.ua.final_only = 0;
  sa.ua.x = xchg_args(e, false);
  sa.ua.fin_local = f.fin;
  sa.ua.sum_lo = 0;
  sa.ua.sum_hi = 0;
  sa.part = f.small_part;
  sa.flags = e->fit_persistent == 3 ? f.small_flags : nullptr;   // 3: flag exchange instead of the grid barrier
  sa.rows_tr = rows_tr;
  sa.rows_dv = rows_dv;
  sa.NE = p.NE;
  f.small_grid = (int)grid;
  f.small_smem = smem;
  f.small_fn = fn;
  f.small = true;
  return RFB_OK;
}

extern "C" int rfb_fit_begin(rfb_model* m, int max_iters, const double* step_sizes, double alpha,
                             double beta, int metric) {
  if (!m || !step_sizes) return fail(RFB_ERR_INVALID, "NULL argument");
  if (max_iters < 1) return fail(RFB_ERR_INVALID, "max_iters must be >= 1");
  if (metric < RFB_METRIC_NONE || metric > RFB_METRIC_CORRCOEF) return fail(RFB_ERR_INVALID, "bad metric");
  if (!m->data[RFB_KIND_TRAIN].bound || !m->data[RFB_KIND_TRAIN].has_y)
    return fail(RFB_ERR_STATE, "no training data / response bound");
  if (!m->have_params) return fail(RFB_ERR_STATE, "no initial parameters set");
  rfb_engine* e = m->e;
  TRY(use_device(e));
  CU(cudaStreamSynchronize(e->stream));
  fit_release(m);
  TRY(ensure_plan(m));
  const Plan& p = m->plan;
  const DataSet& tr = m->data[RFB_KIND_TRAIN];
  const DataSet& dev = m->data[RFB_KIND_DEV];
  for (int s = 0; s < p.n_slots; ++s)
    if (!(s < tr.n_slots && tr.blocks[s])) return fail(RFB_ERR_STATE, "training data has no block for slot %d", s);
  Fit& f = m->fit;
  f.max_iters = max_iters;
  f.metric = metric;
  f.has_dev = dev.bound && dev.has_y;
  if (f.has_dev)
    for (int s = 0; s < p.n_slots; ++s)
      if (!(s < dev.n_slots && dev.blocks[s])) return fail(RFB_ERR_STATE, "dev data has no block for slot %d", s);

  const size_t nd = (size_t)max_iters;
  const size_t np = (size_t)std::max(p.P, 1), nk = (size_t)p.F + 1;
  CU(cudaMalloc(&f.step, sizeof(double) * nd));
  CU(cudaMalloc(&f.cost_train, sizeof(double) * nd));
  CU(cudaMalloc(&f.cost_dev, sizeof(double) * nd));
  CU(cudaMalloc(&f.metric_train, sizeof(double) * nd));
  CU(cudaMalloc(&f.metric_dev, sizeof(double) * nd));
  CU(cudaMalloc(&f.ring_b, sizeof(float) * nd * np));
  CU(cudaMalloc(&f.ring_ic, sizeof(double) * nd * nk));
  CU(cudaMalloc(&f.it, sizeof(int)));
  CU(cudaMalloc(&f.m, sizeof(float) * np));
  CU(cudaMalloc(&f.v, sizeof(float) * np));
  CU(cudaMalloc(&f.mi, sizeof(double) * nk));
  CU(cudaMalloc(&f.vi, sizeof(double) * nk));
  CU(cudaMalloc(&f.fin, sizeof(double) * ((size_t)p.NE + kNumScalars)));
  {
    const size_t ns = 2 * 2 * (nd + 1);
    std::vector<unsigned long long> init(ns);
    for (size_t k = 0; k < ns; ++k) init[k] = (k & 1) ? 0ull : ~0ull;       // start: min-folded, end: max-folded
    CU(cudaMalloc(&f.stamps, sizeof(unsigned long long) * ns));
    CU(cudaMemcpyAsync(f.stamps, init.data(), sizeof(unsigned long long) * ns, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  CU(cudaMemcpyAsync(f.step, step_sizes, sizeof(double) * nd, cudaMemcpyHostToDevice, e->stream));
  {
    // 1 - b^(it + 1): in the parameter dtype ((float)b raised in double, rounded once -- jnp.asarray(b1, m.dtype)
    // ** (i + 1) of the published Adam) and in double for the float64 intercept leaves
    std::vector<float> bcf(2 * nd);
    std::vector<double> bcd(2 * nd);
    for (size_t it = 0; it < nd; ++it) {
      bcf[2 * it] = 1.0f - (float)std::pow((double)0.9f, (double)(it + 1));
      bcf[2 * it + 1] = 1.0f - (float)std::pow((double)0.999f, (double)(it + 1));
      bcd[2 * it] = 1.0 - std::pow(0.9, (double)(it + 1));
      bcd[2 * it + 1] = 1.0 - std::pow(0.999, (double)(it + 1));
    }
    CU(cudaMalloc(&f.bc_f, sizeof(float) * 2 * nd));
    CU(cudaMalloc(&f.bc_d, sizeof(double) * 2 * nd));
    CU(cudaMemcpyAsync(f.bc_f, bcf.data(), sizeof(float) * 2 * nd, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(f.bc_d, bcd.data(), sizeof(double) * 2 * nd, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  CU(cudaMemsetAsync(f.it, 0, sizeof(int), e->stream));
  CU(cudaMemsetAsync(f.m, 0, sizeof(float) * np, e->stream));
  CU(cudaMemsetAsync(f.v, 0, sizeof(float) * np, e->stream));
  CU(cudaMemsetAsync(f.mi, 0, sizeof(double) * nk, e->stream));
  CU(cudaMemsetAsync(f.vi, 0, sizeof(double) * nk, e->stream));
  CU(cudaMemsetAsync(f.fin, 0, sizeof(double) * ((size_t)p.NE + kNumScalars), e->stream));
  TRY(fill_nan(f.cost_train, max_iters, e->stream));
  TRY(fill_nan(f.cost_dev, max_iters, e->stream));
  TRY(fill_nan(f.metric_train, max_iters, e->stream));
  TRY(fill_nan(f.metric_dev, max_iters, e->stream));

  // start from the stored initial parameters
  CU(cudaMemcpyAsync(m->cur.theta, m->host_b.data(), sizeof(float) * p.P, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(m->cur.ic, m->host_ic.data(), sizeof(double) * nk, cudaMemcpyHostToDevice, e->stream));
  publish_params_kernel<<<1, kUpdateThreads, 0, e->stream>>>(publish_args(m, m->cur));
  e->launches++;
  CU(cudaGetLastError());

  UpdateArgs& ua = f.ua;
  ua = publish_args(m, m->cur);
  ua.m = f.m; ua.v = f.v; ua.mi = f.mi; ua.vi = f.vi;
  ua.fin_train = f.fin;
  ua.fin_dev = f.has_dev ? f.fin + p.NE : nullptr;
  ua.n_train = tr.n_total; ua.sy_train = tr.sy; ua.syy_train = tr.syy;
  ua.n_dev = dev.n_total; ua.sy_dev = dev.sy; ua.syy_dev = dev.syy;
  ua.distr = m->distr;
  ua.metric = metric;
  ua.alpha = (float)alpha;
  ua.beta = (float)beta;
  ua.penalise = beta != 0.0 ? 1 : 0;
  ua.step_sizes = f.step;
  ua.bc_f = f.bc_f;
  ua.bc_d = f.bc_d;
  ua.cost_train = f.cost_train; ua.cost_dev = f.cost_dev;
  ua.metric_train = f.metric_train; ua.metric_dev = f.metric_dev;
  ua.ring_b = f.ring_b; ua.ring_ic = f.ring_ic;
  ua.it = f.it;
  ua.max_iters = max_iters;
  ua.final_only = 0;

  // size the sweeps' scratch outside the capture (allocation is not capturable)
  TRY(ensure_sweep_buffers(m, RFB_KIND_TRAIN));
  if (f.has_dev) TRY(ensure_sweep_buffers(m, RFB_KIND_DEV));
  if (m->out_nl == RFB_NL_SOFTMAX) {
    TRY(ensure_softmax_buffers(m, RFB_KIND_TRAIN));
    if (f.has_dev) TRY(ensure_softmax_buffers(m, RFB_KIND_DEV));
  }
  TRY(ensure_smx_buffers(m, RFB_KIND_TRAIN));
  if (f.has_dev) TRY(ensure_smx_buffers(m, RFB_KIND_DEV));
  CU(cudaStreamSynchronize(e->stream));

  const int64_t l0 = e->launches;
  CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_iteration(m, false, nullptr, nullptr);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
  f.kernels_per_iter = (int)(e->launches - l0);
  e->launches = l0;  // captured, not executed
  if (rc != RFB_OK || ce != cudaSuccess || !graph) {
    std::string msg = rc != RFB_OK ? g_err : std::string(cudaGetErrorString(ce));
    if (graph) cudaGraphDestroy(graph);
    fit_release(m);
    return fail(rc != RFB_OK ? rc : RFB_ERR_CUDA, "capturing the fit iteration failed: %s", msg.c_str());
  }
  f.graph = graph;
  CU(cudaGraphInstantiate(&f.exec, f.graph, 0));
  TRY(plan_small_fit(m));
  f.active = true;
  f.executed = 0;
  f.finished = false;
  return RFB_OK;
}

extern "C" int rfb_fit_run(rfb_model* m, int n_iters) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  Fit& f = m->fit;
  if (!f.active) return fail(RFB_ERR_STATE, "rfb_fit_begin has not been called");
  if (f.finished) return fail(RFB_ERR_STATE, "the fit was finished; begin a new one");
  if (n_iters < 0 || f.executed + n_iters > f.max_iters)
    return fail(RFB_ERR_INVALID, "%d more iterations exceed max_iters=%d (executed %d)", n_iters, f.max_iters, f.executed);
  TRY(use_device(m->e));
  if (f.small && n_iters > 0) {
    // the whole batch of iterations in one cooperative launch; state stays in the fit's device buffers
    f.sa.n_iters = n_iters;
    f.sa.epoch = f.small_epoch;
    f.small_epoch += (unsigned long long)n_iters;
    void* params[] = {(void*)&f.sa};
    CU(cudaLaunchCooperativeKernel(f.small_fn, dim3(f.small_grid), dim3(kSmallThreads), params, f.small_smem,
                                   m->e->stream));
    m->e->launches += 1;
    f.executed += n_iters;
    return RFB_OK;
  }
  for (int k = 0; k < n_iters; ++k) {
    CU(cudaGraphLaunch(f.exec, m->e->stream));
    m->e->launches += f.kernels_per_iter;
  }
  f.executed += n_iters;
  return RFB_OK;
}

extern "C" int rfb_fit_is_persistent(rfb_model* m, int* out) {
  if (!m || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  *out = (m->fit.active && m->fit.small) ? 1 : 0;
  return RFB_OK;
}

extern "C" int rfb_fit_run_timed(rfb_model* m, int n_iters, double* ms4) {
  if (!m || !ms4) return fail(RFB_ERR_INVALID, "NULL argument");
  Fit& f = m->fit;
  if (!f.active) return fail(RFB_ERR_STATE, "rfb_fit_begin has not been called");
  if (f.finished) return fail(RFB_ERR_STATE, "the fit was finished; begin a new one");
  if (n_iters < 1 || f.executed + n_iters > f.max_iters)
    return fail(RFB_ERR_INVALID, "%d more iterations exceed max_iters=%d (executed %d)", n_iters, f.max_iters, f.executed);
  rfb_engine* e = m->e;
  TRY(use_device(e));
  std::vector<cudaEvent_t> ev((size_t)n_iters * 5);
  for (auto& x : ev) CU(cudaEventCreate(&x));
  int rc = RFB_OK;
  for (int k = 0; k < n_iters && rc == RFB_OK; ++k) rc = enqueue_iteration(m, false, nullptr, nullptr, &ev[(size_t)k * 5]);
  cudaStreamSynchronize(e->stream);
  for (int j = 0; j < 4; ++j) ms4[j] = 0.0;
  if (rc == RFB_OK) {
    for (int k = 0; k < n_iters; ++k)
      for (int j = 0; j < 4; ++j) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[(size_t)k * 5 + j], ev[(size_t)k * 5 + j + 1]);
        ms4[j] += (double)t / n_iters;
      }
    f.executed += n_iters;
  }
  for (auto& x : ev) cudaEventDestroy(x);
  return rc;
}

extern "C" int rfb_fit_sweep_times(rfb_model* m, int kind, int i0, int n, double* ms_out) {
  if (!m || !ms_out) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind != RFB_KIND_TRAIN && kind != RFB_KIND_DEV) return fail(RFB_ERR_INVALID, "bad kind");
  Fit& f = m->fit;
  if (!f.active || !f.stamps) return fail(RFB_ERR_STATE, "no fit in progress");
  if (i0 < 0 || n < 0 || i0 + n > f.executed) return fail(RFB_ERR_INVALID, "iterations [%d, %d) were not all executed (%d were)", i0, i0 + n, f.executed);
  TRY(use_device(m->e));
  std::vector<unsigned long long> st(2 * (size_t)n);
  CU(cudaStreamSynchronize(m->e->stream));
  CU(cudaMemcpy(st.data(), f.stamps + (size_t)kind * 2 * ((size_t)f.max_iters + 1) + 2 * (size_t)i0,
                sizeof(unsigned long long) * st.size(), cudaMemcpyDeviceToHost));
  for (int k = 0; k < n; ++k) {
    const unsigned long long a = st[2 * k], b = st[2 * k + 1];
    ms_out[k] = (a == ~0ull || b < a) ? std::nan("") : (double)(b - a) * 1e-6;
  }
  return RFB_OK;
}

extern "C" int rfb_fit_iterations(rfb_model* m, int* out) {
  if (!m || !out) return fail(RFB_ERR_INVALID, "NULL argument");
  *out = m->fit.executed;
  return RFB_OK;
}

extern "C" int rfb_fit_finish(rfb_model* m, int i_last) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  Fit& f = m->fit;
  if (!f.active) return fail(RFB_ERR_STATE, "rfb_fit_begin has not been called");
  if (i_last < 0 || i_last >= f.executed) return fail(RFB_ERR_INVALID, "iteration %d was not executed (%d were)", i_last, f.executed);
  rfb_engine* e = m->e;
  const Plan& p = m->plan;
  TRY(use_device(e));
  // parameters after update i_last become current
  CU(cudaMemcpyAsync(m->cur.theta, f.ring_b + (size_t)i_last * std::max(p.P, 1), sizeof(float) * p.P,
                     cudaMemcpyDeviceToDevice, e->stream));
  CU(cudaMemcpyAsync(m->cur.ic, f.ring_ic + (size_t)i_last * (p.F + 1), sizeof(double) * (p.F + 1),
                     cudaMemcpyDeviceToDevice, e->stream));
  publish_params_kernel<<<1, kUpdateThreads, 0, e->stream>>>(publish_args(m, m->cur));
  e->launches++;
  CU(cudaGetLastError());
  const int it = i_last + 1;
  CU(cudaMemcpyAsync(f.it, &it, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  DataSet& tr = m->data[RFB_KIND_TRAIN];
  DataSet& dev = m->data[RFB_KIND_DEV];
  if (!tr.r_buf) CU(cudaMalloc(&tr.r_buf, sizeof(float) * r_buf_floats(tr.n)));
  if (f.has_dev && !dev.r_buf) CU(cudaMalloc(&dev.r_buf, sizeof(float) * r_buf_floats(dev.n)));
  TRY(enqueue_iteration(m, true, tr.r_buf, f.has_dev ? dev.r_buf : nullptr));
  CU(cudaStreamSynchronize(e->stream));
  f.finished = true;
  return RFB_OK;
}

extern "C" int rfb_fit_traces(rfb_model* m, int n, double* cost_train, double* cost_dev,
                              double* metric_train, double* metric_dev) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  Fit& f = m->fit;
  if (!f.active) return fail(RFB_ERR_STATE, "no fit in progress");
  if (n < 0 || n > f.executed) return fail(RFB_ERR_INVALID, "asked for %d iterations, %d executed", n, f.executed);
  rfb_engine* e = m->e;
  TRY(use_device(e));
  const size_t bytes = sizeof(double) * (size_t)n;
  if (cost_train) CU(cudaMemcpyAsync(cost_train, f.cost_train, bytes, cudaMemcpyDeviceToHost, e->stream));
  if (cost_dev) CU(cudaMemcpyAsync(cost_dev, f.cost_dev, bytes, cudaMemcpyDeviceToHost, e->stream));
  if (metric_train) CU(cudaMemcpyAsync(metric_train, f.metric_train, bytes, cudaMemcpyDeviceToHost, e->stream));
  if (metric_dev) CU(cudaMemcpyAsync(metric_dev, f.metric_dev, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return RFB_OK;
}

extern "C" int rfb_fit_params_all(rfb_model* m, int n, float* b_all, double* icpt_all) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  Fit& f = m->fit;
  if (!f.active) return fail(RFB_ERR_STATE, "no fit in progress");
  if (n < 0 || n > f.executed) return fail(RFB_ERR_INVALID, "asked for %d iterations, %d executed", n, f.executed);
  rfb_engine* e = m->e;
  const Plan& p = m->plan;
  TRY(use_device(e));
  if (b_all && p.P > 0)
    CU(cudaMemcpyAsync(b_all, f.ring_b, sizeof(float) * (size_t)n * p.P, cudaMemcpyDeviceToHost, e->stream));
  if (icpt_all)
    CU(cudaMemcpyAsync(icpt_all, f.ring_ic, sizeof(double) * (size_t)n * (p.F + 1), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return RFB_OK;
}

extern "C" int rfb_fit_params_at(rfb_model* m, int i, float* b, double* intercepts) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  Fit& f = m->fit;
  if (!f.active) return fail(RFB_ERR_STATE, "no fit in progress");
  if (i < 0 || i >= f.executed) return fail(RFB_ERR_INVALID, "iteration %d was not executed", i);
  rfb_engine* e = m->e;
  const Plan& p = m->plan;
  TRY(use_device(e));
  if (b && p.P > 0)
    CU(cudaMemcpyAsync(b, f.ring_b + (size_t)i * p.P, sizeof(float) * p.P, cudaMemcpyDeviceToHost, e->stream));
  if (intercepts)
    CU(cudaMemcpyAsync(intercepts, f.ring_ic + (size_t)i * (p.F + 1), sizeof(double) * (p.F + 1),
                       cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return RFB_OK;
}

extern "C" int rfb_fit_predictions(rfb_model* m, int kind, float* r_out) {
  if (!m || !r_out) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind != RFB_KIND_TRAIN && kind != RFB_KIND_DEV) return fail(RFB_ERR_INVALID, "bad kind");
  Fit& f = m->fit;
  if (!f.active || !f.finished) return fail(RFB_ERR_STATE, "rfb_fit_finish has not been called");
  DataSet& ds = m->data[kind];
  if (!ds.r_buf) return fail(RFB_ERR_STATE, "no predictions for kind %d", kind);
  TRY(use_device(m->e));
  CU(cudaMemcpyAsync(r_out, ds.r_buf, sizeof(float) * (size_t)ds.n, cudaMemcpyDeviceToHost, m->e->stream));
  CU(cudaStreamSynchronize(m->e->stream));
  return RFB_OK;
}

extern "C" int rfb_fit_predictions_ptr(rfb_model* m, int kind, uint64_t* dptr, int64_t* n) {
  if (!m || !dptr || !n) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind != RFB_KIND_TRAIN && kind != RFB_KIND_DEV) return fail(RFB_ERR_INVALID, "bad kind");
  Fit& f = m->fit;
  if (!f.active || !f.finished) return fail(RFB_ERR_STATE, "rfb_fit_finish has not been called");
  DataSet& ds = m->data[kind];
  if (!ds.r_buf) return fail(RFB_ERR_STATE, "no predictions for kind %d", kind);
  *dptr = (uint64_t)(uintptr_t)ds.r_buf;
  *n = ds.n;
  return RFB_OK;
}

extern "C" int rfb_fit_predictions_swap(rfb_model* m, int kind, rfb_block* b) {
  if (!m || !b) return fail(RFB_ERR_INVALID, "NULL argument");
  if (kind != RFB_KIND_TRAIN && kind != RFB_KIND_DEV) return fail(RFB_ERR_INVALID, "bad kind");
  Fit& f = m->fit;
  if (!f.active || !f.finished) return fail(RFB_ERR_STATE, "rfb_fit_finish has not been called");
  DataSet& ds = m->data[kind];
  if (!ds.r_buf) return fail(RFB_ERR_STATE, "no predictions for kind %d", kind);
  if (b->e != m->e) return fail(RFB_ERR_INVALID, "block belongs to another engine");
  if ((size_t)b->n_rows * (size_t)b->ld != r_buf_floats(ds.n))
    return fail(RFB_ERR_INVALID, "block holds %lld floats, the prediction buffer %lld",
                (long long)(b->n_rows * b->ld), (long long)r_buf_floats(ds.n));
  TRY(use_device(m->e));
  CU(cudaStreamSynchronize(m->e->stream));
  std::swap(ds.r_buf, b->d);
  return RFB_OK;
}

extern "C" int rfb_fit_end(rfb_model* m) {
  if (!m) return fail(RFB_ERR_INVALID, "model is NULL");
  TRY(use_device(m->e));
  CU(cudaStreamSynchronize(m->e->stream));
  // keep the final parameters as the model's current ones
  if (m->fit.active && m->plan.ok) {
    CU(cudaMemcpy(m->host_b.data(), m->cur.theta, sizeof(float) * m->plan.P, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(m->host_ic.data(), m->cur.ic, sizeof(double) * (m->plan.F + 1), cudaMemcpyDeviceToHost));
  }
  return fit_release(m);
}
