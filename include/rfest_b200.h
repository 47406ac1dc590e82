/*
 * rfest_b200 -- C ABI of the B200-native spline-GLM fit engine.
 *
 * The reference (berenslab/RFEst, pure Python + JAX) has no FFI: its boundary for
 * this path is the Python class rfest.GLM (rfest/GLM/_glm.py:24).  This header
 * is the boundary the Python shim `rfest_b200.GLM` binds with ctypes; each
 * entry point cites the reference code it replaces.  Plain pointers and sizes
 * only -- no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RFB_ERR_* code otherwise;
 *     rfb_last_error() returns a message for the calling thread's last failure.
 *   - `on_device` flags say whether a caller pointer is host memory (0) or a
 *     CUDA device pointer valid on the engine's device (1).
 *   - an engine owns one CUDA device, one stream and (optionally) one NCCL
 *     communicator; handles are not thread-safe (one host thread per GPU rank).
 *   - all arithmetic is float32 with float64 accumulation of the cross-sample
 *     sums; intercepts are float64 scalars (the reference's weak-typed Python
 *     floats under jax_enable_x64).
 */
#ifndef RFEST_B200_H
#define RFEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFB_ABI_VERSION 1

/* error codes */
#define RFB_OK 0
#define RFB_ERR_INVALID (-1)     /* bad argument (reference: ValueError)                */
#define RFB_ERR_UNSUPPORTED (-2) /* shape / option not implemented (NotImplementedError) */
#define RFB_ERR_CUDA (-3)        /* CUDA runtime failure                                 */
#define RFB_ERR_NCCL (-4)        /* NCCL failure / NCCL not loadable                     */
#define RFB_ERR_STATE (-5)       /* call out of order (e.g. fit_run before fit_begin)    */

/* nonlinearities: GLM.fnl, rfest/GLM/_glm.py:103-170 */
#define RFB_NL_NONE 0
#define RFB_NL_SOFTPLUS 1    /* log(1 + exp(x)) + 1e-7, the naive form (_glm.py:124-129) */
#define RFB_NL_EXPONENTIAL 2
#define RFB_NL_SIGMOID 3
#define RFB_NL_TANH 4
#define RFB_NL_RELU 5        /* where(x > 0, x, 1e-7) (_glm.py:152-157)                  */
#define RFB_NL_LEAKY_RELU 6
#define RFB_NL_SOFTMAX 7      /* exp(x) / sum over ALL samples of exp(x) (_glm.py:134-140); as the
                                output nonlinearity only: two sweeps per evaluation           */

/* noise distribution: GLM.cost, rfest/GLM/_glm.py:598-603 */
#define RFB_DISTR_GAUSSIAN 0 /* loss_mse, rfest/loss.py:14-16       */
#define RFB_DISTR_POISSON 1  /* loss_neglogli, rfest/loss.py:4-11   */

/* data set kinds: the `kind` argument of add_design_matrix / forwardpass */
#define RFB_KIND_TRAIN 0
#define RFB_KIND_DEV 1
#define RFB_KIND_TEST 2

/* metrics: GLM.compute_score, rfest/GLM/_glm.py:997-1013; rfest/metrics.py:8-20 */
#define RFB_METRIC_NONE 0
#define RFB_METRIC_MSE 1
#define RFB_METRIC_R2 2
#define RFB_METRIC_CORRCOEF 3

/* indices into the scalar sums returned by rfb_model_eval (double[RFB_N_SUMS]) */
#define RFB_SUM_LOSS_A 0  /* poisson: sum(-log(r') y); gaussian: sum((y - r)^2) */
#define RFB_SUM_LOSS_B 1  /* poisson: sum(r');        gaussian: 0               */
#define RFB_SUM_R 2
#define RFB_SUM_RR 3
#define RFB_SUM_YR 4
#define RFB_SUM_SQERR 5
#define RFB_SUM_Y 6
#define RFB_SUM_YY 7
#define RFB_N_SUMS 8

typedef struct rfb_engine rfb_engine;
typedef struct rfb_block rfb_block; /* a projected design block XS_f: float32 [n_rows x n_cols], HBM resident */
typedef struct rfb_model rfb_model;

const char* rfb_last_error(void);
int rfb_abi_version(void);

/* ---- engine -------------------------------------------------------------------------- */
int rfb_engine_create(int device, rfb_engine** out);
int rfb_engine_destroy(rfb_engine* e);
int rfb_engine_synchronize(rfb_engine* e);
/* kernels launched by this engine since creation (bench.py's gpu_launches claim) */
int rfb_engine_launch_count(rfb_engine* e, int64_t* out);
/* Launch-shape knobs of the sweep kernel, applied to models (re)planned afterwards:
 * "sweep_staged" (1 = rows staged through shared memory by cp.async.bulk, 0 = direct loads),
 * "sweep_stages" (ring depth per warp), "sweep_warps" (warps per CTA), "sweep_ctas" (CTAs per
 * SM, 0 = as many as fit), "sweep_sub" (subunit models on the subunit sweep: 3 = default mapping,
 * 0 = off, 1 / 2 = the mappings it was measured against, 4-7 = the pooled / tensor-memory variants).  Also read
 * from RFB_SWEEP_* environment variables at creation.  Further keys: "project_tc" (spatial projection on
 * tcgen05, default 1), "project_fused" (spatial + temporal projection in one kernel, default 0: measured
 * slower), "project_slab_rows", "h2d_overlap" / "h2d_threads" (host-array ring), "fit_persistent" (small
 * models as one persistent kernel, default 1), "gram_mma" (Gram matrices on the FP64 tensor-core
 * instructions, default 1), "xchg" (multi-GPU exchange: 1 = fused over NVLink peer memory, 0 = NCCL,
 * 3 = the one-CTA variant). */
int rfb_engine_set_option(rfb_engine* e, const char* key, int value);
/* the engine's CUDA stream (cudaStream_t as an integer), for event timing by the caller */
int rfb_engine_stream(rfb_engine* e, uint64_t* out);

/* Page-locked host memory for results the caller reads back at DMA speed (predictions). */
int rfb_host_alloc(size_t bytes, void** out);
int rfb_host_free(void* p);

/* Multi-GPU: time samples are sharded across ranks; every fit
 * iteration ends with one all-reduce of the gradient + loss/metric sums.  NCCL is
 * dlopen()ed (libnccl.so.2 -- the copy torch already loaded); rank 0 makes the
 * 128-byte unique id, the caller broadcasts it (torch.distributed), every rank inits. */
int rfb_comm_unique_id(void* out128);
int rfb_engine_comm_init(rfb_engine* e, const void* id128, int rank, int world);
int rfb_engine_comm_info(rfb_engine* e, int* rank, int* world);
/* One-shot exchange over NVLink peer memory, replacing the NCCL call inside the fit iteration:
 * the reduction kernel stores each rank's vector into every peer's buffer (P2P stores) and the
 * update kernel waits on flags and adds the slots in rank order.  create() allocates this
 * rank's buffer and returns its 64-byte CUDA IPC handle; the caller all-gathers the handles
 * (rank order) and passes all of them to connect().  Optional: without it the engine uses NCCL. */
int rfb_engine_xchg_create(rfb_engine* e, void* handle64);
int rfb_engine_xchg_connect(rfb_engine* e, const void* handles);
/* in-place sum all-reduce of a host double vector over the engine's communicator
 * (no-op when there is none); used for the one-time constants n, sum y, sum y^2 */
int rfb_engine_allreduce_f64(rfb_engine* e, double* host_inout, int count);

/* ---- design blocks (K1) ---------------------------------------------------------------- */
/* Zero-filled block of n_rows x n_cols (row stride padded to a multiple of 4 floats). */
int rfb_block_create(rfb_engine* e, int64_t n_rows, int n_cols, rfb_block** out);
int rfb_block_destroy(rfb_block* b);
int rfb_block_shape(const rfb_block* b, int64_t* n_rows, int* n_cols, int64_t* ld);
/* device pointer of the block's storage (row-major, stride ld floats) */
int rfb_block_device_ptr(const rfb_block* b, uint64_t* out);
/* copy rows [row0, row0 + n) x all columns to / from a dense host matrix [n x n_cols] */
int rfb_block_read(rfb_block* b, int64_t row0, int64_t n, float* host_out);
int rfb_block_write(rfb_block* b, int64_t row0, int64_t n, const float* src, int on_device);

/* Fused lag + spline projection.  Replaces
 *     build_design_matrix(X, dims[0], shift)[burn_in:] @ build_spline_matrix(dims, df, smooth)
 * (rfest/utils.py:5-67, rfest/splines.py:213-305, rfest/GLM/_glm.py:250-258,297-298)
 * without ever materialising the lagged matrix.  With the Kronecker structure
 * S = kron(St, Sxy) the product is two separable contractions:
 *     Z[s, q]              = sum_p stim[s, p] * Sxy[p, q]
 *     XS[t, a * n_q + q]   = sum_i Z[t + i - front, q] * St[i, a],   front = nlag + shift - 1
 *                            (0 if nlag + shift <= 0; rows outside [0, n_raw_total) are zero)
 * Fills block rows [dst_row0, dst_row0 + n_out); block row dst_row0 + k is design row
 * (before burn-in trimming) first_t + k.  `stim` holds raw rows [src0, src0 + n_src) of a
 * recording of n_raw_total rows (a shard passes its rows plus the left halo).
 *   St   [nlag x df_t] row-major;  Sxy [n_pix x n_q] row-major (NULL => n_pix == n_q == 1,
 *   i.e. a purely temporal filter).  Output columns [col0, col0 + df_t * n_q).            */
int rfb_block_project(rfb_block* dst, int64_t dst_row0, int64_t n_out, int col0,
                      const float* stim, int on_device, int64_t src0, int64_t n_src,
                      int64_t n_raw_total, int n_pix, int64_t first_t, int nlag, int shift,
                      const float* St, int df_t, const float* Sxy, int n_q);

/* Bytes rfb_block_project moved host -> device in its last call on this engine (0 for device input). */
int rfb_engine_last_project_h2d_bytes(rfb_engine* e, double* out);
/* out[n_rows x k] = block @ B, B [n_cols x k] row-major, both host arrays: what user code writes as
 * `model.XS[kind][name] @ b` (rfest/GLM/_glm.py:562,1127) -- computed on the device, the block is not
 * copied to the host. */
int rfb_block_matmul(rfb_block* b, const float* B, int k, float* out);

/* ---- model: filters, data, parameters --------------------------------------------------- */
/* GLM(distr, output_nonlinearity), rfest/GLM/_glm.py:25 */
int rfb_model_create(rfb_engine* e, int distr, int output_nl, rfb_model** out);
int rfb_model_destroy(rfb_model* m);
/* Remove all filters and parameter state (the data sets stay bound). */
int rfb_model_clear_filters(rfb_model* m);
/* One fitted filter: `n_coef` coefficients applied to columns [0, n_coef) of the block bound
 * to `slot`, followed by `filter_nl` (forwardpass, _glm.py:555-565).  Subunits are several
 * filters on the same slot (_initialize_subunits aliases XS, _glm.py:348).  Filters are
 * numbered in call order = the reference's `filter_names` order. */
int rfb_model_add_filter(rfb_model* m, int slot, int filter_nl, int n_coef);
/* Bind data for one kind: blocks[slot] for slot < n_slots, and the response y[n] (may be
 * NULL for kind test).  n is this rank's number of samples; the global count (MSE divides by
 * it, rfest/loss.py:15), sum y and sum y^2 are all-reduced over the communicator here. */
int rfb_model_set_data(rfb_model* m, int kind, int n_slots, rfb_block* const* blocks,
                       const float* y, int y_on_device, int64_t n);
/* Replace the response of an already bound data set (same n): what a second `fit(y=...)` on the same
 * design needs -- blocks, plan and scratch stay as they are, the device copy of y is reused. */
int rfb_model_set_y(rfb_model* m, int kind, const float* y, int y_on_device, int64_t n);
/* Parameters: b = concatenated filter coefficients (filter order), intercepts = one per
 * filter followed by the global one (p0 of _initialize_p0, _glm.py:369-393). */
int rfb_model_set_params(rfb_model* m, const float* b, const double* intercepts);
int rfb_model_get_params(rfb_model* m, float* b, double* intercepts);
int rfb_model_n_params(rfb_model* m, int* n_coef_total, int* n_intercepts);

/* forwardpass(p, kind) + the sums cost()/compute_score() need (_glm.py:539-614,997-1013).
 * b/intercepts NULL => current parameters.  r_out (host, n floats) may be NULL.
 * sums: double[RFB_N_SUMS], all-reduced over ranks when a communicator is set. */
int rfb_model_eval(rfb_model* m, int kind, const float* b, const double* intercepts,
                   float* r_out, double* sums);

/* cost(p, kind, precomputed=r) (_glm.py:596-603): the loss terms of GIVEN predictions r[n] against the bound
 * response -- loss_ab = {sum -log(r') y, sum r'} (Poisson) or {sum (y - r)^2, 0} (Gaussian), all-reduced. */
int rfb_model_loss_of_predictions(rfb_model* m, int kind, const float* r, int r_on_device, double* loss_ab);

/* value_and_grad(cost) of the data term (_glm.py:692-696 without the penalty): the same sums plus
 * d loss / d b [n_coef_total] and d loss / d intercepts [n_intercepts] (filters, then global), float64. */
int rfb_model_value_and_grad(rfb_model* m, int kind, const float* b, const double* intercepts,
                             double* sums, double* grad_b, double* grad_icpt);

/* Normal-equation pieces for the least-squares initialisation (compute_mle, _glm.py:438-537):
 * over the concatenated, padded design columns of data set `kind` (ctot of them, slot after slot),
 * gram = XS^T XS [ctot x ctot], colsum = XS^T 1 [ctot], xty = XS^T y [ctot], all float64 and
 * all-reduced over ranks.  Any output may be NULL; with all three NULL only *ctot is set. */
int rfb_model_gram(rfb_model* m, int kind, double* gram, double* colsum, double* xty, int* ctot);

/* XS_f^T diag(U) XS_f [n_coef x n_coef] of one filter (float64, all-reduced), the matrix whose inverse
 * is the coefficient covariance of _get_filter_variance (_glm.py:1058-1164): weighted = 0 gives the
 * plain Gram (Gaussian case); weighted = 1 uses U = 1 / fnl(fnl(XS_f b_f, filter_nl), output_nl)^2 with
 * `b` the full coefficient vector (no intercepts, as in the reference). */
int rfb_model_filter_gram(rfb_model* m, int kind, int filter, const float* b, int weighted, double* gram);

/* Confidence band of the predicted response (_get_response_variance, _glm.py:1166-1232).  For every
 * filter f with data in `kind`: y_se_f[t] = sqrt(xs_t^T V_f xs_t) with V[f] the filter's coefficient
 * covariance (host, float32 [n_coef x n_coef] row-major; NULL entry = width 0), then
 *   pred  = out_nl(sum_f filter_nl(xs_t b_f + c_f)             + c_global)
 *   upper = out_nl(sum_f filter_nl(xs_t b_f + c_f + 2 y_se_f)  + c_global)     lower: - 2 y_se_f
 * (the reference's X_f w_f of the band lines equals XS_f b_f because w_f = S_f b_f, so the lagged
 * matrix is not needed).  V has one entry per filter of the model.  pred / upper / lower receive this
 * rank's n samples each: host pointers, or device pointers when on_device != 0. */
int rfb_model_response_band(rfb_model* m, int kind, const float* b, const double* intercepts,
                            const float* const* V, float* pred, float* upper, float* lower, int on_device);

/* ---- fit loop (GLM.optimize, _glm.py:644-836) --------------------------------------------- */
/* Prepare `max_iters` iterations from the current parameters with zeroed Adam state:
 *   step_sizes[max_iters]  per-iteration learning rate (a constant or a schedule evaluated
 *                          by the caller, _glm.py:669)
 *   alpha, beta            elastic net (rfest/loss.py:19-23); beta == 0 => no penalty
 *   metric                 RFB_METRIC_*
 * Captures one iteration (train sweep, dev sweep, reduction, all-reduce, penalty + Adam)
 * as a CUDA graph. */
int rfb_fit_begin(rfb_model* m, int max_iters, const double* step_sizes, double alpha,
                  double beta, int metric);
/* Run `n_iters` more iterations (graph replays; no host synchronisation inside). */
int rfb_fit_run(rfb_model* m, int n_iters);
/* Same iterations launched eagerly (outside the graph) with CUDA events on the engine's stream
 * around each kernel: ms4 = mean milliseconds of {train sweep, dev sweep, reduction + all-reduce,
 * update}.  For bench.py's per-kernel roofline. */
int rfb_fit_run_timed(rfb_model* m, int n_iters, double* ms4);
/* 1 when the fit in progress runs as the persistent small-model kernel: data sets that fit the SMs'
 * shared memory (launch-latency regime, e.g. the reference README recipe) run every batch of
 * rfb_fit_run iterations as ONE cooperative kernel with grid barriers instead of a graph per iteration
 * (engine option "fit_persistent": 0 off, 1 auto, 2 whenever the data fit). */
int rfb_fit_is_persistent(rfb_model* m, int* out);
/* Duration (ms) of the train / dev sweep kernel of iterations [i0, i0 + n), measured INSIDE the
 * replayed graph: every CTA of a sweep folds %globaltimer into a per-iteration (first entry, last exit)
 * pair, so these are the kernel times of the very iterations rfb_fit_run executed (no events, no
 * eager relaunch).  NaN where no sweep ran. */
int rfb_fit_sweep_times(rfb_model* m, int kind, int i0, int n, double* ms_out);
/* Number of iterations executed so far. */
int rfb_fit_iterations(rfb_model* m, int* out);
/* Make iteration `i_last` (< executed) the final one: its parameters become current, the
 * forward passes that complete trace entry i_last are run, predictions are kept. */
int rfb_fit_finish(rfb_model* m, int i_last);
/* Traces (double[n] each, n <= executed iterations; entries as in the reference:
 * cost_train[i] before update i, the others after it; NaN where not computed). */
int rfb_fit_traces(rfb_model* m, int n, double* cost_train, double* cost_dev,
                   double* metric_train, double* metric_dev);
/* Parameters after update `i` (all_params[i], _glm.py:724,828-830). */
int rfb_fit_params_at(rfb_model* m, int i, float* b, double* intercepts);
/* all iterations [0, n) at once: b_all [n x n_coef_total], icpt_all [n x n_intercepts] */
int rfb_fit_params_all(rfb_model* m, int n, float* b_all, double* icpt_all);
/* Predictions of the final forward pass of rfb_fit_finish (y_pred['opt'][kind]). */
int rfb_fit_predictions(rfb_model* m, int kind, float* r_out);
/* The same predictions without leaving the device: pointer to n floats (padded to a multiple of 4),
 * valid until the next forward pass over that data set. */
int rfb_fit_predictions_ptr(rfb_model* m, int kind, uint64_t* dptr, int64_t* n);
/* Zero-copy hand-over: trade the storage of block `b` ([ceil((n + 4) / 4) x 4] floats) with the prediction
 * buffer of `kind`, so that `b` holds the predictions and the model keeps b's old storage for the next
 * forward pass (y_pred['opt'][kind] of _glm.py:832-834 without an allocation or a copy per fit). */
int rfb_fit_predictions_swap(rfb_model* m, int kind, rfb_block* b);
/* Release graph / traces. */
int rfb_fit_end(rfb_model* m);

#ifdef __cplusplus
}
#endif
#endif /* RFEST_B200_H */
