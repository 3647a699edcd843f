/*
 * neuralfoil_b200 -- C ABI of the B200 (sm_100a) evaluator for NeuralFoil's learned core.
 *
 * The reference (peterdsharpe/NeuralFoil v0.3.3) has no FFI layer: its boundary for this path is the
 * Python callable `neuralfoil.get_aero_from_kulfan_parameters` (reference neuralfoil/main.py:64-332).
 * This header declares what a ctypes stub placed behind that callable binds to; INTEGRATION.md shows
 * the stub.  Nothing here uses torch or C++ types: plain pointers, sizes and integer codes.
 *
 * Every entry point returns NFB_OK (0) or a negative error code; `nfb_last_error()` gives the
 * message of the last failure on the calling thread.
 *
 * Layout conventions ("SoA", the reference's effective layout, main.py:240):
 *   inputs : 23 raw rows, row r = one scalar per query.  Row order
 *            upper_weights[0..7], lower_weights[0..7], leading_edge_weight, TE_thickness,
 *            alpha [deg], Re, n_crit, xtr_upper, xtr_lower          (main.py:171-184)
 *            Each row has its own base pointer and element stride; stride 0 broadcasts one value
 *            (the reference's `np.ones(N_cases) * row`, main.py:198-199).
 *   outputs: 198 rows x N in the reference's dict order (main.py:318-331): analysis_confidence, CL,
 *            CD, CM, Top_Xtr, Bot_Xtr, upper_bl_theta_0..31, upper_bl_H_0..31, upper_bl_ue/vinf_0..31,
 *            lower_bl_theta_*, lower_bl_H_*, lower_bl_ue/vinf_*.  Row r starts at out + r * out_ld
 *            elements; out_ld >= N.
 */
#ifndef NEURALFOIL_B200_H
#define NEURALFOIL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFB_N_RAW_ROWS 23
#define NFB_N_LATENT_INPUTS 25
#define NFB_N_OUTPUT_ROWS 198
#define NFB_N_BL_STATIONS 32
#define NFB_MAX_MODELS 16
#define NFB_MAX_LAYERS 16

enum nfb_status {
  NFB_OK = 0,
  NFB_ERR_INVALID_ARGUMENT = -1, /* bad pointer / size / id; maps to ValueError on the Python side   */
  NFB_ERR_UNSUPPORTED_MODEL = -2,/* layer shapes this build has no kernel for (main.py:206-215 analogue) */
  NFB_ERR_NOT_LOADED = -3,       /* model slot or distribution not set before nfb_eval_*               */
  NFB_ERR_CUDA = -4,             /* CUDA runtime failure; message carries cudaGetErrorString           */
  NFB_ERR_NO_DEVICE = -5         /* no sm_100 device visible; there is NO CPU fallback                  */
};

enum nfb_dtype { NFB_F32 = 0, NFB_F64 = 1 };

/* Which kernel evaluates the network.
 *   FP32_FFMA    fp32 CUDA-core path, every model size; the correctness reference (tolerance: tests/parity.py).
 *   TENSOR_FP16  tcgen05 tensor-core path for 512- and 256-wide models (xxxlarge, xxlarge): fp16 operands (TF32's
 *                mantissa), fp32 accumulation in tensor memory, first layer fp32-exact; looser, stated tolerance.
 *                Other models: NFB_ERR_UNSUPPORTED_MODEL.
 *   TENSOR_FP16X3  the same tcgen05 kernel with every operand split into two fp16 halves and three MMAs per
 *                product (a_hi w_hi + a_lo w_hi + a_hi w_lo), fp32 accumulation: fp32-class accuracy, held to the
 *                FP32_FFMA tolerance.  Same model restriction. */
enum nfb_variant { NFB_VARIANT_FP32_FFMA = 0, NFB_VARIANT_TENSOR_FP16 = 1, NFB_VARIANT_TENSOR_FP16X3 = 2 };

/* May be OR-ed into `variant`: write only the six scalar outputs (analysis_confidence, CL, CD, CM, Top_Xtr, Bot_Xtr --
 * rows 0..5 of the full block, bit-identical to them).  `out` is then a 6 x out_ld block: 24 instead of 792 bytes per
 * query travel back over PCIe in nfb_eval_host, which is what bounds every model but the largest there. */
#define NFB_OUTPUT_SCALARS 0x100

/* Cross-check kernels, for tests and A/B measurements only; may be OR-ed into `variant`.  Both are complete, independent
 * implementations of the same arithmetic that the parity tests hold against the production kernels.
 *   NFB_DEBUG_FFMA_TILED  FP32_FFMA: the CTA-synchronous first implementation (nfb_ffma.cuh), for every batch size;
 *                         must agree with the column-group kernels bit for bit.
 *   NFB_DEBUG_TC_ONE_SM   TENSOR_*: the one-SM (cta_group::1) tcgen05 kernels (nfb_tc.cuh): same operands, another K order,
 *                         so they agree with the CTA-pair kernels to the variant's tolerance, not bit for bit. */
#define NFB_DEBUG_FFMA_TILED 0x1000
#define NFB_DEBUG_TC_ONE_SM 0x2000
#define NFB_VARIANT_FLAGS (NFB_OUTPUT_SCALARS | NFB_DEBUG_FFMA_TILED | NFB_DEBUG_TC_ONE_SM)

#define NFB_VERSION "0.5.0"

typedef struct nfb_context nfb_context;
typedef struct nfb_multi nfb_multi;

/* Threading contract.  A context serialises its own entry points with an internal mutex: any number of host threads may
 * call into the same context, the *_host calls then run one after the other (they share the context's staging buffers and
 * streams), the *_device calls hold the lock only while they enqueue.  The reverse-mode kernels (nfb_eval_scalar_grad_device,
 * nfb_eval_vjp_reverse_device) share one scratch block per context: launches on different streams of one context are chained
 * by an event, i.e. they are correct but do not overlap -- use one context per stream for concurrency.  Different contexts
 * never share state (the distribution statistics travel in each launch's parameters), so one context per thread, per stream
 * or per device needs no coordination at all.  The first reverse-mode call for a larger model or batch than any before
 * allocates (cudaMalloc): warm the context up with one eager call of the same shape before capturing a CUDA graph. */

/* Library / build identification (static string, never NULL). */
const char* nfb_version(void);

/* Message of the last error raised on this thread (static storage, never NULL). */
const char* nfb_last_error(void);

/* Create an evaluator bound to CUDA device `device` (its primary context).  Replaces the reference's
 * import-time state (main.py:25-44).  Fails with NFB_ERR_NO_DEVICE when no sm_100 GPU is present. */
int nfb_create(int device, nfb_context** out_ctx);
void nfb_destroy(nfb_context* ctx);

/* Tuning knobs for experiments and tests (the defaults are the measured optima):
 *   NFB_TUNE_SMALL_BATCH_MAX  batches of at most `value` queries use the small-batch kernel (-1 = default per width)
 *   NFB_TUNE_TC2_MAX_GRID     at most `value` CTAs (rounded down to whole pairs) for the CTA-pair tensor kernels (0 = all SMs)
 *   NFB_TUNE_HOST_THREADS     at most `value` host threads stage pageable caller memory in nfb_eval_host (0 = all cores, max 32)
 *   NFB_TUNE_MAPPED_MAX       nfb_eval_host: batches of at most `value` queries run zero-copy -- the kernel reads and writes the
 *                             pinned staging block directly, one launch and one synchronisation (-1 = default 8, 0 = never) */
enum nfb_tuning { NFB_TUNE_SMALL_BATCH_MAX = 0, NFB_TUNE_TC2_MAX_GRID = 1, NFB_TUNE_HOST_THREADS = 2, NFB_TUNE_MAPPED_MAX = 3 };
int nfb_set_tuning(nfb_context* ctx, int key, int64_t value);

/* Number of SMs / device ordinal the context runs on. */
int nfb_device_info(const nfb_context* ctx, int* device, int* sm_count, int* cc_major, int* cc_minor);

/* Mean (25) and inverse covariance (25x25, row-major) of the scaled inputs, float64.
 * Replaces `_scaled_input_distribution` (main.py:27-32) as used by `_squared_mahalanobis_distance`
 * (main.py:47-61). */
int nfb_set_distribution(nfb_context* ctx, const double* mean25, const double* inv_cov_25x25);

/* Load one network into slot `model_id` (0 <= id < NFB_MAX_MODELS).
 *   n_layers       number of affine layers L (3..7 for the shipped sizes)
 *   dims[L + 1]    25, hidden..., 198
 *   weights[l]     float32, shape (dims[l+1], dims[l]) row-major -- exactly `net.{2l}.weight` of the
 *                  reference's .npz (training/convert_all_pth_files_to_npz.py:16-19)
 *   biases[l]      float32, shape (dims[l+1],)
 * Host pointers; the library repacks them into its own device layout.  Replaces
 * `_nn_parameters[model_size]` (main.py:41-44) and the layer discovery at main.py:206-215. */
int nfb_load_model(nfb_context* ctx, int model_id, int n_layers, const int* dims,
                   const float* const* weights, const float* const* biases);

/* Evaluate N queries.  All data pointers are DEVICE pointers valid on the context's device.
 *   in_rows[23], in_strides[23]  per-row base pointer and element stride (0 or 1, or any >= 0)
 *   in_dtype / out_dtype         NFB_F32 or NFB_F64 element type of the rows
 *   out, out_ld                  198 x out_ld block, row-major
 *   stream                       a cudaStream_t (NULL = default stream); the call only enqueues work
 * Replaces the body of get_aero_from_kulfan_parameters from featurisation to decode
 * (main.py:171-184, 218-316). */
int nfb_eval_device(nfb_context* ctx, int model_id, int variant, int64_t n,
                    const void* const* in_rows, const int64_t* in_strides, int in_dtype,
                    void* out, int64_t out_ld, int out_dtype, void* stream);

/* Same, with HOST pointers (pageable or pinned).  The library moves data in chunks of `chunk`
 * queries (0 = default) over two streams so that host->device copies, the kernel and
 * device->host copies overlap, and returns after the last byte has landed in `out`. */
int nfb_eval_host(nfb_context* ctx, int model_id, int variant, int64_t n,
                  const void* const* in_rows, const int64_t* in_strides, int in_dtype,
                  void* out, int64_t out_ld, int out_dtype, int64_t chunk);

/* Mixed-precision outputs: as nfb_eval_device / nfb_eval_host, but the 192 boundary-layer rows (theta, H, ue/Vinf of both
 * surfaces: output rows 6..197) are written as bfloat16 -- the fp32 results rounded to nearest, relative error <= 2^-9 -- and
 * the six scalars stay float32: 408 instead of 792 bytes per query cross PCIe, which is what bounds the host entry point for
 * every model but the largest on the fp32 kernel.  Any variant (not NFB_OUTPUT_SCALARS).
 *   out_scalars   6 x out_ld float32 (rows 0..5, bit-identical to nfb_eval_*)
 *   out_bl        192 x out_ld bfloat16 (row r - 6 = output row r); bit-identical to the bfloat16 rounding of nfb_eval_*'s rows
 * nfb_eval_host_mixed copies straight between the caller's memory and the device, chunk by chunk on two streams: give it
 * pinned memory (nfb_host_alloc / nfb_host_register); pageable memory works at cudaMemcpy's pageable rate. */
int nfb_eval_device_mixed(nfb_context* ctx, int model_id, int variant, int64_t n,
                          const void* const* in_rows, const int64_t* in_strides, int in_dtype,
                          float* out_scalars, void* out_bl, int64_t out_ld, void* stream);
int nfb_eval_host_mixed(nfb_context* ctx, int model_id, int variant, int64_t n,
                        const void* const* in_rows, const int64_t* in_strides, int in_dtype,
                        float* out_scalars, void* out_bl, int64_t out_ld, int64_t chunk);

/* Outputs AND their derivatives with respect to the 23 raw inputs (forward mode, fp32 arithmetic, every model size).
 * what a design optimiser needs from the path; the reference's users obtain it by tracing main.py
 * through CasADi (studies/sequential_multifidelity_optimization.py:21-56, benchmarking/daedalus_optimization.py:14-60).
 *   values, values_ld   198 x values_ld block as nfb_eval_device's `out` (may be NULL); bit-identical to variant 0
 *   jac                 n x 198 x 23, contiguous: jac[q][row][t] = d output_row / d raw_input_t of query q, inputs in the
 *                       row order of in_rows (upper 0..7, lower 8..15, LE 16, TE 17, alpha 18 [per degree], Re 19, n_crit 20,
 *                       xtr_upper 21, xtr_lower 22)
 * Device pointers / host pointers as for the two calls above. */
int nfb_eval_jacobian_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                             const int64_t* in_strides, int in_dtype, void* values, int64_t values_ld,
                             void* jac, int out_dtype, void* stream);
int nfb_eval_jacobian_host(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                           const int64_t* in_strides, int in_dtype, void* values, int64_t values_ld,
                           void* jac, int out_dtype);

/* Vector-Jacobian product (reverse-mode result, computed by the same forward-mode kernel without ever writing the
 * Jacobian): grad[t][q] = sum_row cot[row][q] * d output_row / d raw_input_t -- the gradient with respect to the 23 raw
 * inputs of any scalar objective whose derivative with respect to the 198 outputs is `cot`.  92 instead of 18,216 bytes
 * per query leave the kernel.
 *   cot, cot_ld     198 x cot_ld cotangents (type out_dtype);  values, values_ld as above (may be NULL);
 *   grad, grad_ld   23 x grad_ld gradients (type out_dtype), rows in the order of in_rows. */
int nfb_eval_vjp_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                        const int64_t* in_strides, int in_dtype, const void* cot, int64_t cot_ld,
                        void* values, int64_t values_ld, void* grad, int64_t grad_ld, int out_dtype, void* stream);

/* Reverse-mode gradient of the six scalar outputs for large batches (nfb_grad.cuh): as nfb_eval_vjp_device with
 * cotangents for rows 0..5 only (analysis_confidence, CL, CD, CM, Top_Xtr, Bot_Xtr) -- the gradient, with respect to the 23
 * raw inputs, of any scalar objective of the aerodynamic scalars, at the cost of about 1.6 forward passes per query instead
 * of the forward-mode kernel's 24.  What a reverse-mode trace of main.py:171-303 returns for such an objective.
 *   cot, cot_ld     6 x cot_ld cotangents (type out_dtype);  values, values_ld: 6 x values_ld outputs (may be NULL);
 *   grad, grad_ld   23 x grad_ld gradients (type out_dtype), rows in the order of in_rows. */
int nfb_eval_scalar_grad_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                                const int64_t* in_strides, int in_dtype, const void* cot, int64_t cot_ld,
                                void* values, int64_t values_ld, void* grad, int64_t grad_ld, int out_dtype, void* stream);

/* nfb_eval_vjp_device's result -- cotangents of all 198 outputs -- from the fused reverse pass (nfb_grad.cuh, FULL): same
 * arguments, same meaning; about 3 forward passes of arithmetic per query instead of 24.  The kernel of choice from a few
 * hundred queries up; the forward-mode kernel remains the one for an optimiser's single query. */
int nfb_eval_vjp_reverse_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                                const int64_t* in_strides, int in_dtype, const void* cot, int64_t cot_ld,
                                void* values, int64_t values_ld, void* grad, int64_t grad_ld, int out_dtype, void* stream);

/* One training step's forward, loss and backward: what the reference's loop computes between `net(x)` and
 * `loss.backward()` (training/train.py:56-114 Net.forward, :194-236 loss_function, :251-255), for the network in slot
 * `model_id`, in fp32 on the GPU.  All data pointers are DEVICE pointers.
 *   x, x_ld             (batch, 25) float32 scaled inputs, row b at x + b * x_ld            [train.py:251]
 *   y_data, y_ld        (batch, 198) float32 targets, NaN = no data for that output         [train.py:252, :197]
 *   loss_weights        198 float32 (train.py:180-190)          huber_delta   0.05 in the reference (train.py:213)
 *   loss_components     198 float64 OUT: the weighted mean loss of every output; their sum is `loss` (train.py:222-230)
 *   grad_weights[l]     OUT, float32 (dims[l+1], dims[l]) row-major = d loss / d net.{2l}.weight
 *   grad_biases[l]      OUT, float32 (dims[l+1],)                    = d loss / d net.{2l}.bias
 *                       (host arrays of n_layers device pointers, the layout nfb_load_model takes the parameters in)
 * Enqueues three kernels on `stream`; a context-owned workspace grows with the batch (49 KB per sample for xxxlarge, 10 KB
 * for xlarge).  The gradients are bit-reproducible from run to run; loss_components is summed with floating-point atomics
 * in double precision (reproducible to ~1e-16 relative).  The optimiser step stays with the caller, who re-uploads the
 * updated parameters with nfb_load_model. */
int nfb_train_step_device(nfb_context* ctx, int model_id, int64_t batch, const float* x, int64_t x_ld,
                          const float* y_data, int64_t y_ld, const float* loss_weights, float huber_delta,
                          double* loss_components, float* const* grad_weights, float* const* grad_biases, void* stream);

/* ---- several GPUs of one box behind one call (BASELINE configs[2] and [4]) ---------------------------------
 * Queries are independent (main.py:198-201: every row of the stacked input is its own case) and the weights are a few MB,
 * so the path shards by contiguous index range with NO collective: device i of G evaluates queries
 * [i * ceil4(n / G), ...) -- nfb_shard_range -- with the replicated model and copies its results device->host straight
 * into its own column range of the caller's single 198 x out_ld block.  The "host gather of outputs" is therefore free:
 * when the call returns the block is complete.  One host thread per device drives that device's nfb_eval_host (its own
 * streams and pinned staging buffers), pinned to the CPUs the device is attached to (sysfs local_cpulist) so that staging
 * copies and first-touch page placement stay on the device's NUMA node.
 *
 * nfb_multi_create      `devices` = n_devices CUDA ordinals (NULL = every visible sm_100 device, n_devices ignored).
 * nfb_multi_context     the i-th device's context (borrowed: do not destroy) for device-resident work on that GPU.
 * nfb_multi_set_distribution / nfb_multi_load_model   as the single-device calls, applied to every device.
 * nfb_multi_eval_host   as nfb_eval_host; bit-identical to the single-device result (a query's arithmetic does not depend
 *                       on its position in a batch), for any n -- including n smaller than the number of devices.
 * nfb_shard_range       the rule itself: [*begin, *end) of shard `index` of `count` over n queries; shard boundaries are
 *                       multiples of 4 queries so that every shard's rows stay 16-byte aligned. */
int nfb_multi_create(const int* devices, int n_devices, nfb_multi** out_multi);
void nfb_multi_destroy(nfb_multi* multi);
int nfb_multi_device_count(const nfb_multi* multi);
nfb_context* nfb_multi_context(nfb_multi* multi, int index);
int nfb_multi_set_distribution(nfb_multi* multi, const double* mean25, const double* inv_cov_25x25);
int nfb_multi_load_model(nfb_multi* multi, int model_id, int n_layers, const int* dims,
                         const float* const* weights, const float* const* biases);
int nfb_multi_eval_host(nfb_multi* multi, int model_id, int variant, int64_t n,
                        const void* const* in_rows, const int64_t* in_strides, int in_dtype,
                        void* out, int64_t out_ld, int out_dtype, int64_t chunk);
int nfb_shard_range(int64_t n, int index, int count, int64_t* begin, int64_t* end);
int64_t nfb_multi_launch_count(const nfb_multi* multi);

/* Pinned host memory helpers for callers without their own allocator. */
int nfb_host_alloc(void** out_ptr, int64_t bytes);
int nfb_host_free(void* ptr);
/* Page-lock memory the caller already owns (cudaHostRegister, portable to every device) -- e.g. a POSIX shared-memory
 * block that the ranks of a one-process-per-GPU job all map: each rank then copies its shard device->host straight into
 * rank 0's result block (neuralfoil_b200/sharding.py: SharedHostBlock). */
int nfb_host_register(void* ptr, int64_t bytes);
int nfb_host_unregister(void* ptr);

/* Number of kernel launches this context has enqueued so far (bench.py's `gpu_launches`). */
int64_t nfb_launch_count(const nfb_context* ctx);

/* Developer aid: when enabled, CTA 0 of the tensor-core kernel records clock64 stamps of its second tile's phases
 * (featurise, each layer's MMA span, each epilogue) into a 64-entry buffer; `out64` (may be NULL) receives a copy. */
int nfb_debug_tc_stamps(nfb_context* ctx, int enable, long long* out64);

/* Measure the device's fp32 FFMA throughput with a register-resident outer-product loop
 * (the roofline denominator for the FFMA variant; MEASURED_PEAKS.json has no fp32 figure).
 * Runs `iters` timed launches and returns the best rate in TFLOP/s. */
int nfb_measure_fp32_peak(nfb_context* ctx, int iters, double* out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* NEURALFOIL_B200_H */
