/* alan_b200.h -- C ABI of the B200-native logPQ engine (libalan_b200.so).
 *
 * The reference (alan-ppl/alan) is pure Python/PyTorch and has NO FFI layer
 * (SURVEY.md §8b): the seam this library plugs into is the Python call from
 * Sample._elbo / Sample._importance_sample_idxs into logPQ_plate / logPQ_sample,
 * selected by the `computation_strategy` object.  Each entry point below names
 * the reference interface it replaces (paths relative to /root/reference/).
 *
 * Rules of the boundary:
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer;
 *   - the library never allocates device memory: the caller (torch) owns inputs,
 *     outputs and the workspace (size from alan_b200_workspace_bytes);
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*);
 *   - return value 0 = success, non-zero = error, message in alan_b200_last_error();
 *   - a plan is immutable after creation; one plan may be used from several host
 *     threads (the autograd thread differs from the forward thread) as long as
 *     each (workspace, stream) pair is used by one thread at a time;
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef ALAN_B200_H
#define ALAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct alan_b200_plan alan_b200_plan;

/* ABI version of this header / library. */
int alan_b200_abi_version(void);

/* Last error message of the calling thread ("" if none). */
const char* alan_b200_last_error(void);

/* Build an executable plan from the int32 blob emitted by alan_b200/plan.py.
 * The blob is the static description of one model + shapes: the factor programs,
 * the contraction steps chosen by the planner, plate sums / Timeseries chains and
 * the adjoint and resampling programs.
 * replaces: the recursive Python walk of logPQ_plate / lp_getter
 *           (src/alan/logpq.py:15-155, 257-332), done once instead of per call. */
int alan_b200_plan_create(const int32_t* blob, size_t n_words, alan_b200_plan** plan);
void alan_b200_plan_destroy(alan_b200_plan* plan);

/* Plate sharding across GPUs: plans built with fused collectives reduce the per-shard [K_parents]
 * tile and the global-parameter gradients ACROSS RANKS inside their programs, through symmetric buffers that every
 * rank maps into its address space (NVLink peer memory; torch.distributed._symmetric_memory allocates and
 * exchanges them).  alan_b200_comm_bytes: size each rank's zero-initialised buffer must have (0: the plan has no
 * such reduction).  alan_b200_plan_set_comm: peer_buffers[r] = rank r's buffer as mapped in THIS process.
 * replaces: the sequential `prev_lpq + lp` accumulation over Split chunks (src/alan/logpq.py:43-57,151-153). */
size_t alan_b200_comm_bytes(const alan_b200_plan* plan);
int alan_b200_plan_set_comm(alan_b200_plan* plan, int rank, int world, void* const* peer_buffers, size_t bytes);

/* Bytes of device workspace the plan needs (factors, adjoints, partials). */
size_t alan_b200_workspace_bytes(const alan_b200_plan* plan);
int alan_b200_num_inputs(const alan_b200_plan* plan);
int alan_b200_num_programs(const alan_b200_plan* plan);
/* Kernels one run of `program` launches (for gpu_launches accounting). */
int alan_b200_program_launches(const alan_b200_plan* plan, int program);

/* Run one program of the plan.  Programs are numbered by the planner; the
 * Python host (alan_b200/runtime.py) knows which is which.  `inputs` holds
 * n_inputs device pointers in plan order, `outputs` the program's outputs.
 * Generic entry used by the three typed entry points below. */
int alan_b200_run(const alan_b200_plan* plan, int program, const void* const* inputs,
                  void* const* outputs, void* workspace, void* stream);

/* Measurement aid: run `program` with a CUDA event recorded on `stream` before and after every
 * op, synchronise, and return the per-op device times in milliseconds (returns the number of
 * ops, or -1 on error).  Used by bench.py for the roofline of the dominant kernel; never on
 * the timed path. */
int alan_b200_profile(const alan_b200_plan* plan, int program, const void* const* inputs,
                      void* const* outputs, const void* const* aux, void* workspace, void* stream,
                      float* ms_per_op, int max_ops);

/* Forward: log-evidence estimate (0-d, plan dtype) written to lp_out.
 * replaces: logPQ_plate(name=None, ...) -> lp      (src/alan/logpq.py:15-60)
 *   factor evaluation   logPQ_gdt / Dist.log_prob   (logpq.py:157-254, dist.py:297-302,
 *                                                    TorchDimDist.py:127-162)
 *   mixture-Q reduction SamplerMP.reduce_logQ       (Sampler.py:118-134)
 *   K contraction       reduce_Ks / logsumexp_dims  (reduce_Ks.py:236-298, utils.py:207-222)
 *   plate sum / chain   lp.sum / chain_logmmexp     (logpq.py:131-153, utils.py:478-510)
 * `segment` selects the part of the forward program before (0) or after (1) the
 * cross-GPU all-reduce of the sharded-plate tile; single-GPU plans have one segment. */
int alan_b200_logpq_fwd(const alan_b200_plan* plan, int segment, const void* const* inputs,
                        void* lp_out, void* workspace, void* stream);

/* Backward of the same path: gradients of lp w.r.t. every input the plan marks
 * as differentiable (distribution arguments, reparameterised samples, source
 * terms J whose gradients are the marginals / moments), written to grads_out
 * in plan order.  grad_lp is a device pointer to the upstream scalar gradient.
 * replaces: torch.autograd over logPQ_plate incl. checkpoint recomputation
 *           (src/alan/logpq.py:62-66; Sample.py:257-272, 334-346). */
int alan_b200_logpq_bwd(const alan_b200_plan* plan, int segment, const void* const* inputs,
                        const void* grad_lp, void* const* grads_out, void* workspace, void* stream);

/* Posterior resampling of K indices, top-down over the plate tree, from the
 * factors the forward pass left in `workspace`.  `uniforms[i]` is the float64
 * tensor u[batch plates..., N] consumed by sampling step i; `idx_out[g]` receives
 * int64 indices [N, plates of group g...] for latent group g (plan order).
 * replaces: logPQ_sample / sample_Ks (src/alan/sample_logpq.py:17-107,
 *           src/alan/reduce_Ks.py:35-83, unravel_index.py:24-101). */
int alan_b200_resample(const alan_b200_plan* plan, const void* const* inputs,
                       const double* const* uniforms, int64_t* const* idx_out,
                       void* workspace, void* stream);

/* Gather samples at resampled indices: out[n, plates..., event] = x[idx[n, plates_g], plates..., event].
 * x has layout [outer, K, inner] (K stride = inner); idx is [N, outer_g] and is broadcast
 * over the plates of x that the group does not carry via (outer_div): idx row = outer / outer_div.
 * replaces: index_into_sample (src/alan/Sample.py:359-381). */
int alan_b200_gather(const void* x, const int64_t* idx, void* out, int elem_bytes,
                     int64_t N, int64_t outer, int64_t K, int64_t inner, int64_t outer_div,
                     void* stream);

/* Byte-typed inputs: dst[i] = (working dtype) src[i] for n uint8 / bool elements (both pointers 16-byte aligned).
 * Binary features, 0/1 observations and small counts cross PCIe as bytes (a quarter of the fp32 traffic) and are widened
 * on the device; the values are exact in either type.
 * replaces: the float copies of the binary covariates / observations that the reference loads and moves to the
 * device (examples/models/movielens/movielens.py:11-22, `prob.to(device)` at :98).  dtype: 0 = f32, 1 = f64. */
int alan_b200_widen_u8(const void* src, void* dst, int64_t n, int dtype, void* stream);

/* QEM parameter update of ONE latent variable, elementwise and in place:
 *     mean_s  <- mean_s * (1 - lr) + lr * new_s            (s = 0, 1: the family's sufficient statistics)
 *     params  <- mean2conv(mean_0, mean_1)                  (conventional parameters of the family)
 * family: 0 Normal (mean, mean2 -> loc, scale), 1 Bernoulli (mean -> probs), 2 Poisson (mean -> rate),
 *         3 Exponential (mean -> rate), 4 HalfNormal (mean2 -> scale), 5 Gamma (mean_log, mean -> concentration, rate),
 *         6 Beta (mean_log, mean_log1m -> concentration1, concentration0).  Unused second buffers may be NULL.
 * replaces: BoundPlate._update_qem_moving_avg + _update_qem_convparams (src/alan/BoundPlate.py:256-296) and
 * conversions.*Conversion.mean2conv (src/alan/conversions.py:46-296).  dtype: 0 = f32, 1 = f64. */
int alan_b200_qem_update(int family, int64_t n, double lr, const void* new0, const void* new1, void* mean0, void* mean1,
                         void* param0, void* param1, int dtype, void* stream);

/* Measurement aid (bench.py roofline denominators, measured in the same process as the bench): sustained rate of one
 * SM pipe over the whole GPU, in lane-operations per second.  which = 0: MUFU.EX2 (the unit that bounds the
 * log-semiring contraction once its d-contraction runs on the tensor cores), 1: FFMA.
 * `scratch` is device memory of at least 2 * 1024 * 4 bytes per SM.  Never on the timed path. */
int alan_b200_pipe_peak(int which, void* scratch, size_t scratch_bytes, double* ops_per_s, void* stream);

/* ---- unit-level ops, exported for the parity tests ---------- */

/* out[o] = log(sum_r exp(x[o, r] - max_r) + eps) + max_r ; x is [n_out, n_red] row-major.
 * replaces: logsumexp_dims (src/alan/utils.py:207-222).  dtype: 0 = f32, 1 = f64. */
int alan_b200_lse_eps(const void* x, void* out, int64_t n_out, int64_t n_red, int dtype, void* stream);

/* Timeseries chain: ms is [outer, T, K, K] row-major (Kprev, Kcurr); out is [outer, K]:
 * logsumexp_{Kcurr}( chain_logmmexp(ms) ), pairwise tree with the odd tail carried.
 * `levels` is scratch of alan_b200_chain_scratch_elems(outer, T, K) elements.
 * replaces: chain_logmmexp + t.logsumexp (src/alan/utils.py:478-510, logpq.py:134-143). */
int64_t alan_b200_chain_scratch_elems(int64_t outer, int64_t T, int64_t K);
int alan_b200_logmmexp_chain(const void* ms, void* levels, void* out, int64_t outer, int64_t T,
                             int64_t K, int dtype, void* stream);

/* out[c] = sum_e Normal(loc[.], scale[.]).log_prob(value[.]) over a broadcast
 * [n_cells, n_event] iteration space; each operand has (cell stride, event stride)
 * in elements (0 = broadcast).
 * replaces: TorchDimDist.log_prob for Normal (src/alan/TorchDimDist.py:127-162). */
int alan_b200_normal_logpdf_bcast(const void* value, const void* loc, const void* scale, void* out,
                                  int64_t n_cells, int64_t n_event,
                                  const int64_t* value_strides, const int64_t* loc_strides,
                                  const int64_t* scale_strides, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALAN_B200_H */
