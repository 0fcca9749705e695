/*
 * spq_b200.h -- C ABI of the B200 (sm_100a) switchable-precision fake-quant linear path.
 *
 * One shared library (libspq_b200.so), plain pointers and sizes, no torch types.  Every pointer
 * is a DEVICE pointer unless its name ends in _host; tensors are borrowed (caller-owned,
 * row-major, contiguous unless a leading dimension is given), nothing is allocated inside.
 * Every entry point enqueues on `stream` (a cudaStream_t passed as void*) and returns
 * SPQ_OK (0) or an error code; spq_last_error() gives the message.  There is no CPU fallback:
 * on a machine without an sm_100 device the compute entry points return SPQ_ERR_CUDA.
 *
 * Process model: ONE PROCESS PER GPU (the data-parallel layout of this path).  The SM count, the
 * co-resident cluster count of the CTA-pair GEMM and the per-kernel shared-memory attributes are
 * cached process-wide for the device that is current at the first call; driving several devices
 * from one process through this library is not supported.
 *
 * The reference (Laurence-Wu/LLM-QAT-on-gpt2) has no FFI; its boundary for this path is the
 * Python class API of part1_switchable_precision.  Each entry point below names the reference
 * code it replaces (file:line relative to the upstream repo root, p1 =
 * part1_switchable_precision).  The Python classes of the same names in
 * llm_qat_on_gpt2_b200/ bind these symbols with ctypes.
 */
#ifndef SPQ_B200_H
#define SPQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPQ_ABI_VERSION 6
#define SPQ_API __attribute__((visibility("default")))

typedef void* spq_stream_t;          /* cudaStream_t */
typedef uint16_t spq_half_t;         /* IEEE binary16 bit pattern */

enum { SPQ_OK = 0, SPQ_ERR_INVALID = 1, SPQ_ERR_CUDA = 2, SPQ_ERR_UNSUPPORTED = 3 };
/* how a per-channel parameter (scale, zero-point, statistic) broadcasts over a [rows, cols] view */
enum { SPQ_PER_TENSOR = 0, SPQ_PER_ROW = 1, SPQ_PER_COL = 2 };
enum { SPQ_MINMAX = 0, SPQ_LOG = 1 };
/* what the GEMM-operand output of a quantise kernel holds */
enum { SPQ_OPERAND_CODE = 0,     /* (q - zero_point): the integer code, exact in fp16 for <= 11 bits */
       SPQ_OPERAND_DEQUANT = 1,  /* the dequantised value */
       SPQ_OPERAND_RAW = 2,      /* the unquantised input (32-bit path) */
       SPQ_OPERAND_CODE_E4M3 = 3 };  /* the integer code as ONE e4m3 byte (exact for |code| <= 16: <= 4-bit min-max);
                                        the operand pointer is a byte buffer, leading dimensions in bytes */

SPQ_API int spq_abi_version(void);
SPQ_API const char* spq_last_error(void);
/* sm count and compute capability of the current device */
SPQ_API int spq_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* device-side watchdog (synchronises).  A GEMM pipeline wait that times out (seconds) TRAPS the kernel: the launch
 * fails and every later CUDA call of the process -- this one included -- returns the error, so a stalled pipeline
 * can neither hang the GPU nor produce a number.  *aborted_host reports the residual flag (2 = the dynamic
 * shared-memory base was not 1024-byte aligned). */
SPQ_API int spq_debug_status(int* aborted_host);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
SPQ_API int64_t spq_launch_count(void);

/* ---- (a) calibration statistics -------------------------------------------------------------
 * Replaces LearnableFakeQuantize._collect_statistics_batch / _reduce_min_max
 * (p1/quantization.py:152-162, 174-209).  x is viewed as [rows, cols]; the statistic is per
 * column (input activations, LoRA A/B: channel_dim = -1 / 1), per row (weights: channel_dim = 0)
 * or per tensor.  log_mode: statistics of log2(max(|x|, eps)), updated only if any(|x| > eps)
 * (p1/quantization.py:177-197); when nothing exceeds eps and accumulate == 0 the statistics are
 * filled with log2(eps).  accumulate: 0 = first batch (overwrite), 1 = running minimum/maximum.
 * state[0] |= 1 when this batch had data above eps (always for min-max mode).
 * NaN propagates as in torch.min/max.  Results are bit-exact (min/max are order independent;
 * log2 is the correctly rounded float32 logarithm, taken after the reduction -- it is monotone).
 * x_is_half: x holds float16 (the fp16 attention output feeding c_proj); values are widened exactly,
 * so the statistics equal those of x.float().  Per-row statistics take float32 only.
 */
SPQ_API size_t spq_stats_workspace_bytes(int64_t rows, int64_t cols, int bcast);
SPQ_API int spq_minmax_stats(const void* x, int x_is_half, int64_t rows, int64_t cols, int bcast, int log_mode, float eps,
                     float* stat_min, float* stat_max, int accumulate, int32_t* state,
                     void* workspace, size_t workspace_bytes, spq_stream_t stream);

/* Single-batch calibration of many small tensors in ONE launch: start_calibration + one collected batch +
 * finish_calibration for each job (the LoRA A/B quantisers of every linear are recalibrated on their own weights
 * each training step, p1/train_sp.py:125-163).  Same statistics, log transform, scale and zero-point as
 * spq_minmax_stats + spq_finish_calibration, bit for bit.  `jobs_dev` is an array of SpqCalibJob in device
 * memory; max_blocks >= the largest per-job block count (per column: ceil(cols/32); per row: ceil(rows/8);
 * per tensor: 1).  flags_dev[job] = 1 unless a log-mode job saw no |x| > eps (the caller then redoes that job
 * through the single-tensor path, which implements the reference's default-shape behaviour). */
typedef struct SpqCalibJob {
    const float* x;          /* [rows, cols] float32, dense */
    int64_t rows, cols;
    float* rmin; float* rmax; float* scale; float* zp;   /* outputs, one per channel */
    int32_t bcast;           /* SPQ_PER_COL / SPQ_PER_ROW / SPQ_PER_TENSOR */
    int32_t qtype;           /* SPQ_MINMAX / SPQ_LOG */
    int32_t symmetric;
    float levels;            /* quant_max - quant_min of the min-max range (2^(b-1)-1 symmetric, 2^b-1 otherwise) */
    float eps;
    int32_t pad_;
} SpqCalibJob;
SPQ_API int spq_calibrate_many(const SpqCalibJob* jobs_dev, int32_t n_jobs, int32_t max_blocks, int32_t* flags_dev,
                       spq_stream_t stream);

/* Replaces LearnableFakeQuantize.finish_calibration (p1/quantization.py:104-139): scale and
 * zero-point (min-max: symmetric or asymmetric; log: zero_point = log_min, scale = log_range)
 * from n running statistics.  True IEEE division, as torch-CPU. */
SPQ_API int spq_finish_calibration(const float* running_min, const float* running_max, int64_t n, int qtype,
                           int symmetric, int bits, float eps, float* scale, float* zero_point,
                           spq_stream_t stream);

/* ---- (b) quantise ---------------------------------------------------------------------------
 * Replaces MinMaxQuantizationFunction.forward / LogQuantizationFunction.forward
 * (p1/quantization_methods.py:8-22, 33-79).  x is [rows, cols]; scale / zero_point broadcast per
 * `bcast` (for log: zero_point = log_min, scale = log_range).  Outputs (each nullable):
 *   dequant : the fake-quantised float32 tensor the reference returns
 *   codes   : int32 -- min-max: the clamped rounded code q; log: the level index L
 *   sign    : int8  -- log only: sign(x) in {-1,0,1}, 0 also where |x| < 1e-5 (zero mask)
 *   operand : fp16 GEMM operand = fp16(v * row_mul[row] * col_mul[col] * mul), v chosen by
 *             operand_kind (code - zero_point, dequantised value, or raw x); row_mul/col_mul nullable.
 *             operand_transposed != 0 writes it as [cols, rows]; operand_ld > 0 is its leading
 *             dimension in elements (0 = dense), so odd widths can be padded to TMA's 16-byte strides.
 * Codes are bit-exact w.r.t. the reference arithmetic (IEEE div, round-half-even, separate
 * mul/add roundings; log2 correctly rounded via an exact slow path next to rounding ties).
 */
SPQ_API int spq_fake_quantize(const float* x, int64_t rows, int64_t cols,
                      const float* scale, const float* zero_point, int bcast,
                      int qtype, int bits, int symmetric,
                      float* dequant, int32_t* codes, int8_t* sign,
                      spq_half_t* operand, int operand_kind, const float* row_mul, const float* col_mul,
                      float mul, int operand_transposed, int64_t operand_ld, spq_stream_t stream);

/* Fused activation-side quantise for SPLinearWithLoRA.forward (p1/lora.py:141,149): one elementwise
 * pass over x [M, K] (per-column or per-tensor calibrated scale) produces
 *   a_q   [M, K] fp16 : the quantised GEMM operand (code or dequant) * col_mul[k] * mul, and
 *   a_raw [M, K] fp16 : the UNquantised x * raw_col_mul[k] (the LoRA branch reads x, not q(x));
 *                       raw_col_mul is a power of two per input channel chosen from the calibrated
 *                       bound (spq_prep_linear_scales), the conversion saturates.  a_raw is optional.
 * x is float32, or float16 when x_is_half (widened exactly: same results as on x.float()). */
SPQ_API int spq_quantize_act(const void* x, int x_is_half, int64_t M, int64_t K,
                     const float* scale, const float* zero_point, int bcast,
                     int qtype, int bits, int symmetric, int operand_kind, const float* col_mul, float mul,
                     spq_half_t* a_q, spq_half_t* a_raw, const float* raw_col_mul, spq_stream_t stream);

/* Scale preparation for the fused linear (host-side glue of p1/lora.py:141-150 made one launch):
 * from the input quantiser's calibrated (scale, zero_point) [in_n = 1 or K], the per-row absmax of
 * the dequantised weight [N] and, optionally, |q(A)| [K, r] of the active LoRA adapter, computes the
 * per-K factor the weight operand absorbs, the activation operand multiplier, the power-of-two
 * multiplier of the raw fp16 operand per input channel (and its inverse), the power-of-two row
 * normaliser pw[n] (and 1/pw) and lora_vec (8 r floats) = tau | 1/tau | lora_scaling/tau | pa | 1/pa |
 * pb | 1/pb | pa*tau  (tau: static pre-scale of t = x q(A) from the calibrated input bound; pa[j]: normaliser of
 * column j of q(A)[k,j] / raw_mul[k], the operand of the LoRA down-projection; pb[j], written when
 * bq = q(B) [r, N] is given: normaliser of row j of lora_scaling * q(B), the operand of dT = dY q(B)^T). */
SPQ_API int spq_prep_linear_scales(const float* in_scale, const float* in_zero_point, int64_t in_n,
                           int qtype, int bits, int symmetric, int64_t K,
                           const float* w_rowmax, int64_t N, const float* aq, const float* bq, int64_t r,
                           float lora_scaling, float* absorb, float* act_mul, float* raw_mul,
                           float* inv_raw_mul, float* pw, float* inv_pw, float* lora_vec, spq_stream_t stream);

/* CPT variant (p2/cpt_model.py:92-114): the scale vectors of one CPTLinear's shared-LoRA level in one launch.
 * aq [K, r] = q(A), bq [N, r] = q(B) (dequantised, row-major), absorb [K] = per-K factor of the quantised activation operand,
 * xbound [K] = calibrated bound of |q(x)|; out [8 r] = pa | 1/pa | tau | 1/tau | pa tau | scaling/tau | pb | 1/pb (all powers
 * of two except scaling/tau): see csrc/spq_prep.cu.  r must divide 1024. */
SPQ_API int spq_cpt_lora_scales(const float* aq, const float* bq, const float* absorb, const float* xbound, int64_t K, int64_t N,
                        int64_t r, float scaling, float* out, spq_stream_t stream);

/* STE backward (p1/quantization_methods.py:25-28, 82-90): identity (min-max) or clamp to
 * [-10, 10] (log).  out may alias grad. */
SPQ_API int spq_ste_backward(const float* grad, int64_t n, int qtype, float* out, spq_stream_t stream);

/* ---- (c) fused quant-GEMM -------------------------------------------------------------------
 * Replaces F.linear(q(x), q(W), b) + LoRALayer.forward (p1/lora.py:45-54, 144-150):
 *   D[m,n] = epi( sum_k A[m,k] B[n,k]  +  sum_j A2[m,j] B2[n,j] )
 *   epi(v) = act( clamp(v * alpha * row_scale[m] * col_scale[n], +-clamp_abs) + bias[n] + C[m,n] )
 * A, B, A2, B2 are fp16, K-contiguous ("K-major"), leading dimensions in elements (multiples of
 * 8); the second segment (the LoRA up-projection folded in as extra K) is optional (K2 = 0).
 * row_scale, col_scale, bias, C are nullable; clamp_abs <= 0 disables the clamp; activation 0 = none,
 * 1 = exact erf GELU (nn.GELU(), p1/models_sp.py:114, fused here for the no-grad MLP path).  D is float32,
 * or fp16 (saturating) when d_is_half.  When N is not a multiple of 4 and ldd >= 4*ceil(N/4), the 1-3
 * floats of row padding after column N may be overwritten (16-byte store granularity).  TMA-fed tcgen05.mma (kind::f16, fp32 accumulation in
 * TMEM), persistent over the SMs.
 */
SPQ_API int spq_qgemm(const spq_half_t* A, int64_t lda, const spq_half_t* B, int64_t ldb,
              int64_t M, int64_t N, int64_t K,
              const spq_half_t* A2, int64_t lda2, const spq_half_t* B2, int64_t ldb2, int64_t K2,
              float alpha, const float* row_scale, const float* col_scale, const float* bias,
              float clamp_abs, const float* C, int64_t ldc,
              void* D, int64_t ldd, int d_is_half, int activation, spq_stream_t stream);

/* The same fused GEMM on e4m3 integer-code operands (north_star (b)/(c): "codes held in e4m3 where they fit",
 * tcgen05.mma.kind::f8f6f4 at twice the fp16 rate): A [M, K] and B [N, K] are e4m3 BYTES (SPQ_OPERAND_CODE_E4M3;
 * lda / ldb in bytes, multiples of 16), exact for the <= 4-bit min-max codes of the per-tensor-scale evaluation
 * configurations (p1/deploy.py:210,238; part3_eval_sp/main_sp_eval.py:60): with K * 7^2 < 2^24 the fp32 accumulator holds
 * the integer dot product exactly, and the scales s_x * s_w[n] arrive through col_scale.  The optional second K
 * segment (A2, B2: the fp16 LoRA operands) is accumulated with kind::f16 into the same TMEM tile.  No residual. */
SPQ_API int spq_qgemm_f8(const uint8_t* A, int64_t lda, const uint8_t* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 const spq_half_t* A2, int64_t lda2, const spq_half_t* B2, int64_t ldb2, int64_t K2, float alpha,
                 const float* row_scale, const float* col_scale, const float* bias, void* D, int64_t ldd,
                 int d_is_half, int activation, spq_stream_t stream);

/* Transposed-operand GEMM for the weight-gradient shaped products of the STE backward
 * (dA = x^T dT, dB = t^T dY, optional dW = dY^T q(x); torch autograd of p1/lora.py:51-52, 144):
 *   D[i,j] = clamp( alpha * alpha_dev[0] * i_scale[i] * j_scale[j] * sum_m P[m,i] Q[m,j], +-clamp_abs )
 * P: [Mred, I] fp16, Q: [Mred, J] fp16, row-major (leading dimensions multiples of 8): both are
 * MN-major tcgen05 operands loaded by TMA without a transpose.  D is a dense float32 matrix
 * addressed D[i * d_stride_i + j * d_stride_j], either [I, J] row-major or its transpose [J, I].
 * The reduction over Mred is split across CTAs; each split writes its partial tile to its own plane of
 * `workspace` (spq_gemm_tn_workspace_bytes) and a second kernel folds the planes in a fixed order, so the
 * result is bitwise reproducible (no atomics).  clamp_abs > 0 applies the log quantiser's STE clamp
 * (p1/quantization_methods.py:82-90) to the finished gradient.  gq_scale_i (nullable, [I]) fuses the CPT variant's
 * GradientQuantizer (p2/quantization.py:14-26) in front of that clamp: symmetric gq_bits-bit min-max fake
 * quantisation of the gradient with one calibrated scale per row i.  alpha_dev (device scalar), i_scale, j_scale
 * are nullable.  accumulate != 0: D += result (the gradient lands in an existing .grad buffer, e.g. a slice of the flat
 * gradient buffer of the training step).  Per-reduction-row scales cannot be applied here: fold them into P or Q. */
SPQ_API size_t spq_gemm_tn_workspace_bytes(int64_t Mred, int64_t I, int64_t J);
SPQ_API int spq_gemm_tn(const spq_half_t* P, int64_t ldp, const spq_half_t* Q, int64_t ldq,
                int64_t Mred, int64_t I, int64_t J, float alpha, const float* alpha_dev,
                const float* i_scale, const float* j_scale, float clamp_abs,
                const float* gq_scale_i, int gq_bits, int accumulate,
                float* D, int64_t d_stride_i, int64_t d_stride_j, void* workspace, size_t workspace_bytes,
                spq_stream_t stream);

/* ---- SwitchableLayerNorm (p1/switchable_batchnorm.py:102-109) ------------------------------ */
SPQ_API int spq_layernorm_fwd(const float* x, int64_t rows, int64_t cols, const float* weight, const float* bias,
                      float eps, float* y, float* mean, float* rstd, spq_stream_t stream);
SPQ_API size_t spq_layernorm_bwd_workspace_bytes(int64_t rows, int64_t cols);
/* accumulate_params != 0: dweight / dbias += the column sums (gradient accumulation in place) */
SPQ_API int spq_layernorm_bwd(const float* dy, const float* x, const float* weight, const float* mean,
                      const float* rstd, int64_t rows, int64_t cols, float* dx, float* dweight,
                      float* dbias, int accumulate_params, void* workspace, size_t workspace_bytes, spq_stream_t stream);

/* the column-sum fold of spq_layernorm_bwd as a call of its own: after spq_layernorm_bwd(..., dweight = dbias = NULL, ...)
 * left the per-CTA sums in `workspace` (private to that call), this writes / accumulates dweight and dbias */
SPQ_API int spq_layernorm_bwd_finalize(const void* workspace, size_t workspace_bytes, int64_t rows, int64_t cols, float* dweight,
                               float* dbias, int accumulate_params, spq_stream_t stream);

/* ---- SwitchableLayerNorm fused into its consumer's activation-side kernel (no float32 round trip of the normalised
 * rows): ln_1 -> c_attn, ln_2 -> c_fc, ln_f -> LM head (p1/models_sp.py:139-147, 316-319; p1/lora.py:141-149).
 * Normalised dim: a multiple of 4, <= 2048.  y_out (optional) receives the float32 normalised rows as well.
 *
 * spq_ln_quantize_act: a_q / a_raw exactly as spq_quantize_act would write them for layernorm(x) (same arguments). */
SPQ_API int spq_ln_quantize_act(const float* x, int64_t M, int64_t K, const float* ln_weight, const float* ln_bias, float ln_eps,
                        const float* scale, const float* zero_point, int bcast, int qtype, int bits, int symmetric,
                        int operand_kind, const float* col_mul, float mul, spq_half_t* a_q, spq_half_t* a_raw,
                        const float* raw_col_mul, float* y_out, spq_stream_t stream);
/* spq_ln_rowscale_stats: out / row_scale as spq_rowscale_f16 on layernorm(x) (dense rows); stats_mode 1 / 2 also folds
 * the per-column min / max of layernorm(x) (2: of |.|, log2 applied after the fold, `state` |= any(|.| > stat_eps)) into
 * stat_min / stat_max [K] exactly as spq_minmax_stats (bcast = SPQ_PER_COL, log_mode = stats_mode == 2) would. */
SPQ_API size_t spq_ln_rowscale_stats_workspace_bytes(int64_t M, int64_t K);
SPQ_API int spq_ln_rowscale_stats(const float* x, int64_t M, int64_t K, const float* ln_weight, const float* ln_bias, float ln_eps,
                          spq_half_t* out, float* row_scale, int stats_mode, float stat_eps, float* stat_min,
                          float* stat_max, int accumulate, int32_t* state, float* y_out, void* workspace,
                          size_t workspace_bytes, spq_stream_t stream);

/* ---- gradient-side operand: out[m, 0:N] = fp16(g[m,n] * 2^-e[m]), e from the row's absmax,
 * row_scale[m] = 2^e[m] (an all-zero row reports the smallest scale, 2^-108, so that `max over rows` and
 * `scale / max` ignore it; inf / NaN rows report 1); rows of `out` are ld_out elements apart (0 = N).  Any N (a two-pass
 * kernel takes over when the row does not fit the register-resident one or is unaligned).
 * g is float32, or float16 when g_is_half (register-resident shapes only). */
SPQ_API int spq_rowscale_f16(const void* g, int g_is_half, int64_t M, int64_t N, spq_half_t* out, int64_t ld_out,
                     float* row_scale, spq_stream_t stream);
/* same, and max_scale[0] = max over rows of row_scale (what the token-reduction GEMMs fold the scales against) */
SPQ_API int spq_rowscale_f16_max(const void* g, int g_is_half, int64_t M, int64_t N, spq_half_t* out, int64_t ld_out,
                         float* row_scale, float* max_scale, spq_stream_t stream);

/* spq_rowscale_f16_max of g * gelu'(y): the gradient entering c_fc when its output went through nn.GELU() (p1/models_sp.py:
 * 114-123 under autograd), taken from the gradient g of the GELU output and the pre-activation y -- replaces torch's
 * gelu_backward pass and the float32 gradient it writes.  Dense float32 rows, N % 4 == 0, N <= 8192. */
SPQ_API int spq_rowscale_dgelu_f16_max(const float* g, const float* y, int64_t M, int64_t N, spq_half_t* out, float* row_scale,
                               float* max_scale, spq_stream_t stream);

/* fp16 operands of the LoRA gradient GEMMs in one pass (STE backward of p1/lora.py:45-54):
 *   dt16 = fp16(dtn * dt_mul)                       (dX += dT q(A)^T; token scale applied in that GEMM's epilogue)
 *   dt2  = fp16(dtn * dt_mul * row_scale / max)     (dA = x^T dT: reduction over tokens, scale folded in)
 *   t2   = fp16(t16 * row_scale / max)              (dB = t^T dY)
 * dtn [M, r] float32 is the down-projected gradient per token scale, t16 [M, r] (leading dimension ld_t16) the fp16
 * down-projection saved by the forward; row_scale / max_scale come from spq_rowscale_f16_max.  Any output may be NULL. */
SPQ_API int spq_lora_bwd_prep(const float* dtn, const spq_half_t* t16, int64_t ld_t16, const float* row_scale,
                      const float* max_scale, int64_t M, int64_t r, float dt_mul, spq_half_t* dt16, spq_half_t* dt2,
                      spq_half_t* t2, spq_stream_t stream);

/* LM head with the log-sum-exp folded into the GEMM epilogue ("fused CE over V = 50257
 * with the tied LM-head GEMM"; replaces lm_head + the softmax passes of nn.CrossEntropyLoss,
 * p1/models_sp.py:436-449).  Same product and epilogue as spq_qgemm with float32 D (rows 16-byte aligned and
 * padded: ldd >= 4*ceil(N/4)); additionally lse_part [M, lse_ld] receives, per row and per column half-tile, the
 * pair (max, sum exp(v - max)) of the values stored.  spq_qgemm_lse_parts(M, N) is the number of pairs per row;
 * spq_cross_entropy_from_parts folds them and reads only the target logit of each row. */
SPQ_API int64_t spq_qgemm_lse_parts(int64_t M, int64_t N);
SPQ_API int spq_qgemm_lse(const spq_half_t* A, int64_t lda, const spq_half_t* B, int64_t ldb, int64_t M, int64_t N,
                  int64_t K, float alpha, const float* row_scale, const float* col_scale, const float* bias,
                  float* D, int64_t ldd, float* lse_part, int64_t lse_ld, spq_stream_t stream);
SPQ_API int spq_cross_entropy_from_parts(const float* parts, int64_t P, int64_t part_ld, const float* logits, int64_t M,
                                 int64_t V, int64_t ld, const int64_t* targets, int64_t ignore_index,
                                 float* row_loss, float* row_valid, spq_stream_t stream);

/* Distillation loss of the SP training step (p1/distillation_manager.py:64-80):
 * row_loss[m] = KL( softmax(t[m]/T) || softmax(s[m]/T) ) and, when grad != NULL, the dense [M, V] gradient
 * grad[m,v] = grad_scale / T * (softmax(s[m]/T)[v] - softmax(t[m]/T)[v])  (grad_scale = T^2 / rows for the
 * reference's batchmean * T^2).  Rows with m % seq_len == seq_len - 1 are ignored (seq_len = 0: none) -- the
 * reference scores positions 0..T-2.  Logits may have padded rows (ld_s, ld_t in elements). */
/* Feature term of the distillation loss (p1/distillation_manager.py:82-116: F.mse_loss(student_hidden[l], teacher_hidden[l])
 * for ONE layer l drawn per micro-step): out[0] = mean((a[l] - b[l])^2) with l = select[0] read on the DEVICE (clamped to
 * [0, n_pairs)), so that a captured CUDA graph serves every draw.  a / b: host arrays of n_pairs (<= 32) device pointers to
 * float32 tensors of numel elements, 16-byte aligned.  Deterministic (fixed-order fold). */
SPQ_API size_t spq_mse_select_workspace_bytes(void);
SPQ_API int spq_mse_select(const float* const* a, const float* const* b, int n_pairs, const int32_t* select, int64_t numel,
                   float* out, void* workspace, size_t workspace_bytes, spq_stream_t stream);

SPQ_API int spq_distill_kl(const float* s_logits, int64_t ld_s, const float* t_logits, int64_t ld_t, int64_t M, int64_t V,
                   float temperature, int64_t seq_len, float grad_scale, float* row_loss, float* grad,
                   spq_stream_t stream);

/* Softmax loss over the LM-head logits whose gradient is emitted as the fp16 operand of the LM-head backward GEMM
 * (the float32 [B, T, V] dlogits matrix is never materialised).
 *   kind 0: distillation KL of p1/distillation_manager.py:64-80 (as spq_distill_kl);  d = softmax(s/T) - softmax(t/T)
 *   kind 1: next-token cross-entropy of p1/models_sp.py:441-449 (targets already shifted; rows whose target is
 *           ignore_index or out of range are not scored);                              d = softmax(s) - onehot(target)
 * row_loss[m] as spq_distill_kl / spq_cross_entropy_fwd, row_valid[m] (nullable) 1 for scored rows;
 * g16[m, 0:V] = fp16(d * 2^(8-E[m])), row_scale[m] = 2^(E[m]-8), E[m] from an analytic upper bound of max_v |d| (the row maxima
 * of the softmaxes: |g16| <= 256 always, >= 64 for cross-entropy) (unscored rows: zeros, scale 2^-108),
 * max_scale[0] = max_m row_scale[m]; g16 rows are ld_g (even, >= V) elements apart, the padding is zeroed.
 * The caller's scalar factor (T/rows for kind 0, 1/scored-rows for kind 1) multiplies row_scale. */
SPQ_API int spq_softmax_loss_grad16(int kind, const float* s_logits, int64_t ld_s, const float* t_logits, int64_t ld_t,
                            const int64_t* targets, int64_t ignore_index, int64_t M, int64_t V, float temperature,
                            int64_t seq_len, float* row_loss, float* row_valid, spq_half_t* g16, int64_t ld_g,
                            float* row_scale, float* max_scale, spq_stream_t stream);

/* ---- consumer of the path: next-token cross-entropy, forward only -----------
 * Replaces nn.CrossEntropyLoss over re-materialised shifted logits (p1/models_sp.py:441-449) for
 * no-grad evaluation: row_loss[m] = logsumexp(logits[m, 0:V]) - logits[m, targets[m]], row_valid[m] = 1,
 * or both 0 where targets[m] == ignore_index.  Rows of `logits` are `ld` floats apart (ld >= V). */
SPQ_API int spq_cross_entropy_fwd(const float* logits, int64_t M, int64_t V, int64_t ld, const int64_t* targets,
                          int64_t ignore_index, float* row_loss, float* row_valid, spq_stream_t stream);

/* ---- optimiser step of the SP training step on flat buffers -----------------------------
 * Replaces `scaler.unscale_; clip_grad_norm_(model.parameters(), 1.0); AdamW.step` of p1/train_sp.py:390-393 for
 * the path's trainable parameters (LoRA A/B, LayerNorm pairs) when they live in one flat fp32 buffer (their .grad
 * tensors are slices of a second one, which is also what the data-parallel all-reduce sends).
 * spq_grad_sumsq: out_sumsq[0] = scale^2 * sum g[i]^2, fixed-order tree (bitwise reproducible).
 * spq_adamw_flat: one pass over a segment; g' = g * grad_scale * min(1, max_norm / (sqrt(total_sumsq[0]) + 1e-6))
 * (max_norm <= 0 or total_sumsq NULL: no clipping), then torch.optim.AdamW's update with decoupled weight decay
 * and bias corrections for `step` (1-based count of updates this segment has received). */
SPQ_API size_t spq_sumsq_workspace_bytes(void);
SPQ_API int spq_grad_sumsq(const float* g, int64_t n, float scale, float* out_sumsq, void* workspace, size_t workspace_bytes,
                   spq_stream_t stream);
SPQ_API int spq_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                   const float* total_sumsq, float max_norm, spq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPQ_B200_H */
