/*
 * dpivae_b200.h -- C ABI of the B200-native DPI-VAE training-step library (libdpivae_b200.so).
 *
 * The reference (JanKoune/DPI-VAE) has NO native/FFI interface: its boundary for this path is the
 * Python surface `models/vae.py` (DPIVAE.loss/forward/encode/sample) + `dpivae.py`
 * (setup_model/train_model/evaluate_model).  Each entry point below states which reference
 * lines it replaces; the Python mirror in dpivae_b200/ binds them with ctypes.
 *
 * Conventions: plain C, no torch types.  All tensor pointers are DEVICE pointers (fp32, row-major,
 * contiguous) unless the name ends in `_host`.  Work is enqueued on the caller's `stream`
 * (a cudaStream_t passed as void*).  No entry point allocates device memory per call: scratch is
 * a caller-provided workspace sized by dpivae_workspace_bytes().  Return value 0 = ok; non-zero =
 * error, message in dpivae_last_error().  No exception crosses the ABI.
 */
#ifndef DPIVAE_B200_H
#define DPIVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPIVAE_MAX_ZX 4      /* physics latents            */
#define DPIVAE_MAX_ZCY 8     /* nz_c, nz_y each            */
#define DPIVAE_MAX_Z 16      /* nz_x + nz_c + nz_y         */
#define DPIVAE_MAX_NDX 64    /* response vector length     */
#define DPIVAE_MAX_NDCY 4    /* nd_c, nd_y each            */
#define DPIVAE_MAX_PHYS_LAYERS 6

/* One hidden-layer ReLU MLP: in -> hid -> out.  Weights live in the flat parameter buffer in
 * torch nn.Linear layout: w0 (hid,in), b0 (hid), w1 (out,hid), b1 (out); offsets in floats.
 * For encoders / prior nets `w1` is the concatenation [f_mean ; f_sigma ; f_cov] (models/encoders.py:20-22). */
typedef struct {
  int32_t in_dim, hid, out_dim, _pad;
  int64_t w0, b0, w1, b1;
} dpivae_mlp2_t;

enum { DPIVAE_MODEL_P = 0, DPIVAE_MODEL_S = 1 };
enum { DPIVAE_PHYS_MLP = 0, DPIVAE_PHYS_MASS_SPRING = 1, DPIVAE_PHYS_BEAM = 2 };
enum { DPIVAE_PRIOR_UNIFORM = 0, DPIVAE_PRIOR_NORMAL = 1 };

/* Everything `setup_model` (dpivae.py:89-283) fixes for a model instance. */
typedef struct {
  int32_t model_type;                 /* DPIVAE_MODEL_P: encoder, encoder_c, encoder_y ; _S: one encoder */
  int32_t nz_x, nz_c, nz_y, nd_x, nd_c, nd_y, nd_p;
  int32_t idx_c_phys[DPIVAE_MAX_NDCY]; /* columns of raw c appended to zx (models/vae.py:171) */
  dpivae_mlp2_t enc[3];               /* P: x,c,y (nz_k latents each) ; S: enc[0] over Z = nz_x+nz_c+nz_y */
  dpivae_mlp2_t prior[2];             /* FactorizedNN prior nets p(zc|c), p(zy|y) (models/encoders.py:96-128) */
  dpivae_mlp2_t fx, dec_c, dec_y;     /* GradRevAdditive.fx0/fx1, Decoder c / y (models/decoders.py) */
  int64_t log_sigma_x;                /* offset of the scalar parameter (models/vae.py:70) */
  int64_t n_params;                   /* length of the flat trainable buffer */
  float mean_x[DPIVAE_MAX_NDX], std_x[DPIVAE_MAX_NDX];   /* StandardScaler stats (utils/transforms.py:64-68) */
  float mean_c[DPIVAE_MAX_NDCY], std_c[DPIVAE_MAX_NDCY];
  float mean_y[DPIVAE_MAX_NDCY], std_y[DPIVAE_MAX_NDCY];
  float lb[DPIVAE_MAX_ZX], ub[DPIVAE_MAX_ZX];            /* Logistic -> ShiftScale bounds (dpivae.py:184-187) */
  int32_t prior_kind[DPIVAE_MAX_ZX];                     /* prior over zx (utils/priors.py:19-23) */
  float prior_a[DPIVAE_MAX_ZX], prior_b[DPIVAE_MAX_ZX];  /* uniform: low, high ; normal: loc, scale */
  float lambda_g0;                    /* constant GRL scale (utils/transforms.py:207-219,235) */
  int32_t has_lambda_x;               /* optional regulariser on xh_d (models/vae.py:217-219) */
  float lambda_x;
  int32_t phys_kind;
  int32_t phys_n_layers;              /* MLP surrogate: number of Linear layers (Tanh between) */
  int32_t phys_dims[DPIVAE_MAX_PHYS_LAYERS + 1];
  float phys_grid[DPIVAE_MAX_NDX];    /* mass_spring: t ; beam: x = linspace(0, 1, nd_x) */
} dpivae_model_desc_t;

typedef struct dpivae_model* dpivae_handle_t;

/* One minibatch (dpivae.py:403-404).  x/c/y may be the whole resident dataset with `idx` selecting
 * rows (the reference's multinomial gather, fused), or already-gathered rows with idx == NULL. */
typedef struct {
  const float* x;        /* (rows, nd_x) */
  const float* c;        /* (rows, nd_c) */
  const float* y;        /* (rows, nd_y) ; may be NULL for forward/sample/encode */
  const int64_t* idx;    /* (B) row indices into x/c/y, or NULL */
  int64_t B;             /* rows this call processes (this rank's shard) */
  int64_t B_global;      /* rows of the global minibatch: loss normaliser + noise indexing */
  int64_t row_offset;    /* global index of this shard's first row */
  int32_t n_mc;          /* Monte-Carlo samples per row */
  int32_t cond;          /* forward(cond=True): zc drawn from the prior net (models/vae.py:165-167) */
  int64_t row_stride;    /* global index of local row r = row_offset + r * row_stride; 0 or 1 = a contiguous block.
                          * Cyclic shards (rank k of N: row_offset = k, row_stride = N) keep the four noise elements that
                          * torch's normal_ kernel draws from ONE Philox evaluation on one rank */
} dpivae_batch_t;

/* Reparameterisation noise.  mode 0: injected buffers eps[k] of shape (n_mc, B_global, nz_k)
 * (P: k = x,c,y ; S: eps[0] of width Z ; eps[3] = the cond draw).  mode 1: in-kernel Philox4x32-10
 * reproducing torch.cuda's normal_() stream: tensor k starts at philox offset `offset[k]`, and the
 * element -> (subsequence, counter, lane) map uses `grid_threads[k]` = 256 * grid of the torch
 * launch (ATen/native/cuda/DistributionTemplates.h); dpivae_philox_plan() fills both. */
typedef struct {
  int32_t mode;
  int32_t _pad;
  const float* eps[4];
  uint64_t seed;
  uint64_t offset[4];
  uint32_t grid_threads[4];
} dpivae_rng_t;

typedef struct {
  float beta_x, alpha_x, alpha_c, alpha_y;  /* models/vae.py:177-231; beta_c/beta_y are unused there */
} dpivae_loss_weights_t;

/* Optional outputs; any pointer may be NULL. */
typedef struct {
  float* row_loss;   /* (6, B): loss, KL_x, R_x, R_c, R_y, reg per datapoint (models/vae.py:222-231) */
  float* scalars;    /* (8): ELBO/(B_global*(nd_x+nd_c+nd_y)), KL_x, 0, 0, R_x, R_c, R_y, reg (each /B_global) -- dpivae.py:419-426 */
  float* xh_p; float* xh_d;            /* (n, B, nd_x) */
  float* ch; float* log_sigma_c;       /* (n, B, nd_c) */
  float* yh; float* log_sigma_y;       /* (n, B, nd_y) */
  float* zx; float* zc; float* zy;     /* (n, B, nz_*) */
  float* dens_z;                       /* (n, B) */
} dpivae_outputs_t;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int dpivae_create(const dpivae_model_desc_t* desc, dpivae_handle_t* out);
int dpivae_destroy(dpivae_handle_t h);
const char* dpivae_last_error(void);
#define DPIVAE_ABI_VERSION 3   /* 2: dpivae_regression_metrics returns per-output raw values; dpivae_sizeof_model_desc; 3: dpivae_batch_t.row_stride */
int dpivae_abi_version(void);
size_t dpivae_sizeof_model_desc(void);   /* sizeof(dpivae_model_desc_t) in the binary: bindings check their struct layout */

/* Frozen physics surrogate (cases/bridge/__init__.py:163-174, models/nn.py:67-80): host arrays in
 * nn.Linear layout, concatenated layer by layer; input scaler mean/std of length phys_dims[0]. */
int dpivae_set_physics_mlp(dpivae_handle_t h, const float* weights_host, const float* biases_host,
                           const float* in_mean_host, const float* in_std_host);

/* Flat parameter / gradient / Adam-state buffers (n_params floats each; grads/m/v may be NULL for
 * inference-only use).  Replaces the per-tensor nn.Parameter storage of the reference. */
int dpivae_bind(dpivae_handle_t h, float* params, float* grads, float* exp_avg, float* exp_avg_sq);

/* Adam parameter groups (dpivae.py:335-363): group g covers flat range [begin[g], end[g]). */
int dpivae_set_groups(dpivae_handle_t h, int32_t n_groups, const int64_t* begin, const int64_t* end,
                      const float* lr, const float* weight_decay);

/* ---- the hot path --------------------------------------------------------------------------- */
size_t dpivae_workspace_bytes(dpivae_handle_t h, int64_t B, int32_t n_mc);

/* DPIVAE.loss (models/vae.py:177-231) [+ forward outputs].  with_grad != 0 additionally runs the
 * fused backward of ELBO/(B_global*(nd_x+nd_c+nd_y)) (dpivae.py:419,429) and leaves THIS SHARD's
 * gradient sum in the bound grads buffer and the 8 scalars (shard sums, already normalised by the
 * global batch) in out->scalars -- ready for one allreduce(sum) across shards. */
int dpivae_loss(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng,
                const dpivae_loss_weights_t* w, int32_t with_grad, const dpivae_outputs_t* out,
                void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam.step (dpivae.py:373,436) over the flat buffers, per-group lr / L2 weight decay,
 * betas (0.9, 0.999), eps 1e-8.  `step` is 1-based.  max_grad_norm > 0 applies clip_grad_norm_
 * (dpivae.py:432-433) first. */
int dpivae_adam_step(dpivae_handle_t h, int64_t step, float max_grad_norm, void* stream);

/* dpivae.py:390-436 minus the minibatch draw: loss(with_grad) + Adam in one call (single shard). */
int dpivae_train_step(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng,
                      const dpivae_loss_weights_t* w, int64_t step, float max_grad_norm,
                      const dpivae_outputs_t* out, void* workspace, size_t workspace_bytes, void* stream);

/* Device-resident training loop (dpivae.py:390-436 iterated): ONE training step -- minibatch gather from a
 * resident index pool (dpivae.py:403-404), loss, backward, clip, Adam -- captured in a CUDA graph whose per-step
 * values (optimizer step -> Adam bias corrections, generator offset, pool row) live in device memory and are
 * advanced by the graph's first kernel, so replaying it needs no host arguments and no host synchronisation.
 *   rng          mode 1; offset[] / grid_threads[] = plan of the FIRST step (dpivae_philox_plan)
 *   philox_inc   generator offset one step consumes (plan's return value - its offset_in)
 *   w            loss weights, constant over the captured steps (annealing == None, utils/annealing.py:12-14)
 *   idx_pool     [pool_rows][B] int64 rows of the resident data set, step t uses row (t-1) % pool_rows; NULL = batch->idx / identity
 *   idx_cur      [B] int64 scratch the step's kernels gather through (required with idx_pool)
 *   scalars      the 8 loss scalars of the latest step; must directly follow the bound gradient buffer
 *   step_log     optional [log_cap][9] ring: row (t-1) % log_cap = 8 loss scalars + log_sigma_x after step t
 *                (the per-iteration values the reference logs, dpivae.py:439-451)
 *   unroll       >= 1: besides the single-step graph, a second graph of `unroll` consecutive steps is captured;
 *                dpivae_step_graph_launch uses it for every full group of `unroll` steps (fewer host launches)
 * The workspace, data and index buffers must stay alive and unchanged in address while the graph exists. */
typedef struct dpivae_step_graph* dpivae_step_graph_t;
int dpivae_step_graph_create(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng, uint64_t philox_inc,
                             const dpivae_loss_weights_t* w, int64_t first_step, float max_grad_norm,
                             const int64_t* idx_pool, int64_t pool_rows, int64_t* idx_cur, float* scalars, float* step_log,
                             int64_t log_cap, int32_t unroll, void* workspace, size_t workspace_bytes, void* stream,
                             dpivae_step_graph_t* out);
/* Re-base the device state: the next replay is optimizer step `next_step` and draws at rng->offset[]. */
int dpivae_step_graph_reset(dpivae_step_graph_t g, const dpivae_rng_t* rng, int64_t next_step, void* stream);
/* Enqueue n_steps replays on `stream` (no host synchronisation). */
int dpivae_step_graph_launch(dpivae_step_graph_t g, int32_t n_steps, void* stream);
int dpivae_step_graph_destroy(dpivae_step_graph_t g);

/* transform_inputs -> encode (models/vae.py:161-162, 125-151): zx (n,B,nz_x), zc, zy, dens_z (n,B).
 * x_is_standardised != 0 mirrors DPIVAE.encode(x_t, n), which receives already-scaled inputs. */
int dpivae_encode(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng,
                  int32_t x_is_standardised, float* zx, float* zc, float* zy, float* dens_z,
                  void* workspace, size_t workspace_bytes, void* stream);

/* DPIVAE.decode (models/vae.py:153-158) on caller-supplied latents: zx_in (n,B,nz_x+nd_p) = [zx | c_phys],
 * zc (n,B,nz_c), zy (n,B,nz_y) -> out->xh_p, xh_d (n,B,nd_x), ch, log_sigma_c (n,B,nd_c), yh, log_sigma_y (n,B,nd_y)
 * (NULL outputs are skipped).  Forward only, fp32 FFMA kernels. */
int dpivae_decode(dpivae_handle_t h, const float* zx_in, const float* zc, const float* zy, int64_t B, int32_t n_mc,
                  const dpivae_outputs_t* out, void* workspace, size_t workspace_bytes, void* stream);

/* DPIVAE.prior_net (models/vae.py:99-110): raw c (B,nd_c) and optional raw y (B,nd_y) -> standardise ->
 * FactorizedNN (models/encoders.py:121-128) -> loc (B,nz), scale_tril (B,nz,nz) = diag(sigma + 1e-8). */
int dpivae_prior_net(dpivae_handle_t h, const float* c, const float* y, int64_t B, float* loc_c, float* scale_tril_c,
                     float* loc_y, float* scale_tril_y, void* workspace, size_t workspace_bytes, void* stream);

/* GaussianEncoder.sample without output transform (models/encoders.py:73-93) on given parameters:
 * z = loc + L eps (n,B,nz), dens = log N(z; loc, L L^T) (n,B); eps (n,B,nz) standard normal draws. */
int dpivae_gaussian_sample(const float* loc, const float* scale_tril, const float* eps, int32_t n_mc, int64_t B, int32_t nz,
                           float* z, float* dens, void* stream);

/* Post-processing of sample / encode outputs on the device (the callers right behind the hot path).
 * dpivae_mc_mean: out (B,d) = mean over the leading MC axis of v (n,B,d)  (dpivae.py:548 `y_sample.mean(0)`,
 *   dpivae.py:640-661 `z.mean(0)`).
 * dpivae_regression_metrics: utils/metrics.py:11-32 = sklearn r2_score, mean_squared_error, mean_absolute_error with
 *   multioutput="raw_values" (one value per output column) -> out3d = [R2 (d) | MSE (d) | MAE (d)], 3*d floats;
 *   scratch = 4*d doubles (device).
 * dpivae_linreg_r2: dpivae.py:672-690 with regressor == "linear": ordinary least squares (with intercept) of one
 *   target column on k <= 8 latent columns of the training set, R2 of the fit on the test set.  y_* point at the
 *   target column, ldy_* = its row stride in floats.  scratch = 80 doubles (device). */
int dpivae_mc_mean(const float* v, int32_t n_mc, int64_t B, int32_t d, float* out, void* stream);
int dpivae_regression_metrics(const float* y_true, const float* y_pred, int64_t N, int32_t d, double* scratch, float* out3d,
                              void* stream);
int dpivae_linreg_r2(const float* X_train, const float* y_train, int64_t ldy_train, int64_t N_train, const float* X_test,
                     const float* y_test, int64_t ldy_test, int64_t N_test, int32_t k, double* scratch, float* r2_out,
                     void* stream);

/* On-device synthetic data generator = utils/data.py:9-52 `sample_response` for a Tanh-MLP `full_model` behind a
 * StandardScaler (cases/<case>/__init__.py): z_j ~ Uniform(lo_j, hi_j) (utils/priors.py:32-36), x = full_model(z) + sigma_x N(0,1),
 * c = z[idx_c] + sigma_c N(0,1), y = z[idx_y] + sigma_y N(0,1).  Random numbers are the torch.cuda Philox stream of
 *   torch.rand(n) x n_factors ; torch.randn(n, nd_x) ; torch.randn(n, nd_c) ; torch.randn(n, nd_y)
 * at generator (seed, offset_in); *offset_out is the generator offset after those draws.
 *   w, b        surrogate weights (nn.Linear layouts, layer after layer) and biases, DEVICE pointers
 *   workspace   dpivae_datagen_workspace_bytes(desc, n) bytes: two activation buffers n x max width */
typedef struct {
  int32_t n_factors, n_layers, nd_x, nd_c, nd_y, _pad;
  int32_t dims[DPIVAE_MAX_PHYS_LAYERS + 2];   /* n_factors, hidden ..., nd_x */
  float lo[16], hi[16], in_mean[16], in_std[16];
  int32_t idx_c[DPIVAE_MAX_NDCY], idx_y[DPIVAE_MAX_NDCY];
  float sigma_x, sigma_c, sigma_y, _pad2;
} dpivae_datagen_desc_t;
size_t dpivae_datagen_workspace_bytes(const dpivae_datagen_desc_t* desc, int64_t n);
int dpivae_sample_response(const dpivae_datagen_desc_t* desc, const float* w, const float* b, int64_t n, uint64_t seed,
                           uint64_t offset_in, int32_t sm_count, int32_t max_threads_per_sm, float* z, float* x, float* c,
                           float* y, void* workspace, size_t workspace_bytes, void* stream, uint64_t* offset_out);

/* Philox bookkeeping for rng mode 1: given the torch CUDA generator's current offset and the SM
 * count / max threads per SM of the device, fill rng->offset / grid_threads for the draws of one
 * forward (P: 3 tensors, S: 1, +1 if cond) and return the generator offset after them. */
uint64_t dpivae_philox_plan(dpivae_handle_t h, int64_t B_global, int32_t n_mc, int32_t cond,
                            uint64_t offset_in, int32_t sm_count, int32_t max_threads_per_sm,
                            dpivae_rng_t* rng);

/* Measurement hooks (bench.py): with timing enabled every hot-path call brackets each of its kernels
 * with CUDA events on the caller's stream; dpivae_last_kernel_ms synchronises those events and returns
 * the durations [enc_fwd, dec_fused, enc_bwd, reduce, adam, lat_fwd, lat_bwd] (7 floats) of the last call in
 * milliseconds (lat_*: the per-pair latent kernels of the tensor-core path, 0 in fp32 mode). */
int dpivae_set_timing(dpivae_handle_t h, int32_t enable);
/* Profiling: device buffer of 16 uint64 cycle counters accumulated per phase of the fused decoder kernel
 * (thread 0 of every CTA, clock64 after each barrier); NULL switches it off. */
int dpivae_set_phase_buffer(dpivae_handle_t h, void* dev_counters16);
int dpivae_last_kernel_ms(dpivae_handle_t h, float* out7);

/* FP32 FFMA peak of the current device (micro-benchmark, 2 FLOP per FFMA), in TFLOP/s: the roofline
 * denominator for the fp32-parity mode, which MEASURED_PEAKS.json does not carry. */
int dpivae_ffma_peak_tflops(float* tflops_out, void* stream);

/* Arithmetic of the decoder-side GEMMs (MLP stacks of models/decoders.py, models/nn.py).
 *   DPIVAE_MATH_FP32     : fp32 FFMA on the CUDA cores (default; bit-reproducible, 1e-5 parity).
 *   DPIVAE_MATH_TC_FP16X3: tcgen05 tensor cores, every fp32 operand split into two fp16 planes and each
 *                          GEMM issued as hi*hi + lo*hi + hi*lo into an fp32 TMEM accumulator
 *                          (fp32-level accuracy; parity tolerance stated in tests/test_gpu_tc.py).
 *   DPIVAE_MATH_TC_FP16  : tcgen05 with plain fp16 operands, fp32 accumulate (reduced precision, separately
 *                          stated tolerance).
 * Shapes / options the tensor-core kernel does not cover (n_mc outside [8,128], cond, lambda_x, per-sample
 * decoder outputs, encode-only) run on the fp32 kernel whatever the mode; dpivae_last_used_tensor_cores()
 * says which kernel the last call used. */
enum { DPIVAE_MATH_FP32 = 0, DPIVAE_MATH_TC_FP16X3 = 1, DPIVAE_MATH_TC_FP16 = 2 };
int dpivae_set_math_mode(dpivae_handle_t h, int32_t mode);
int dpivae_last_used_tensor_cores(dpivae_handle_t h);

/* Number of kernels launched by the last dpivae_loss / train_step / encode call on this handle. */
int dpivae_last_launch_count(dpivae_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* DPIVAE_B200_H */
