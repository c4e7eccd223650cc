// Internal kernel-parameter structs shared by api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpv {

constexpr int MAXZ = 16;
constexpr int MAXL = MAXZ * (MAXZ + 1) / 2;  // packed lower-triangular entries incl. diagonal
constexpr int RBMAX = 16;                    // rows per row block in the decoder kernel
constexpr int NSCAL = 8;                     // per-CTA scalar partials appended to the grad partials

// One-hidden-layer MLP staged in shared memory (weights transposed: Wt[k][n], n contiguous).
struct Mlp2S {
  int K0, H, O;            // true dims
  int ldw0, ldw1;          // leading dims of w0t / w1t
  int s_w0t, s_b0, s_w1t, s_b1;  // smem offsets (floats)
  long long g_w0, g_b0, g_w1, g_b1;  // offsets in the flat param / grad buffers
};

struct PhysLayerS {
  int K, N, ldw, s_wt, s_b;
  long long g_w, g_b;      // offsets in the frozen-weights device buffer
};

// Device-resident values of a captured (CUDA-graph) training step: advanced on the device by advance_kernel, so a
// replayed graph needs no per-step host arguments (dpivae.py:390-436 loop counter, Adam bias corrections, generator offset)
struct StepState {
  long long step;                      // 1-based optimizer step of the CURRENT replay
  unsigned long long philox_off[4];    // generator offsets of the current step's noise tensors
  float step_size[16];                 // lr_g / (1 - beta1^step)
  float bc2_sqrt;                      // sqrt(1 - beta2^step)
  float _pad;
};

struct RngP {
  const StepState* ss;     // non-null: Philox offsets come from the device-resident step state
  int mode;                // 0 injected buffers, 1 philox (torch.cuda normal_ stream)
  const float* eps[4];
  unsigned long long seed;
  unsigned long long offset[4];
  unsigned int grid_threads[4];
};

struct OutP {
  float* row_loss;
  float *xh_p, *xh_d, *ch, *lsc, *yh, *lsy, *zx, *zc, *zy, *dens;
};

// ---- decoder-side fused kernel -------------------------------------------------------------
struct DecParams {
  // model
  int model_type, nz_x, nz_c, nz_y, Z, nd_x, nd_c, nd_y, nd_p;
  int idx_c_phys[4];
  int n_blk, blk_start[3], blk_size[3], blk_loff[3];  // latent blocks (P: x,c,y ; S: one)
  int nL;                                              // packed L entries (all blocks)
  unsigned char L_blk[MAXL], L_i[MAXL], L_j[MAXL];
  float lb[4], ub[4];
  int prior_kind[4];
  float prior_a[4], prior_b[4];
  float lambda_g0;
  int has_lambda_x;
  float lambda_x;
  int phys_kind, phys_n_layers;
  PhysLayerS pl[6];
  float phys_in_mean[8], phys_in_std[8];
  float grid[64];
  Mlp2S fx, dc, dy;
  long long g_lsx;              // offset of log_sigma_x in the flat buffers
  int henc[3], hpri[2], O_tot;  // feature rows in headpre / gpre
  // --full_cov_prior (dpivae.py:151-153): the conditional priors carry a full lower-triangular factor; strict-lower entries of
  // prior `which` at rp_pL / f_pL + pl_off[which] + i (i - 1) / 2 + j, head rows hpri[which] + [mean nz | sigma nz | cov nz * nz]
  int prior_full, npL, pl_off[2], rp_pL, f_pL;
  // smem plan (float offsets); rows have leading dimension LDP
  int s_zero_end;               // everything below is zero-filled once
  int s_A[4];                   // physics MLP hidden activations (layer outputs 0..n-2)
  int s_XHP, s_HD, s_XHD;
  int s_EPS, s_EPSC, s_U, s_ZXIN, s_S0, s_ZD, s_OC, s_OY, s_DZD, s_DZC, s_DZY, s_DZX;
  int s_SC;                     // scalar rows: KLP, RXP, RCP, RYP, REGP, WP, LSXP + 4 partial rows
  int s_ROWPAR, s_ROWRAW, s_ROWACC, s_FEAT, s_ROWX, s_PH;
  int rp_loc, rp_L, rp_pmu, rp_psig, n_rowpar;
  int f_loc, f_L, f_pmu, f_psig, n_feat;  // per-pair gradient feature rows
  int s_total;
  // runtime
  const float* params;
  const float* frozen;
  const float *x, *c, *y;
  const long long* idx;
  long long B, Bg, row_off;
  long long row_stride;         // global row of local row r = row_off + r * row_stride (>= 1)
  int n_mc, cond, with_grad, RB, n_chunks, latent_only;
  long long n_rowblocks;
  RngP rng;
  float beta_x, alpha_x, alpha_c, alpha_y;
  const float* headpre;   // [O_tot][B]
  float* gpre;            // [O_tot][B]
  float* part;            // per-CTA partial grads + scalars
  long long part_stride;
  long long n_params;
  OutP out;
  long long* phase;        // optional [PH_COUNT] cycle counters (profiling), else nullptr
  // tensor-core path: per-tile records exchanged between lat_fwd_kernel / dec_tc_kernel / lat_bwd_kernel
  unsigned char* rec;      // [tiles][rec_stride]: latent operand hi/lo planes (8 KB) + raw c|y per pair
  long long rec_stride;
  float* dzrec;            // [tiles][nz_c + nz_y + nz_x][128]: dL/dz per pair
  float* epsbuf;           // [tiles][Z][128]: reparameterisation noise
  float* rowkl;            // [B]: per-row KL (MC mean)
  unsigned int* gpre_max;  // bits of max |gpre| over the encoder heads of the batch (lat_bwd -> enc_tc_bwd operand scale), or nullptr
  // decode-only calls (DPIVAE.decode, models/vae.py:153-158): user latents (n, B, .) replace the encoder's
  const float *zin_x, *zin_c, *zin_y;
  // thread-per-pair latent kernels (lat_kernels.cu): noise in the LOCAL (m, row, i) order of each noise tensor, one buffer
  // per latent block; eps_ready = already filled by lat_noise_fill_kernel (else the forward fills it when a backward follows)
  float* eps_local[3];
  int eps_ready;
  // cyclic noise pre-pass (lat_noise_fill_cyclic_kernel): per-block launch constants, computed once on the host
  unsigned int cyc_nloc[3], cyc_gtn[3], cyc_step[3];   // GT / S generator threads of this rank, GT / nz, GT / (nz S)
};

// DPIVAE.prior_net post-processing and GaussianEncoder.sample on given (loc, scale_tril) (optim_kernels.cu)
void launch_prior_post(const float* headpre, long long B, int row0, int nz, int full, float* loc, float* tril, cudaStream_t s);
void launch_gaussian_sample(const float* loc, const float* tril, const float* eps, int n, long long B, int nz, float* z,
                            float* dens, cudaStream_t s);

// ---- tensor-core decoder kernel (dec_tc_kernel.cu): DecParams + its own shared-memory plan (BYTE offsets) ----
struct TcParams {
  DecParams d;
  int terms;                 // 3 = fp16 hi/lo split (fp32-accurate), 1 = plain fp16 inputs
  int KZ, c_ones, c_s0;      // latent operand: columns, constant-one column, first physics-input column
  int w_fx0, w_fx1, w_p[4];  // weight operands (hi plane), lo plane at + l_*
  int l_fx0, l_fx1, l_p[4];
  int a_big, a_g, l_big, l_g;   // activation / gradient operands (the auxiliary decoders' mask / product operands alias a_g)
  int a_rec, rec_buf;        // two tile-record buffers of rec_buf bytes each
  int f_inv, f_bias_x, f_bias_p1, f_bias_p2, f_aw0, f_ab0, f_aw1, f_ab1, f_w0f, f_wp0f;
  int f_dza, f_sc, f_sca, f_red;   // f_dza: [2][nz_c + nz_y][128] aux dL/dz, f_sca: [2][2][128] aux R_c / R_y (double-buffered by tile parity)
  int o_bar;
  int total;
};
void launch_dec_tc(const TcParams& p, int grid, cudaStream_t s);
int configure_dec_tc_kernel();
size_t lat_smem_bytes(const DecParams& p, bool bwd);
void launch_lat_fwd(const DecParams& p, long long n_tiles, cudaStream_t s);
void launch_lat_bwd(const DecParams& p, long long n_tiles, cudaStream_t s);
bool lat_pair_supported(const DecParams& p);
void launch_lat_noise_fill(const DecParams& p, bool cyclic, cudaStream_t s);
int configure_lat_kernels();
void launch_lat_encode(const DecParams& p, cudaStream_t s);
bool dec_tc_has_variant(int phys_kind, int nd_x);

// ---- encoder-side kernels (forward and backward over "MLP2 units") ------------------------------
struct EncUnit {
  int K0, H, O;
  int src;                 // 0 = x, 1 = c, 2 = y
  int hid_row, out_row;    // first feature row in hid / headpre buffers
  long long g_w0, g_b0, g_w1, g_b1;
};

struct EncParams {
  int n_units;
  EncUnit u[5];
  int nd_x, nd_c, nd_y;
  int x_is_standardised;
  float mean_x[64], istd_x[64], mean_c[4], istd_c[4], mean_y[4], istd_y[4];
  float std_x[64], std_c[4], std_y[4];
  const float* params;
  const float *x, *c, *y;
  const long long* idx;
  long long B;
  float* hid;       // [H_tot][B]
  float* headpre;   // [O_tot][B]
  const float* gpre;
  float* part;
  long long part_stride;
  long long n_params;
  int with_hid;     // forward: also store hidden activations (needed by the backward)
};

// ---- tensor-core encoder kernels (enc_tc_kernels.cu) ----------------------------------------------
struct EncTcParams {
  int n_units, K0, KX, Hc, Oc, terms;
  int H[3], O[3], h_off[4], o_off[3], out_row[3];
  long long g_w0[3], g_b0[3], g_w1[3], g_b1[3];
  float mean_x[64], std_x[64];
  int x_is_standardised;
  const float* params;
  const float* x;
  const long long* idx;
  long long B;
  float* headpre;
  unsigned char* hidrec;   // optional per-tile record of the hidden activations (X8 hi/lo planes) for the backward
  long long hid_stride;
  int hid_lo;
  // shared-memory plan (bytes)
  int w_0, l_0, w_1, l_1, a_x, l_x, f_b1, f_red, f_orow, o_bar, total;
  int f_raw, fb_raw;       // raw-weight scratch of the set-up (byte offset, -1 = no room: read the flat buffer directly)
  // backward kernel: inputs, plan
  const float* gpre;
  float* part;
  long long part_stride;
  int e_g;                 // gpre is scaled by 2^e_g before the fp16 split (static estimate from the batch size) ...
  const unsigned int* gpre_max;   // ... unless lat_bwd_kernel measured max |gpre| of this batch: then 2^e_g max = 2^11
  int wb_1, lb_1, ab_h, ab_g, lb_g, ab_x, lb_x, fb_red, fb_orow, ob_bar, total_b;
};
void launch_enc_tc_bwd(const EncTcParams& p, int grid, cudaStream_t s);

// ---- fused encode-only kernel (enc_fused_kernel.cu): encoder MMAs + latent sampling, one launch ----
struct EncFusedParams {
  EncTcParams q;
  float istd_x[64];
  int model_type, nz[3];
  int hn0[3], hN[3];           // head MMA of unit u: accumulator columns [hn0, hn0 + hN) (8-aligned start, multiple of 16 wide)
  float lb[4], ub[4];
  RngP rng;
  long long Bg, row_off, row_stride;
  int n_mc;
  int o_ms, o_bars;            // byte offsets: scaler statistics [2][64] floats, mbarrier block (both after the EncTcParams plan)
  float *zx, *zc, *zy, *dens;  // (n_mc, B, nz_*) latents and (n_mc, B) density, any may be null
  long long* phase;            // optional [16] cycle counters of CTA 0 (tools/encode_probe.py), else nullptr
  int dbg;                     // PROF instantiation only: 1 skip x staging, 2 skip ReLU epilogue, 4 skip latent math (timing experiments)
  long long* trace;            // optional [6][402] event trace of CTA 0 (phase + 32), else nullptr
  float* eps_local[3];         // pre-generated noise per latent block, local (n_mc, B, nz_b) order (noise_fill_kernel), or null
};
bool enc_fused_supports(const EncFusedParams& p);
int launch_enc_fused(const EncFusedParams& p, int grid, cudaStream_t s);
int configure_enc_fused_kernels();
void launch_enc_tc_fwd(const EncTcParams& p, int grid, cudaStream_t s);
int configure_enc_tc_kernels();

// ---- reduce + Adam ------------------------------------------------------------------------------
struct AdamParams {
  float* params;
  const float* grads;
  float *m, *v;
  const unsigned char* group;   // per param group id
  float step_size[16];          // lr / (1 - beta1^t)
  float wd[16];
  float bc2_sqrt;
  float beta1, beta2, eps;
  long long n_params;
  const float* clip_coef;       // device scalar or nullptr
  // captured-step mode: step sizes from the device state; per-step log row = 8 loss scalars + log_sigma_x
  const StepState* ss;
  const float* scalars;
  float* log;                   // [log_cap][9] ring, or nullptr
  long long log_cap;
  long long lsx_index;
};

struct ReduceParams {
  const float* part;
  long long part_stride;
  int n_cta[3];                 // number of CTA partials per owner class
  long long base[3];            // first partial vector (in units of part_stride) of each owner class
  long long n_params;
  const unsigned char* owner;   // per param: 0 = decoder kernel, 1 = prior-net units, 2 = encoder units
  float* grads;
  float* scalars;               // 8 floats
  float inv_B, inv_BD;
  // single-shard train step without gradient clipping: the Adam update of a parameter is applied by the thread that just
  // reduced its gradient (one launch and one pass over the gradients less); `adam` is ignored unless fuse_adam != 0
  int fuse_adam;
  AdamParams adam;
};

struct AdvanceParams {
  StepState* ss;
  unsigned long long philox_inc;    // generator offset consumed by one step
  float lr[16];
  int n_groups;
  const long long* idx_pool;        // [pool_rows][B] minibatch indices, or nullptr
  long long pool_rows, B;
  long long* idx_cur;               // [B] indices of the current step (read by the step's kernels)
};
void launch_advance(const AdvanceParams& p, cudaStream_t s);

// ---- post-processing on the device (metrics_kernels.cu) -----------------------------------------
void launch_mc_mean(const float* v, int n, long long BD, float* out, cudaStream_t s);
void launch_regression_metrics(const float* y, const float* p, long long N, int d, double* scratch, float* out3, cudaStream_t s);
void launch_linreg_r2(const float* Xtr, const float* ytr, long long ldy_tr, long long Ntr, const float* Xte, const float* yte,
                      long long ldy_te, long long Nte, int k, double* scratch, float* r2, cudaStream_t s);

// ---- on-device synthetic data generator (datagen_kernels.cu) -------------------------------------
struct DataGenParams {
  long long n;
  int nf, nd_x, nd_c, nd_y;
  float lo[16], hi[16], in_mean[16], in_std[16];
  int idx_c[4], idx_y[4];
  float sigma_x, sigma_c, sigma_y;
  unsigned long long seed, off_u[16], off_x, off_c, off_y;
  unsigned int T_u, T_x, T_c, T_y;
  float *z, *a0, *x, *c, *y;
};
void launch_datagen_latents(const DataGenParams& p, cudaStream_t s);
void launch_mlp_layer(const float* A, const float* W, const float* b, float* O, long long n, int K, int N, bool tanh_act, cudaStream_t s);
void launch_datagen_finish(const DataGenParams& p, cudaStream_t s);

size_t dec_smem_bytes(const DecParams& p);
void launch_dec(const DecParams& p, int grid, cudaStream_t s);
// overlap = true: programmatic dependent launch that runs CONCURRENTLY with the kernel before it on the stream (the
// caller guarantees that the two are independent; the kernels wait for their predecessor only at their very end)
void launch_enc_fwd(const EncParams& p, int grid, size_t smem, cudaStream_t s, bool overlap = false);
void launch_enc_bwd(const EncParams& p, int grid, size_t smem, cudaStream_t s, bool overlap = false);
size_t enc_smem_bytes(const EncParams& p, bool bwd);
// dedicated kernels of the two conditional-prior nets (prior_kernels.cu)
bool prior_kernels_support(const EncParams& prior_units);
void launch_prior_fwd(const EncParams& prior_units, int sm_count, cudaStream_t s, bool overlap = false);
void launch_prior_bwd(const EncParams& prior_units, int grid, cudaStream_t s, bool overlap = false);   // grid = number of per-CTA partial rows
void launch_reduce(const ReduceParams& p, cudaStream_t s);
void launch_gradnorm(const float* grads, long long n, float max_norm, float* clip_coef, cudaStream_t s);
void launch_adam(const AdamParams& p, cudaStream_t s);
int configure_kernels();
float ffma_peak_tflops(cudaStream_t s);

}  // namespace dpv
