// C ABI of libdpivae_b200.so (include/dpivae_b200.h): handle management, shared-memory /
// workspace planning and the launch sequences of the hot path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dpivae_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace dpv {
int configure_dec_kernel();
int configure_enc_kernels();
}  // namespace dpv


using namespace dpv;

static thread_local std::string g_err;
static int fail(const std::string& m) {
  g_err = m;
  return 1;
}
#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct dpivae_model {
  dpivae_model_desc_t d;
  int sm_count = 0;
  DecParams dec;
  EncParams enc;
  int n_enc_units = 0;     // encoder units come first in enc.u, prior units after
  int H_tot = 0, O_tot = 0;
  float* d_frozen = nullptr;
  unsigned char* d_owner = nullptr;
  unsigned char* d_group = nullptr;
  float* d_clip = nullptr;
  long long* d_phase = nullptr;
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
  int n_groups = 0;
  float lr[16], wd[16];
  int last_launches = 0;
  long long part_stride = 0;
  EncTcParams enc_tc;       // tensor-core encoder plan
  EncFusedParams enc_fused; // fused encode-only kernel (encoder MMAs + latent sampling)
  int enc_fused_ok = 0;
  int enc_tc_ok = 0, enc_tc_bwd_ok = 0;
  int math_mode = 0;        // DPIVAE_MATH_*
  TcParams tc;              // tensor-core decoder plan
  int tc_ok = 0;            // model shape supported by dec_tc_kernel
  int last_dec_tc = 0;      // the last hot-path call ran the tensor-core decoder kernel
  int timing = 0;
  cudaEvent_t ev[14] = {};   // start/stop pairs: enc_fwd, dec, enc_bwd, reduce, adam, lat_fwd, lat_bwd
  int ev_used[7] = {0, 0, 0, 0, 0, 0, 0};
  // set while a step graph is being captured: kernels read step / offsets from this device-resident state
  StepState* cur_ss = nullptr;
  float* cur_log = nullptr;
  long long cur_log_cap = 0;
};

// A captured training step (advance -> loss fwd/bwd -> Adam) replayable without host arguments.
struct dpivae_step_graph {
  dpivae_model* h = nullptr;
  StepState* d_state = nullptr;
  cudaGraph_t graph = nullptr, graph_u = nullptr;
  cudaGraphExec_t exec = nullptr, exec_u = nullptr;   // one step / `unroll` consecutive steps
  int unroll = 1;
  uint64_t philox_inc = 0;
  int launches_per_step = 0;
};

struct KTimer {
  dpivae_model* h; int k; cudaStream_t st;
  KTimer(dpivae_model* h_, int k_, cudaStream_t st_) : h(h_), k(k_), st(st_) {
    if (h->timing) { cudaEventRecord(h->ev[2 * k], st); h->ev_used[k] = 1; }
  }
  ~KTimer() { if (h->timing) cudaEventRecord(h->ev[2 * k + 1], st); }
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int check_mlp2(const dpivae_mlp2_t& m, int in_dim, int hid, int out_dim, long long n_params, const char* name) {
  if (m.in_dim != in_dim || m.hid != hid || m.out_dim != out_dim)
    return fail(std::string("unit ") + name + ": unexpected dims");
  const long long ends[4] = {m.w0 + (long long)in_dim * hid, m.b0 + hid, m.w1 + (long long)hid * out_dim, m.b1 + out_dim};
  const long long begs[4] = {m.w0, m.b0, m.w1, m.b1};
  for (int i = 0; i < 4; ++i)
    if (begs[i] < 0 || ends[i] > n_params) return fail(std::string("unit ") + name + ": offsets out of range");
  return 0;
}

static void fill_mlp2s(Mlp2S& s, const dpivae_mlp2_t& m, int& o) {
  s.K0 = m.in_dim; s.H = m.hid; s.O = m.out_dim;
  s.ldw0 = m.hid + 4;
  s.ldw1 = pad4(m.out_dim) + 4;
  s.s_w0t = o; o += pad4(m.in_dim) * s.ldw0;
  s.s_b0 = o; o += m.hid;
  s.s_w1t = o; o += m.hid * s.ldw1;
  s.s_b1 = o; o += pad4(m.out_dim);
  s.g_w0 = m.w0; s.g_b0 = m.b0; s.g_w1 = m.w1; s.g_b1 = m.b1;
}

static int build_plan(dpivae_model* h) {
  const dpivae_model_desc_t& d = h->d;
  DecParams& P = h->dec;
  memset(&P, 0, sizeof(P));
  const int Z = d.nz_x + d.nz_c + d.nz_y;
  if (d.nz_x < 1 || d.nz_x > DPIVAE_MAX_ZX || d.nz_c < 1 || d.nz_c > DPIVAE_MAX_ZCY || d.nz_y < 1 ||
      d.nz_y > DPIVAE_MAX_ZCY || Z > DPIVAE_MAX_Z)
    return fail("latent dimensions out of the supported range");
  if (d.nd_x < 4 || d.nd_x > DPIVAE_MAX_NDX || d.nd_c < 1 || d.nd_c > DPIVAE_MAX_NDCY || d.nd_y < 1 ||
      d.nd_y > DPIVAE_MAX_NDCY || d.nd_p < 0 || d.nd_p > d.nd_c)
    return fail("data dimensions out of the supported range");
  if (d.model_type != DPIVAE_MODEL_P && d.model_type != DPIVAE_MODEL_S) return fail("Invalid model_type");
  P.model_type = d.model_type;
  P.nz_x = d.nz_x; P.nz_c = d.nz_c; P.nz_y = d.nz_y; P.Z = Z;
  P.nd_x = d.nd_x; P.nd_c = d.nd_c; P.nd_y = d.nd_y; P.nd_p = d.nd_p;
  for (int i = 0; i < 4; ++i) P.idx_c_phys[i] = d.idx_c_phys[i];
  for (int i = 0; i < d.nd_p; ++i)
    if (d.idx_c_phys[i] < 0 || d.idx_c_phys[i] >= d.nd_c) return fail("idx_c_phys out of range");
  if (d.model_type == DPIVAE_MODEL_P) {
    P.n_blk = 3;
    P.blk_start[0] = 0; P.blk_size[0] = d.nz_x;
    P.blk_start[1] = d.nz_x; P.blk_size[1] = d.nz_c;
    P.blk_start[2] = d.nz_x + d.nz_c; P.blk_size[2] = d.nz_y;
  } else {
    P.n_blk = 1;
    P.blk_start[0] = 0; P.blk_size[0] = Z;
  }
  int nL = 0, hrow = 0;
  for (int b = 0; b < P.n_blk; ++b) {
    const int nzb = P.blk_size[b];
    P.blk_loff[b] = nL;
    for (int i = 0; i < nzb; ++i)
      for (int j = 0; j <= i; ++j) {
        P.L_blk[nL] = (unsigned char)b; P.L_i[nL] = (unsigned char)i; P.L_j[nL] = (unsigned char)j;
        ++nL;
      }
    P.henc[b] = hrow;
    hrow += 2 * nzb + nzb * nzb;
  }
  P.nL = nL;
  // conditional prior nets: FactorizedNN heads [mean | sigma] or, with --full_cov_prior, FullCovarianceNN heads [mean | sigma | cov]
  const bool pfull = d.prior[0].out_dim == 2 * d.nz_c + d.nz_c * d.nz_c && d.prior[1].out_dim == 2 * d.nz_y + d.nz_y * d.nz_y;
  P.prior_full = pfull ? 1 : 0;
  P.pl_off[0] = 0; P.pl_off[1] = d.nz_c * (d.nz_c - 1) / 2;
  P.npL = pfull ? P.pl_off[1] + d.nz_y * (d.nz_y - 1) / 2 : 0;
  P.hpri[0] = hrow; hrow += 2 * d.nz_c + (pfull ? d.nz_c * d.nz_c : 0);
  P.hpri[1] = hrow; hrow += 2 * d.nz_y + (pfull ? d.nz_y * d.nz_y : 0);
  P.O_tot = hrow;
  h->O_tot = hrow;
  for (int i = 0; i < DPIVAE_MAX_ZX; ++i) {
    P.lb[i] = d.lb[i]; P.ub[i] = d.ub[i];
    P.prior_kind[i] = d.prior_kind[i]; P.prior_a[i] = d.prior_a[i]; P.prior_b[i] = d.prior_b[i];
  }
  P.lambda_g0 = d.lambda_g0; P.has_lambda_x = d.has_lambda_x; P.lambda_x = d.lambda_x;
  P.phys_kind = d.phys_kind;
  for (int i = 0; i < DPIVAE_MAX_NDX; ++i) P.grid[i] = d.phys_grid[i];
  P.g_lsx = d.log_sigma_x;
  if (d.log_sigma_x < 0 || d.log_sigma_x >= d.n_params) return fail("log_sigma_x offset out of range");

  // ---- validate units ----
  const int nzd = d.nz_c + d.nz_y, nzin = d.nz_x + d.nd_p;
  if (check_mlp2(d.fx, nzd, 128, d.nd_x, d.n_params, "decoder_x")) return 1;
  if (check_mlp2(d.dec_c, d.nz_c, 64, 2 * d.nd_c, d.n_params, "decoder_c")) return 1;
  if (check_mlp2(d.dec_y, d.nz_y, 64, 2 * d.nd_y, d.n_params, "decoder_y")) return 1;
  if (check_mlp2(d.prior[0], d.nd_c, 64, 2 * d.nz_c + (pfull ? d.nz_c * d.nz_c : 0), d.n_params, "prior_net_c")) return 1;
  if (check_mlp2(d.prior[1], d.nd_y, 64, 2 * d.nz_y + (pfull ? d.nz_y * d.nz_y : 0), d.n_params, "prior_net_y")) return 1;
  if (d.model_type == DPIVAE_MODEL_P) {
    const int nzs[3] = {d.nz_x, d.nz_c, d.nz_y};
    for (int e = 0; e < 3; ++e)
      if (check_mlp2(d.enc[e], d.nd_x, 64, 2 * nzs[e] + nzs[e] * nzs[e], d.n_params, "encoder")) return 1;
  } else {
    if (check_mlp2(d.enc[0], d.nd_x, 128, 2 * Z + Z * Z, d.n_params, "encoder")) return 1;
  }

  // ---- shared-memory plan of the decoder kernel ----
  int o = 0;
  fill_mlp2s(P.fx, d.fx, o);
  fill_mlp2s(P.dc, d.dec_c, o);
  fill_mlp2s(P.dy, d.dec_y, o);
  int act_rows = 0;
  const int act_start_marker = -1;
  (void)act_start_marker;
  if (d.phys_kind == DPIVAE_PHYS_MLP) {
    if (d.phys_n_layers < 1 || d.phys_n_layers > 5) return fail("physics MLP: 1..5 linear layers supported");
    if (d.phys_dims[0] != nzin || d.phys_dims[d.phys_n_layers] != d.nd_x) return fail("physics MLP: in/out dims mismatch");
    P.phys_n_layers = d.phys_n_layers;
    long long gw = 0;
    for (int l = 0; l < d.phys_n_layers; ++l) {
      PhysLayerS& L = P.pl[l];
      L.K = d.phys_dims[l]; L.N = d.phys_dims[l + 1];
      if (L.K < 1 || L.N < 1 || L.K > 128 || L.N > 128) return fail("physics MLP: layer width out of range");
      L.ldw = pad4(L.N) + 4;
      L.s_wt = o; o += pad4(L.K) * L.ldw;
      L.s_b = o; o += pad4(L.N);
      L.g_w = gw; gw += (long long)L.K * L.N;
      L.g_b = gw; gw += L.N;
    }
  } else if (d.phys_kind == DPIVAE_PHYS_MASS_SPRING) {
    if (d.nz_x != 1 || d.nd_p != 0) return fail("mass_spring physics expects nz_x = 1, nd_p = 0");
  } else if (d.phys_kind == DPIVAE_PHYS_BEAM) {
    if (d.nz_x != 2 || d.nd_p != 0) return fail("beam physics expects nz_x = 2, nd_p = 0");
  } else {
    return fail("unknown physics decoder kind");
  }
  const int act_start = o;
  if (d.phys_kind == DPIVAE_PHYS_MLP)
    for (int l = 0; l + 1 < d.phys_n_layers; ++l) {
      P.s_A[l] = o; o += pad4(d.phys_dims[l + 1]) * LDP; act_rows += pad4(d.phys_dims[l + 1]);
    }
  P.s_XHP = o; o += pad4(d.nd_x) * LDP; act_rows += pad4(d.nd_x);
  P.s_HD = o; o += 128 * LDP; act_rows += 128;
  P.s_XHD = o; o += pad4(d.nd_x) * LDP; act_rows += pad4(d.nd_x);
  P.s_FEAT = act_start;
  P.rp_loc = 0; P.rp_L = Z; P.rp_pmu = Z + nL; P.rp_psig = P.rp_pmu + nzd; P.rp_pL = P.rp_psig + nzd; P.n_rowpar = P.rp_pL + P.npL;
  P.f_loc = 0; P.f_L = Z; P.f_pmu = Z + nL; P.f_psig = P.f_pmu + nzd; P.f_pL = P.f_psig + nzd; P.n_feat = P.f_pL + P.npL;
  if (P.n_feat > act_rows) return fail("internal: gradient feature rows exceed the activation region");
  auto rows = [&](int r) { int s = o; o += r * LDP; return s; };
  P.s_EPS = rows(pad4(Z));
  P.s_EPSC = rows(pad4(d.nz_c));
  P.s_U = rows(pad4(d.nz_x));
  P.s_ZXIN = rows(pad4(nzin));
  P.s_S0 = rows(pad4(nzin));
  P.s_ZD = rows(pad4(nzd) + 4);
  P.s_OC = rows(pad4(2 * d.nd_c));
  P.s_OY = rows(pad4(2 * d.nd_y));
  P.s_DZD = rows(pad4(nzd));
  P.s_DZC = rows(pad4(d.nz_c));
  P.s_DZY = rows(pad4(d.nz_y));
  P.s_DZX = rows(pad4(nzin));
  P.s_SC = rows(16);
  P.s_ROWPAR = o; o += P.n_rowpar * RBMAX;
  P.s_ROWRAW = o; o += (d.nd_c + d.nd_y) * RBMAX;
  P.s_ROWACC = o; o += pad4(P.n_feat + 5);          // multi-chunk blocks only (one row per block)
  P.s_ROWX = o; o += d.nd_x * RBMAX;
  P.s_PH = o; o += 32;                              // 16 x int64 phase counters
  if (P.n_feat * LDP + (P.n_feat + 5) * RBMAX > act_rows * LDP) return fail("internal: row accumulators exceed the activation region");
  P.s_total = o;
  P.s_zero_end = o;
  if ((size_t)o * sizeof(float) > 232448) {
    char buf[128];
    snprintf(buf, sizeof buf, "decoder kernel shared-memory plan needs %zu bytes (> 232448)", (size_t)o * 4);
    return fail(buf);
  }
  P.n_params = d.n_params;

  // ---- encoder-side units ----
  EncParams& E = h->enc;
  memset(&E, 0, sizeof(E));
  int nu = 0, hid_row = 0;
  const int n_enc = d.model_type == DPIVAE_MODEL_P ? 3 : 1;
  for (int e = 0; e < n_enc; ++e) {
    EncUnit& U = E.u[nu++];
    U.K0 = d.enc[e].in_dim; U.H = d.enc[e].hid; U.O = d.enc[e].out_dim; U.src = 0;
    U.hid_row = hid_row; hid_row += U.H;
    U.out_row = P.henc[e];
    U.g_w0 = d.enc[e].w0; U.g_b0 = d.enc[e].b0; U.g_w1 = d.enc[e].w1; U.g_b1 = d.enc[e].b1;
  }
  h->n_enc_units = nu;
  for (int k = 0; k < 2; ++k) {
    EncUnit& U = E.u[nu++];
    U.K0 = d.prior[k].in_dim; U.H = d.prior[k].hid; U.O = d.prior[k].out_dim; U.src = 1 + k;
    U.hid_row = hid_row; hid_row += U.H;
    U.out_row = P.hpri[k];
    U.g_w0 = d.prior[k].w0; U.g_b0 = d.prior[k].b0; U.g_w1 = d.prior[k].w1; U.g_b1 = d.prior[k].b1;
  }
  E.n_units = nu;
  h->H_tot = hid_row;
  E.nd_x = d.nd_x; E.nd_c = d.nd_c; E.nd_y = d.nd_y;
  for (int i = 0; i < d.nd_x; ++i) { E.mean_x[i] = d.mean_x[i]; E.std_x[i] = d.std_x[i]; E.istd_x[i] = 1.0f / d.std_x[i]; }
  for (int i = 0; i < d.nd_c; ++i) { E.mean_c[i] = d.mean_c[i]; E.std_c[i] = d.std_c[i]; E.istd_c[i] = 1.0f / d.std_c[i]; }
  for (int i = 0; i < d.nd_y; ++i) { E.mean_y[i] = d.mean_y[i]; E.std_y[i] = d.std_y[i]; E.istd_y[i] = 1.0f / d.std_y[i]; }
  E.n_params = d.n_params;
  if (enc_smem_bytes(E, true) > 232448) return fail("encoder kernel shared-memory plan exceeds 227 KB");
  h->part_stride = (long long)align_up((size_t)d.n_params + NSCAL, 64);

  // ---- tensor-core encoder plan ----
  {
    EncTcParams& Q = h->enc_tc;
    memset(&Q, 0, sizeof(Q));
    h->enc_tc_ok = 0;
    const int nu = h->n_enc_units;
    Q.n_units = nu; Q.K0 = d.nd_x; Q.KX = d.nd_x + 16; Q.terms = 3;
    int hoff = 0, ooff = 0;
    bool ok = (d.nd_x == 32 || d.nd_x == 64);
    for (int u = 0; u < nu; ++u) {
      const EncUnit& U = E.u[u];
      Q.H[u] = U.H; Q.O[u] = U.O; Q.h_off[u] = hoff; Q.o_off[u] = ooff; Q.out_row[u] = U.out_row;
      Q.g_w0[u] = U.g_w0; Q.g_b0[u] = U.g_b0; Q.g_w1[u] = U.g_w1; Q.g_b1[u] = U.g_b1;
      hoff += U.H; ooff += U.O;
      ok = ok && U.K0 == d.nd_x && (U.H % 16 == 0);
    }
    Q.h_off[nu] = hoff;
    Q.Hc = hoff; Q.Oc = (ooff + 15) & ~15;
    ok = ok && Q.Hc % 16 == 0 && Q.Hc <= 256 && Q.Oc <= 256 && 2 * Q.Hc + Q.Oc <= 512;
    for (int i = 0; i < d.nd_x; ++i) { Q.mean_x[i] = d.mean_x[i]; Q.std_x[i] = d.std_x[i]; }
    int b = 0;
    auto plane2 = [&](int chunks, int rows, int& off, int& lo) {
      const int bytes = chunks * rows * 16;
      off = b; lo = bytes; b += 2 * bytes;
    };
    plane2(Q.KX / 8, Q.Hc, Q.w_0, Q.l_0);
    plane2(Q.Hc / 8, Q.Oc, Q.w_1, Q.l_1);
    plane2(Q.KX / 8, 128, Q.a_x, Q.l_x);
    Q.f_b1 = b; b += Q.Oc * 4;
    Q.f_orow = b; b += Q.Oc * 4;
    Q.f_red = b; b += 256;
    Q.o_bar = b; b += 16;
    Q.total = (b + 127) & ~127;
    // raw-weight scratch of the set-up: every unit's [w0 | b0 | w1 | b1] block (contiguous in the flat buffer)
    {
      long long raw = 0;
      bool contiguous = true;
      for (int u = 0; u < nu; ++u) {
        const EncUnit& U = E.u[u];
        raw += (long long)U.H * U.K0 + U.H + (long long)U.O * U.H + U.O;
        contiguous = contiguous && U.g_b0 == U.g_w0 + (long long)U.H * U.K0 && U.g_w1 == U.g_b0 + U.H &&
                     U.g_b1 == U.g_w1 + (long long)U.O * U.H;
      }
      ok = ok && contiguous;
      Q.f_raw = -1;
      if (Q.total + raw * 4 <= 232448) { Q.f_raw = Q.total; Q.total = (int)((Q.total + raw * 4 + 127) & ~127LL); }
    }
    Q.hid_lo = (Q.Hc / 8) * 128 * 16;
    Q.hid_stride = 2 * (long long)Q.hid_lo;
    if (ok && Q.total <= 232448) h->enc_tc_ok = 1;
    // fused encode-only kernel: same plan + its mbarrier block
    {
      EncFusedParams& F = h->enc_fused;
      memset(&F, 0, sizeof(F));
      F.q = Q;
      for (int i = 0; i < d.nd_x; ++i) F.istd_x[i] = 1.0f / d.std_x[i];
      F.model_type = d.model_type;
      F.nz[0] = d.nz_x; F.nz[1] = d.nz_c; F.nz[2] = d.nz_y;
      for (int i = 0; i < d.nz_x && i < 4; ++i) { F.lb[i] = h->dec.lb[i]; F.ub[i] = h->dec.ub[i]; }
      // own shared-memory plan: two X operand buffers (the raw-weight scratch of the set-up aliases them)
      {
        EncTcParams& G = F.q;
        int bb = 0;
        auto plane2f = [&](int chunks, int rows, int& off, int& lo) { const int bytes = chunks * rows * 16; off = bb; lo = bytes; bb += 2 * bytes; };
        plane2f(G.KX / 8, G.Hc, G.w_0, G.l_0);
        plane2f(G.Hc / 8, G.Oc, G.w_1, G.l_1);
        plane2f(G.KX / 8, 128, G.a_x, G.l_x);
        bb += 2 * G.l_x;                                   // second X buffer
        G.f_b1 = bb; bb += G.Oc * 4;
        G.f_orow = bb; bb += G.Oc * 4;
        G.f_red = bb; bb += 256;
        G.o_bar = bb; bb += 16;
        G.total = (bb + 127) & ~127;
        long long raw = 0;
        for (int u = 0; u < nu; ++u) raw += (long long)E.u[u].H * E.u[u].K0 + E.u[u].H + (long long)E.u[u].O * E.u[u].H + E.u[u].O;
        G.f_raw = raw * 4 <= 4LL * G.l_x ? G.a_x : -1;
        F.o_ms = G.total;
        F.o_bars = G.total + 512;
        for (int u = 0; u < 3 && u < nu; ++u) {
          F.hn0[u] = G.o_off[u] & ~7;
          F.hN[u] = ((G.o_off[u] + G.O[u] - F.hn0[u]) + 15) & ~15;
          if (F.hn0[u] + F.hN[u] > G.Oc) F.hn0[u] = (G.Oc - F.hN[u]) & ~7;
        }
      }
      h->enc_fused_ok = h->enc_tc_ok && enc_fused_supports(F);
    }
    // backward plan
    b = 0;
    plane2(Q.Hc / 8, Q.Oc, Q.wb_1, Q.lb_1);
    Q.ab_h = b; b += (int)Q.hid_stride;
    plane2((Q.Oc > 64 ? 64 : Q.Oc) / 8, 128, Q.ab_g, Q.lb_g);   // Oc > 64: two column passes through one 64-column buffer
    plane2(Q.KX / 8, 128, Q.ab_x, Q.lb_x);
    Q.fb_orow = b; b += Q.Oc * 4;
    Q.fb_red = b; b += 256;
    Q.ob_bar = b; b += 32;
    Q.total_b = (b + 127) & ~127;
    {
      long long raw = 0;
      for (int u = 0; u < nu; ++u) raw += (long long)E.u[u].O * E.u[u].H;
      Q.fb_raw = -1;
      if (Q.total_b + raw * 4 <= 232448) { Q.fb_raw = Q.total_b; Q.total_b = (int)((Q.total_b + raw * 4 + 127) & ~127LL); }
      else if (raw * 4 <= Q.hid_stride) Q.fb_raw = Q.ab_h;   // the hidden-record buffer is idle until the first tile's bulk copy
    }
    const int nchunk = Q.Hc > 128 ? 2 : 1;
    h->enc_tc_bwd_ok = h->enc_tc_ok && Q.total_b <= 232448 && Q.Oc <= 128 && Q.Hc + nchunk * (Q.Oc + Q.KX) <= 512 &&
                       256 * (Q.Oc > 64 ? 64 : 32) * 4 <= (int)Q.hid_stride;
  }

  // ---- shared-memory plan of the tensor-core decoder kernel (byte offsets) ----
  TcParams& T = h->tc;
  memset(&T, 0, sizeof(T));
  h->tc_ok = 0;
  {
    const bool mlp = d.phys_kind == DPIVAE_PHYS_MLP;
    const int need = nzd + 1 + (mlp ? nzin : 0);
    bool ok = need <= 16 && dec_tc_has_variant(d.phys_kind, d.nd_x) && !d.has_lambda_x && d.nd_c <= 2 && d.nd_y <= 2 &&
              d.nz_c <= 4 && d.nz_y <= 4 && d.nz_x <= 4;
    if (mlp) ok = ok && d.phys_n_layers == 4 && d.phys_dims[1] == 64 && d.phys_dims[2] == 32 && d.phys_dims[3] == 64;
    if (ok) {
      const int KZ = 16, d1 = mlp ? 64 : 0, d2 = mlp ? 32 : 0, d3 = mlp ? 64 : 0;
      T.KZ = KZ; T.c_ones = nzd; T.c_s0 = nzd + 1;
      int b = 0;
      auto plane2 = [&](int chunks, int rows, int& off, int& lo) {   // hi + lo planes
        const int bytes = chunks * rows * 16;
        off = b; lo = bytes; b += 2 * bytes;
      };
      plane2(KZ / 8, 128, T.w_fx0, T.l_fx0);
      if (mlp) {
        // [W_fx1 hi | W_p3 hi | W_fx1 lo | W_p3 lo]: both have nd_x rows, so as MN-major operands the two hi (lo) planes
        // form ONE operand with 128 + d3 columns (fused dgrad of the x head through both decoders)
        const int bf = 16 * d.nd_x * 16, bp = (d3 / 8) * d.nd_x * 16;
        T.w_fx1 = b; T.w_p[3] = b + bf; T.l_fx1 = T.l_p[3] = bf + bp; b += 2 * (bf + bp);
        plane2(KZ / 8, d1, T.w_p[0], T.l_p[0]);
        plane2(d1 / 8, d2, T.w_p[1], T.l_p[1]);
        plane2(d2 / 8, d3, T.w_p[2], T.l_p[2]);
      } else {
        plane2(16, d.nd_x, T.w_fx1, T.l_fx1);
      }
      plane2(16, 128, T.a_big, T.l_big);
      plane2(8, 128, T.a_g, T.l_g);        // 32 KB; between two x heads it holds the auxiliary decoders' wgrad operands
      T.rec_buf = 8192 + (d.nd_c + d.nd_y) * 512;   // tile record: latent operand hi/lo planes + raw c|y per pair
      T.a_rec = b; b += 2 * T.rec_buf;
      auto f32 = [&](int floats) { int off = b; b += ((floats + 3) & ~3) * 4; return off; };
      T.f_inv = f32(16); T.f_bias_x = f32(64); T.f_bias_p1 = f32(32); T.f_bias_p2 = f32(64);
      T.f_aw0 = f32(2 * 64 * 4); T.f_ab0 = f32(2 * 64); T.f_aw1 = f32(2 * 64 * 4); T.f_ab1 = f32(8);
      T.f_w0f = f32(4 * 128); T.f_wp0f = f32(64 * 4);   // f_w0f: exchange scratch of the physics-latent gradients
      T.f_dza = f32(2 * nzd * 128); T.f_sc = f32(8 * 128); T.f_sca = f32(2 * 2 * 128); T.f_red = f32(256);
      T.o_bar = b; b += 96;
      T.total = (b + 127) & ~127;
      // R0 / R1 end-of-kernel scratch inside BIG; aux operands (64 x 128 mask plane + 24-column hi / lo product planes) inside G
      if (T.total <= 232448 && 256 * 44 * 4 <= 2 * T.l_big && 16384 + 2 * 6144 <= 2 * T.l_g && lat_smem_bytes(P, true) <= 160 * 1024) h->tc_ok = 1;
    }
  }
  return 0;
}

extern "C" {

int dpivae_abi_version(void) { return DPIVAE_ABI_VERSION; }
size_t dpivae_sizeof_model_desc(void) { return sizeof(dpivae_model_desc_t); }
const char* dpivae_last_error(void) { return g_err.c_str(); }

int dpivae_create(const dpivae_model_desc_t* desc, dpivae_handle_t* out) {
  if (!desc || !out) return fail("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: libdpivae_b200 has no CPU fallback");
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) return fail("libdpivae_b200 is built for sm_100a (B200) only");
  dpivae_model* h = new dpivae_model();
  h->d = *desc;
  h->sm_count = prop.multiProcessorCount;
  if (build_plan(h)) { delete h; return 1; }
  if (configure_dec_kernel() || configure_enc_kernels() || configure_dec_tc_kernel() || configure_lat_kernels() ||
      configure_enc_tc_kernels() || configure_enc_fused_kernels()) { delete h; return fail("cudaFuncSetAttribute(max dynamic smem) failed"); }
  // owner map: which kernel's partials hold each parameter's gradient
  std::vector<unsigned char> owner((size_t)desc->n_params, 0);
  for (int u = 0; u < h->enc.n_units; ++u) {
    const EncUnit& U = h->enc.u[u];
    const unsigned char cls = u < h->n_enc_units ? 2 : 1;
    for (long long e = 0; e < (long long)U.K0 * U.H; ++e) owner[U.g_w0 + e] = cls;
    for (long long e = 0; e < U.H; ++e) owner[U.g_b0 + e] = cls;
    for (long long e = 0; e < (long long)U.H * U.O; ++e) owner[U.g_w1 + e] = cls;
    for (long long e = 0; e < U.O; ++e) owner[U.g_b1 + e] = cls;
  }
  CUDA_OK(cudaMalloc(&h->d_owner, owner.size()));
  CUDA_OK(cudaMemcpy(h->d_owner, owner.data(), owner.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&h->d_group, owner.size()));
  CUDA_OK(cudaMemset(h->d_group, 0, owner.size()));
  CUDA_OK(cudaMalloc(&h->d_clip, sizeof(float)));
  h->n_groups = 1;
  h->lr[0] = 1e-3f; h->wd[0] = 0.0f;
  *out = h;
  return 0;
}

int dpivae_destroy(dpivae_handle_t h) {
  if (!h) return 0;
  cudaFree(h->d_frozen); cudaFree(h->d_owner); cudaFree(h->d_group); cudaFree(h->d_clip);
  for (int i = 0; i < 14; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  delete h;
  return 0;
}

int dpivae_set_physics_mlp(dpivae_handle_t h, const float* w, const float* b, const float* in_mean, const float* in_std) {
  if (!h || !w || !b || !in_mean || !in_std) return fail("null argument");
  if (h->d.phys_kind != DPIVAE_PHYS_MLP) return fail("model has no MLP physics decoder");
  std::vector<float> flat;
  long long wo = 0, bo = 0;
  for (int l = 0; l < h->d.phys_n_layers; ++l) {
    const int K = h->d.phys_dims[l], N = h->d.phys_dims[l + 1];
    flat.insert(flat.end(), w + wo, w + wo + (long long)K * N);
    flat.insert(flat.end(), b + bo, b + bo + N);
    wo += (long long)K * N; bo += N;
  }
  if (h->d_frozen) cudaFree(h->d_frozen);
  CUDA_OK(cudaMalloc(&h->d_frozen, flat.size() * sizeof(float)));
  CUDA_OK(cudaMemcpy(h->d_frozen, flat.data(), flat.size() * sizeof(float), cudaMemcpyHostToDevice));
  for (int i = 0; i < h->d.phys_dims[0] && i < 8; ++i) {
    h->dec.phys_in_mean[i] = in_mean[i];
    h->dec.phys_in_std[i] = in_std[i];
  }
  return 0;
}

int dpivae_bind(dpivae_handle_t h, float* params, float* grads, float* exp_avg, float* exp_avg_sq) {
  if (!h || !params) return fail("null argument");
  h->params = params; h->grads = grads; h->m = exp_avg; h->v = exp_avg_sq;
  return 0;
}

int dpivae_set_groups(dpivae_handle_t h, int32_t n_groups, const int64_t* begin, const int64_t* end, const float* lr,
                      const float* weight_decay) {
  if (!h || n_groups < 1 || n_groups > 16) return fail("1..16 parameter groups supported");
  std::vector<unsigned char> grp((size_t)h->d.n_params, 255);
  for (int g = 0; g < n_groups; ++g) {
    if (begin[g] < 0 || end[g] > h->d.n_params || begin[g] > end[g]) return fail("group range out of bounds");
    for (int64_t i = begin[g]; i < end[g]; ++i) grp[i] = (unsigned char)g;
    h->lr[g] = lr[g];
    h->wd[g] = weight_decay[g];
  }
  for (size_t i = 0; i < grp.size(); ++i)
    if (grp[i] == 255) return fail("parameter groups do not cover the flat buffer");
  h->n_groups = n_groups;
  CUDA_OK(cudaMemcpy(h->d_group, grp.data(), grp.size(), cudaMemcpyHostToDevice));
  return 0;
}

struct WsLayout {
  size_t hid, headpre, gpre, rowloss, part, scal, rec, dzrec, epsbuf, rowkl, hidrec, gmax, total;
  int grid_enc, grid_dec, RB, n_chunks;
  long long n_rowblocks;
  int tc_RB, tc_grid;           // tensor-core decoder kernel: rows per 128-pair tile, CTAs
  long long tc_rowblocks;
};

static WsLayout ws_layout(dpivae_handle_t h, int64_t B, int32_t n_mc) {
  WsLayout L;
  size_t o = 0;
  auto take = [&](size_t floats) { size_t s = o; o = align_up(o + floats * sizeof(float), 256); return s; };
  L.hid = take((size_t)h->H_tot * B);
  L.headpre = take((size_t)h->O_tot * B);
  L.gpre = take((size_t)h->O_tot * B);
  L.rowloss = take((size_t)6 * B);
  L.tc_RB = 128 / (n_mc < 1 ? 1 : (n_mc > 128 ? 128 : n_mc));
  if (L.tc_RB < 1) L.tc_RB = 1;
  L.tc_rowblocks = (B + L.tc_RB - 1) / L.tc_RB;
  L.tc_grid = (int)(L.tc_rowblocks < h->sm_count ? L.tc_rowblocks : h->sm_count);
  if (L.tc_grid < 1) L.tc_grid = 1;
  const long long ntiles = (B + TILE - 1) / TILE;
  // encoder kernels: two resident CTAs per SM when their shared-memory plan allows it (latency hiding)
  const int enc_per_sm = enc_smem_bytes(h->enc, true) <= 110 * 1024 ? 2 : 1;
  L.grid_enc = (int)(ntiles < (long long)enc_per_sm * h->sm_count ? ntiles : (long long)enc_per_sm * h->sm_count);
  int RB = TILE / (n_mc < 1 ? 1 : n_mc);
  RB = RB < 1 ? 1 : (RB > RBMAX ? RBMAX : RB);
  L.RB = RB;
  L.n_chunks = (int)(((long long)RB * n_mc + TILE - 1) / TILE);
  L.n_rowblocks = (B + RB - 1) / RB;
  L.grid_dec = (int)(L.n_rowblocks < h->sm_count ? L.n_rowblocks : h->sm_count);
  if (L.grid_enc < 1) L.grid_enc = 1;
  if (L.grid_dec < 1) L.grid_dec = 1;

  L.part = take((size_t)(4 * h->sm_count) * h->part_stride);
  L.scal = take(16);
  L.gmax = take(4);
  L.rec = L.dzrec = L.epsbuf = L.rowkl = 0;
  if (h->tc_ok && n_mc >= 8 && n_mc <= 128) {
    const size_t nt = (size_t)L.tc_rowblocks;
    const int Z = h->d.nz_x + h->d.nz_c + h->d.nz_y;
    L.rec = take(nt * (size_t)h->tc.rec_buf / 4);
    L.dzrec = take(nt * (size_t)Z * 128);
    L.epsbuf = take(nt * (size_t)Z * 128 + 16);   // tile records, or per-block local-order buffers (16-byte aligned blocks)
    L.rowkl = take((size_t)B);
  }
  L.hidrec = 0;
  if (h->enc_tc_bwd_ok) L.hidrec = take((size_t)((B + 127) / 128) * (size_t)h->enc_tc.hid_stride / 4);
  L.total = o;
  return L;
}

size_t dpivae_workspace_bytes(dpivae_handle_t h, int64_t B, int32_t n_mc) {
  if (!h || B < 1 || n_mc < 1) return 0;
  return ws_layout(h, B, n_mc).total;
}

static int run_loss(dpivae_handle_t h, const dpivae_batch_t* bt, const dpivae_rng_t* rng, const dpivae_loss_weights_t* w,
                    int with_grad, int latent_only, int x_std, const dpivae_outputs_t* out, void* ws, size_t ws_bytes,
                    cudaStream_t st, const AdamParams* fuse_adam = nullptr) {
  if (!h || !bt || !rng) return fail("null argument");
  if (!h->params) return fail("dpivae_bind has not been called");
  if (bt->B < 1 || bt->n_mc < 1 || bt->B_global < bt->B) return fail("bad batch sizes");
  const long long row_stride = bt->row_stride > 1 ? bt->row_stride : 1;
  if (bt->row_offset < 0 || bt->row_offset + (bt->B - 1) * row_stride >= bt->B_global) return fail("row_offset / row_stride leave the global batch");
  if (!bt->x || (!latent_only && !bt->c)) return fail("x / c must be given");
  if (with_grad && (!bt->y || !h->grads)) return fail("training needs y and a bound gradient buffer");
  if (h->d.phys_kind == DPIVAE_PHYS_MLP && !h->d_frozen && !latent_only) return fail("dpivae_set_physics_mlp has not been called");
  const WsLayout L = ws_layout(h, bt->B, bt->n_mc);
  if (!ws || ws_bytes < L.total) return fail("workspace too small");
  if (rng->mode == 0) {
    const int need = h->d.model_type == DPIVAE_MODEL_P ? 3 : 1;
    for (int k = 0; k < need; ++k)
      if (!rng->eps[k]) return fail("rng mode 0 needs injected eps buffers");
    if (bt->cond && !rng->eps[3]) return fail("cond=True needs eps[3]");
  }
  char* base = (char*)ws;
  float* hid = (float*)(base + L.hid);
  float* headpre = (float*)(base + L.headpre);
  float* gpre = (float*)(base + L.gpre);
  float* part = (float*)(base + L.part);
  float* scal = (float*)(base + L.scal);
  int launches = 0;

  const bool extra_out = out && (out->xh_p || out->xh_d || out->ch || out->log_sigma_c || out->yh || out->log_sigma_y);
  // (full-covariance conditional priors, --full_cov_prior, run the fp32 decoder kernel: the latent kernels of the
  // tensor-core path implement the diagonal priors of the reference's default)
  const bool use_tc = h->math_mode != DPIVAE_MATH_FP32 && h->tc_ok && !latent_only && !bt->cond && !extra_out &&
                      bt->n_mc >= 8 && bt->n_mc <= 128 && !h->dec.prior_full;
  h->last_dec_tc = use_tc ? 1 : 0;
  const int grid_dec = use_tc ? L.tc_grid : L.grid_dec;

  EncParams E = h->enc;
  E.params = h->params;
  E.x = bt->x; E.c = bt->c; E.y = bt->y; E.idx = (const long long*)bt->idx;
  E.B = bt->B;
  E.hid = hid; E.headpre = headpre; E.gpre = gpre;
  E.with_hid = with_grad;
  E.x_is_standardised = x_std;
  E.part = part + (long long)grid_dec * h->part_stride;
  E.part_stride = h->part_stride;
  if (latent_only) E.n_units = h->n_enc_units;  // prior nets not needed for encode
  const size_t enc_smem = enc_smem_bytes(h->enc, true);
  for (int k = 0; k < 4; ++k) h->ev_used[k] = 0;
  h->ev_used[5] = h->ev_used[6] = 0;
  // (max |gpre| of this batch, lat_bwd -> enc_tc_bwd: zeroed by the latent forward kernel, which precedes both)
  // encoder forward: tensor-core kernel for the encoder units (+ the FFMA kernel for the two prior nets) in the
  // tensor-core math modes when no backward follows; the FFMA kernel for everything otherwise
  const bool enc_tc = h->math_mode != DPIVAE_MATH_FP32 && h->enc_tc_ok && (!with_grad || h->enc_tc_bwd_ok);
  // the two conditional-prior nets: dedicated streaming kernels (prior_kernels.cu) when their shape allows
  EncParams E2 = E;
  E2.n_units = h->enc.n_units - h->n_enc_units;
  for (int u = 0; u < E2.n_units; ++u) E2.u[u] = h->enc.u[h->n_enc_units + u];
  const bool prior_fast = prior_kernels_support(E2) && bt->B >= 2048 && !h->dec.prior_full;   // small batches: units in parallel CTA columns instead
  const long long nt128 = (bt->B + 127) / 128;
  const int grid_etc = (int)(nt128 < h->sm_count ? nt128 : h->sm_count);
  const bool fused_encode = latent_only && enc_tc && h->enc_fused_ok && !getenv("DPIVAE_NO_FUSED_ENCODE");
  if (fused_encode) {
    // encode-only: encoder MMAs and latent sampling in one warp-specialised kernel (enc_fused_kernel.cu)
    KTimer t(h, 0, st);
    EncFusedParams F = h->enc_fused;
    F.q.params = h->params; F.q.x = bt->x; F.q.idx = (const long long*)bt->idx; F.q.B = bt->B;
    F.q.x_is_standardised = x_std; F.q.terms = h->math_mode == DPIVAE_MATH_TC_FP16X3 ? 3 : 1;
    F.rng.ss = h->cur_ss; F.rng.mode = rng->mode; F.rng.seed = rng->seed;
    for (int k = 0; k < 4; ++k) { F.rng.eps[k] = rng->eps[k]; F.rng.offset[k] = rng->offset[k]; F.rng.grid_threads[k] = rng->grid_threads[k] ? rng->grid_threads[k] : 256; }
    F.Bg = bt->B_global; F.row_off = bt->row_offset; F.row_stride = row_stride; F.n_mc = bt->n_mc;
    F.phase = h->d_phase;
    F.dbg = (h->d_phase && getenv("DPIVAE_ENC_DBG")) ? atoi(getenv("DPIVAE_ENC_DBG")) : 0;
    F.trace = (h->d_phase && getenv("DPIVAE_ENCODE_TRACE")) ? h->d_phase + 32 : nullptr;   // the probe's buffer holds 32 + 6 * 402 counters
    F.zx = out ? out->zx : nullptr; F.zc = out ? out->zc : nullptr; F.zy = out ? out->zy : nullptr; F.dens = out ? out->dens_z : nullptr;
    // in-kernel Philox noise: generated ahead by noise_fill_kernel (one evaluation per four elements, torch's own
    // mapping) into the hidden-activation region of the workspace, which this path does not use
    const long long Zt = h->d.nz_x + h->d.nz_c + h->d.nz_y;
    if (rng->mode == 1 && row_stride == 1 && (long long)bt->n_mc * Zt <= (long long)h->H_tot && !getenv("DPIVAE_NO_NOISE_PREPASS")) {
      float* e0 = hid;
      for (int b = 0; b < 3; ++b) { F.eps_local[b] = e0; e0 += ((size_t)bt->n_mc * bt->B * F.nz[b] + 3) & ~(size_t)3; }   // 16-byte aligned blocks (the few floats of rounding spill into headpre, equally unused here)
    }
    h->last_launches = launch_enc_fused(F, grid_etc, st);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  {
    KTimer t(h, 0, st);
    if (enc_tc) {
      EncTcParams Q = h->enc_tc;
      Q.params = h->params; Q.x = bt->x; Q.idx = (const long long*)bt->idx; Q.B = bt->B;
      Q.headpre = headpre; Q.hidrec = with_grad ? (unsigned char*)(base + L.hidrec) : nullptr; Q.x_is_standardised = x_std;
      Q.terms = h->math_mode == DPIVAE_MATH_TC_FP16X3 ? 3 : 1;
      launch_enc_tc_fwd(Q, grid_etc, st);
      ++launches;
      if (!latent_only) {
        // independent of the encoder kernel (other units, other output rows): launched to run alongside it
        if (prior_fast) launch_prior_fwd(E2, h->sm_count, st, true);
        else launch_enc_fwd(E2, L.grid_enc, enc_smem, st, true);
        ++launches;
      }
    } else if (prior_fast && !latent_only) {
      EncParams E1 = E;
      E1.n_units = h->n_enc_units;
      launch_enc_fwd(E1, L.grid_enc, enc_smem, st);
      launch_prior_fwd(E2, h->sm_count, st);
      launches += 2;
    } else {
      launch_enc_fwd(E, L.grid_enc, enc_smem, st);
      ++launches;
    }
  }

  DecParams D = h->dec;
  D.params = h->params; D.frozen = h->d_frozen;
  D.x = bt->x; D.c = bt->c; D.y = bt->y; D.idx = (const long long*)bt->idx;
  D.B = bt->B; D.Bg = bt->B_global; D.row_off = bt->row_offset; D.row_stride = row_stride;
  D.n_mc = bt->n_mc; D.cond = bt->cond; D.with_grad = with_grad;
  D.RB = L.RB; D.n_chunks = L.n_chunks; D.n_rowblocks = L.n_rowblocks;
  D.latent_only = latent_only;
  D.rng.ss = h->cur_ss;
  D.rng.mode = rng->mode; D.rng.seed = rng->seed;
  for (int k = 0; k < 4; ++k) { D.rng.eps[k] = rng->eps[k]; D.rng.offset[k] = rng->offset[k]; D.rng.grid_threads[k] = rng->grid_threads[k] ? rng->grid_threads[k] : 256; }
  if (w) { D.beta_x = w->beta_x; D.alpha_x = w->alpha_x; D.alpha_c = w->alpha_c; D.alpha_y = w->alpha_y; }
  else { D.beta_x = D.alpha_x = D.alpha_c = D.alpha_y = 1.0f; }
  D.headpre = headpre; D.gpre = gpre;
  D.part = part; D.part_stride = h->part_stride;
  D.phase = h->d_phase;
  memset(&D.out, 0, sizeof(D.out));
  if (out) {
    D.out.row_loss = out->row_loss;
    D.out.xh_p = out->xh_p; D.out.xh_d = out->xh_d; D.out.ch = out->ch; D.out.lsc = out->log_sigma_c;
    D.out.yh = out->yh; D.out.lsy = out->log_sigma_y; D.out.zx = out->zx; D.out.zc = out->zc; D.out.zy = out->zy;
    D.out.dens = out->dens_z;
  }
  if (latent_only) {
    // encode-only: per-pair streaming kernel (no decoders, no priors)
    KTimer t(h, 1, st);
    launch_lat_encode(D, st);
  } else if (use_tc) {
    TcParams T = h->tc;
    D.RB = L.tc_RB; D.n_chunks = 1; D.n_rowblocks = L.tc_rowblocks;
    D.rec = (unsigned char*)(base + L.rec); D.rec_stride = h->tc.rec_buf;
    D.dzrec = (float*)(base + L.dzrec); D.epsbuf = (float*)(base + L.epsbuf); D.rowkl = (float*)(base + L.rowkl);
    D.gpre_max = (with_grad && enc_tc) ? (unsigned int*)(base + L.gmax) : nullptr;
    // compile-time shapes: thread-per-pair latent kernels, noise in the local (m, row, i) order of each noise tensor;
    // an unsharded Philox call fills it ahead with one Philox evaluation per four elements (torch's own mapping)
    static const bool lat_v1 = getenv("DPIVAE_LAT_V1") != nullptr, no_prepass = getenv("DPIVAE_NO_NOISE_PREPASS") != nullptr;
    {
      KTimer t(h, 5, st);
      if (!lat_v1 && lat_pair_supported(D)) {
        float* e0 = (float*)(base + L.epsbuf);
        for (int b = 0; b < D.n_blk; ++b) { D.eps_local[b] = e0; e0 += ((size_t)bt->n_mc * bt->B * D.blk_size[b] + 3) & ~(size_t)3; }
        // noise ahead of the latent kernel, one Philox evaluation per four elements: unsharded calls, and CYCLIC row shards
        // (rank k of S: rows k, k + S, ...) whenever S * nz divides torch's generator grid -- the four elements of an
        // evaluation are grid_threads apart in the flattened (m, row, i) tensor, i.e. grid_threads / nz rows apart, so they
        // then fall on the same rank.  Contiguous shards keep the per-element generator inside the forward.
        int prepass = 0;
        // (below 16 k pairs the per-element generator inside the forward is cheaper than one more launch: same stream)
        if (rng->mode == 1 && !no_prepass && (long long)bt->B * bt->n_mc >= 16384) {
          if (bt->B_global == bt->B && bt->row_offset == 0) prepass = 1;
          else if (row_stride > 1 && bt->B * row_stride == bt->B_global && bt->row_offset < row_stride &&
                   (unsigned long long)bt->n_mc * (unsigned long long)bt->B_global < (1ull << 31)) {
            prepass = 2;
            for (int b = 0; b < D.n_blk; ++b)
              if (D.rng.grid_threads[b] % (unsigned long long)(D.blk_size[b] * row_stride) != 0) prepass = 0;
          }
        }
        if (prepass) {
          launch_lat_noise_fill(D, prepass == 2, st);
          D.eps_ready = 1;
          ++launches;
        }
      }
      launch_lat_fwd(D, L.tc_rowblocks, st);
    }
    ++launches;
    T.d = D;
    T.terms = h->math_mode == DPIVAE_MATH_TC_FP16X3 ? 3 : 1;
    { KTimer t(h, 1, st); launch_dec_tc(T, grid_dec, st); }
    if (with_grad) {
      { KTimer t(h, 6, st); launch_lat_bwd(D, L.tc_rowblocks, st); }
      ++launches;
    }
  } else {
    KTimer t(h, 1, st);
    launch_dec(D, grid_dec, st);
  }
  ++launches;

  if (with_grad) {
    KTimer t(h, 2, st);
    if (enc_tc) {
      EncTcParams Q = h->enc_tc;
      Q.params = h->params; Q.x = bt->x; Q.idx = (const long long*)bt->idx; Q.B = bt->B;
      Q.x_is_standardised = x_std; Q.hidrec = (unsigned char*)(base + L.hidrec); Q.gpre = gpre;
      Q.part = part + (long long)(grid_dec + L.grid_enc) * h->part_stride; Q.part_stride = h->part_stride;
      // gpre ~ O(1) / (B_global * D): bring it to O(16) before the fp16 split
      Q.e_g = (int)lrint(log2((double)bt->B_global * (double)(h->d.nd_x + h->d.nd_c + h->d.nd_y))) + 4;
      Q.gpre_max = use_tc ? (const unsigned int*)(base + L.gmax) : nullptr;   // measured by lat_bwd_kernel (tensor-core decoder path)
      launch_enc_tc_bwd(Q, grid_etc, st);
      ++launches;
      // runs alongside the encoder backward kernel (released once that kernel has seen lat_bwd complete)
      if (prior_fast) launch_prior_bwd(E2, L.grid_enc, st, true);
      else launch_enc_bwd(E2, L.grid_enc, enc_smem, st, true);
      ++launches;
    } else if (prior_fast) {
      EncParams E1 = E;
      E1.n_units = h->n_enc_units;
      launch_enc_bwd(E1, L.grid_enc, enc_smem, st);
      launch_prior_bwd(E2, L.grid_enc, st);
      launches += 2;
    } else {
      launch_enc_bwd(E, L.grid_enc, enc_smem, st);
      ++launches;
    }
  }
  if (!latent_only) {
    ReduceParams R;
    R.part = part; R.part_stride = h->part_stride;
    R.n_cta[0] = grid_dec; R.base[0] = 0;
    R.n_cta[1] = L.grid_enc; R.base[1] = grid_dec;
    if (with_grad && enc_tc) { R.n_cta[2] = grid_etc; R.base[2] = grid_dec + L.grid_enc; }
    else { R.n_cta[2] = L.grid_enc; R.base[2] = grid_dec; }
    R.n_params = h->d.n_params;
    R.owner = h->d_owner;
    R.grads = with_grad ? h->grads : nullptr;
    R.scalars = (out && out->scalars) ? out->scalars : scal;
    R.inv_B = 1.0f / (float)bt->B_global;
    R.inv_BD = 1.0f / ((float)bt->B_global * (float)(h->d.nd_x + h->d.nd_c + h->d.nd_y));
    R.fuse_adam = 0;
    if (fuse_adam && with_grad) { R.fuse_adam = 1; R.adam = *fuse_adam; }
    { KTimer t(h, 3, st); launch_reduce(R, st); }
    ++launches;
  }
  h->last_launches = launches;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int dpivae_loss(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng, const dpivae_loss_weights_t* w,
                int32_t with_grad, const dpivae_outputs_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  return run_loss(h, batch, rng, w, with_grad, 0, 0, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

static void fill_adam(dpivae_handle_t h, int64_t step, AdamParams& A);

int dpivae_adam_step(dpivae_handle_t h, int64_t step, float max_grad_norm, void* stream) {
  if (!h || !h->params || !h->grads || !h->m || !h->v) return fail("Adam needs bound params / grads / exp_avg / exp_avg_sq");
  if (step < 1) return fail("step is 1-based");
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  AdamParams A;
  fill_adam(h, step, A);
  if (max_grad_norm > 0.0f) {
    launch_gradnorm(h->grads, h->d.n_params, max_grad_norm, h->d_clip, st);
    A.clip_coef = h->d_clip;
    ++launches;
  }
  h->ev_used[4] = 0;
  { KTimer t(h, 4, st); launch_adam(A, st); }
  ++launches;
  h->last_launches = launches;
  CUDA_OK(cudaGetLastError());
  return 0;
}

static void fill_adam(dpivae_handle_t h, int64_t step, AdamParams& A) {
  memset(&A, 0, sizeof(A));
  A.params = h->params; A.grads = h->grads; A.m = h->m; A.v = h->v; A.group = h->d_group;
  const double b1 = 0.9, b2 = 0.999;
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  for (int g = 0; g < h->n_groups; ++g) { A.step_size[g] = (float)((double)h->lr[g] / bc1); A.wd[g] = h->wd[g]; }
  A.bc2_sqrt = (float)sqrt(bc2);
  A.beta1 = 0.9f; A.beta2 = 0.999f; A.eps = 1e-8f;
  A.n_params = h->d.n_params;
  A.clip_coef = nullptr;
  A.ss = h->cur_ss; A.log = h->cur_log; A.log_cap = h->cur_log_cap; A.lsx_index = h->d.log_sigma_x;
  A.scalars = h->grads + h->d.n_params;   // captured-step mode requires the [grads | 8 scalars] layout (checked there)
}

int dpivae_train_step(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng, const dpivae_loss_weights_t* w,
                      int64_t step, float max_grad_norm, const dpivae_outputs_t* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  // no gradient clipping: the Adam update rides on the gradient reduction (one launch, one pass over the gradients less)
  static const bool no_fuse = getenv("DPIVAE_NO_FUSED_ADAM") != nullptr;
  if (!(max_grad_norm > 0.0f) && !no_fuse && h && h->params && h->grads && h->m && h->v && step >= 1) {
    AdamParams A;
    fill_adam(h, step, A);
    h->ev_used[4] = 0;
    return run_loss(h, batch, rng, w, 1, 0, 0, out, workspace, workspace_bytes, (cudaStream_t)stream, &A);
  }
  if (run_loss(h, batch, rng, w, 1, 0, 0, out, workspace, workspace_bytes, (cudaStream_t)stream)) return 1;
  const int l0 = h->last_launches;
  if (dpivae_adam_step(h, step, max_grad_norm, stream)) return 1;
  h->last_launches += l0;
  return 0;
}

int dpivae_step_graph_create(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng, uint64_t philox_inc,
                             const dpivae_loss_weights_t* w, int64_t first_step, float max_grad_norm,
                             const int64_t* idx_pool, int64_t pool_rows, int64_t* idx_cur, float* scalars, float* step_log,
                             int64_t log_cap, int32_t unroll, void* workspace, size_t workspace_bytes, void* stream,
                             dpivae_step_graph_t* out) {
  if (!h || !batch || !rng || !out || !scalars) return fail("null argument");
  if (unroll < 1 || unroll > 64) return fail("unroll must be in 1..64");
  if (rng->mode != 1) return fail("a captured step draws its noise in-kernel (rng mode 1)");
  if (first_step < 1) return fail("first_step is 1-based");
  if (idx_pool && (!idx_cur || pool_rows < 1)) return fail("idx_pool needs idx_cur and pool_rows >= 1");
  if (step_log && log_cap < 1) return fail("step_log needs log_cap >= 1");
  if (scalars != h->grads + h->d.n_params) return fail("captured steps need the scalars right behind the gradient buffer");
  cudaStream_t st = (cudaStream_t)stream;
  dpivae_step_graph* g = new dpivae_step_graph();
  g->h = h;
  g->philox_inc = philox_inc;
  if (cudaMalloc(&g->d_state, sizeof(StepState)) != cudaSuccess) { delete g; return fail("cudaMalloc(step state) failed"); }
  *out = g;
  if (dpivae_step_graph_reset(g, rng, first_step, stream)) { dpivae_step_graph_destroy(g); *out = nullptr; return 1; }
  CUDA_OK(cudaStreamSynchronize(st));
  const int timing = h->timing;
  h->timing = 0;
  h->cur_ss = g->d_state; h->cur_log = step_log; h->cur_log_cap = log_cap;
  int rc = 0;
  auto capture = [&](int reps, cudaGraph_t* graph_out) -> int {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail("cudaStreamBeginCapture failed");
    int r = 0;
    for (int rep = 0; rep < reps && !r; ++rep) {
      AdvanceParams A;
      memset(&A, 0, sizeof(A));
      A.ss = g->d_state; A.philox_inc = philox_inc; A.n_groups = h->n_groups;
      for (int k = 0; k < h->n_groups; ++k) A.lr[k] = h->lr[k];
      A.idx_pool = (const long long*)idx_pool; A.pool_rows = pool_rows; A.B = batch->B; A.idx_cur = (long long*)idx_cur;
      launch_advance(A, st);
      dpivae_batch_t b = *batch;
      if (idx_pool) b.idx = idx_cur;
      dpivae_outputs_t o;
      memset(&o, 0, sizeof(o));
      o.scalars = scalars;
      r = dpivae_train_step(h, &b, rng, w, first_step, max_grad_norm, &o, workspace, workspace_bytes, stream);
      g->launches_per_step = h->last_launches + 1;
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (!r && e != cudaSuccess) r = fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    *graph_out = graph;
    return r;
  };
  rc = capture(1, &g->graph);
  if (!rc && unroll > 1) rc = capture(unroll, &g->graph_u);
  g->unroll = unroll > 1 ? unroll : 1;
  h->cur_ss = nullptr; h->cur_log = nullptr; h->cur_log_cap = 0;
  h->timing = timing;
  if (!rc && cudaGraphInstantiate(&g->exec, g->graph, 0) != cudaSuccess) rc = fail("cudaGraphInstantiate failed");
  if (!rc && g->graph_u && cudaGraphInstantiate(&g->exec_u, g->graph_u, 0) != cudaSuccess) rc = fail("cudaGraphInstantiate failed");
  if (rc) { dpivae_step_graph_destroy(g); *out = nullptr; return 1; }
  return 0;
}

int dpivae_step_graph_reset(dpivae_step_graph_t g, const dpivae_rng_t* rng, int64_t next_step, void* stream) {
  if (!g || !rng || next_step < 1) return fail("bad argument");
  StepState s;
  memset(&s, 0, sizeof(s));
  s.step = next_step - 1;                                                  // advance_kernel increments before use
  for (int k = 0; k < 4; ++k) s.philox_off[k] = rng->offset[k] - g->philox_inc;
  // pageable host source: the copy is staged before the call returns
  CUDA_OK(cudaMemcpyAsync(g->d_state, &s, sizeof(s), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

int dpivae_step_graph_launch(dpivae_step_graph_t g, int32_t n_steps, void* stream) {
  if (!g || !g->exec || n_steps < 0) return fail("bad argument");
  int left = n_steps;
  if (g->exec_u)
    for (; left >= g->unroll; left -= g->unroll) CUDA_OK(cudaGraphLaunch(g->exec_u, (cudaStream_t)stream));
  for (; left > 0; --left) CUDA_OK(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  g->h->last_launches = n_steps * g->launches_per_step;
  return 0;
}

int dpivae_step_graph_destroy(dpivae_step_graph_t g) {
  if (!g) return 0;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  if (g->exec_u) cudaGraphExecDestroy(g->exec_u);
  if (g->graph_u) cudaGraphDestroy(g->graph_u);
  cudaFree(g->d_state);
  delete g;
  return 0;
}

int dpivae_encode(dpivae_handle_t h, const dpivae_batch_t* batch, const dpivae_rng_t* rng, int32_t x_is_standardised,
                  float* zx, float* zc, float* zy, float* dens_z, void* workspace, size_t workspace_bytes, void* stream) {
  dpivae_outputs_t o;
  memset(&o, 0, sizeof(o));
  o.zx = zx; o.zc = zc; o.zy = zy; o.dens_z = dens_z;
  return run_loss(h, batch, rng, nullptr, 0, 1, x_is_standardised, &o, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dpivae_decode(dpivae_handle_t h, const float* zx_in, const float* zc, const float* zy, int64_t B, int32_t n_mc,
                  const dpivae_outputs_t* out, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !zx_in || !zc || !zy || !out) return fail("null argument");
  if (!h->params) return fail("dpivae_bind has not been called");
  if (B < 1 || n_mc < 1) return fail("bad batch sizes");
  if (h->d.phys_kind == DPIVAE_PHYS_MLP && !h->d_frozen) return fail("dpivae_set_physics_mlp has not been called");
  const WsLayout L = ws_layout(h, B, n_mc);
  if (!ws || ws_bytes < L.total) return fail("workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)ws;
  // the fused FFMA decoder kernel in forward-only mode; its encoder-side inputs (head pre-activations, data rows)
  // are zero-filled workspace regions, its losses are discarded
  float* zeros = (float*)(base + L.hid);
  CUDA_OK(cudaMemsetAsync(base + L.hid, 0, (size_t)h->H_tot * B * sizeof(float), st));
  CUDA_OK(cudaMemsetAsync(base + L.headpre, 0, (size_t)h->O_tot * B * sizeof(float), st));
  DecParams D = h->dec;
  D.params = h->params; D.frozen = h->d_frozen;
  D.x = zeros; D.c = zeros; D.y = zeros; D.idx = nullptr;
  D.B = B; D.Bg = B; D.row_off = 0; D.row_stride = 1;
  D.n_mc = n_mc; D.cond = 0; D.with_grad = 0;
  D.RB = L.RB; D.n_chunks = L.n_chunks; D.n_rowblocks = L.n_rowblocks;
  D.latent_only = 0;
  D.rng.ss = nullptr; D.rng.mode = 1; D.rng.seed = 0;   // the sampled latents are overwritten by the caller's: any noise will do
  for (int k = 0; k < 4; ++k) { D.rng.eps[k] = nullptr; D.rng.offset[k] = 0; D.rng.grid_threads[k] = 256; }
  D.beta_x = D.alpha_x = D.alpha_c = D.alpha_y = 1.0f;
  D.headpre = (float*)(base + L.headpre); D.gpre = (float*)(base + L.gpre);
  D.part = (float*)(base + L.part); D.part_stride = h->part_stride;
  D.phase = nullptr;
  memset(&D.out, 0, sizeof(D.out));
  D.out.xh_p = out->xh_p; D.out.xh_d = out->xh_d; D.out.ch = out->ch; D.out.lsc = out->log_sigma_c;
  D.out.yh = out->yh; D.out.lsy = out->log_sigma_y;
  D.zin_x = zx_in; D.zin_c = zc; D.zin_y = zy;
  launch_dec(D, L.grid_dec, st);
  h->last_launches = 1;
  h->last_dec_tc = 0;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int dpivae_prior_net(dpivae_handle_t h, const float* c, const float* y, int64_t B, float* loc_c, float* scale_tril_c,
                     float* loc_y, float* scale_tril_y, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !c || !loc_c || !scale_tril_c) return fail("null argument");
  if (y && (!loc_y || !scale_tril_y)) return fail("y given without output buffers");
  if (!h->params) return fail("dpivae_bind has not been called");
  if (B < 1) return fail("bad batch size");
  const WsLayout L = ws_layout(h, B, 1);
  if (!ws || ws_bytes < L.total) return fail("workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)ws;
  EncParams E = h->enc;
  E.params = h->params;
  E.x = nullptr; E.c = c; E.y = y; E.idx = nullptr;
  E.B = B;
  E.hid = (float*)(base + L.hid); E.headpre = (float*)(base + L.headpre); E.gpre = nullptr;
  E.with_hid = 0; E.x_is_standardised = 0;
  E.part = nullptr; E.part_stride = h->part_stride;
  EncParams E2 = E;
  E2.n_units = E.n_units - h->n_enc_units;
  for (int u = 0; u < E2.n_units; ++u) E2.u[u] = E.u[h->n_enc_units + u];
  if (prior_kernels_support(E2)) launch_prior_fwd(E2, h->sm_count, st);
  else launch_enc_fwd(E2, L.grid_enc, enc_smem_bytes(h->enc, true), st);
  launch_prior_post(E.headpre, B, h->dec.hpri[0], h->d.nz_c, h->dec.prior_full, loc_c, scale_tril_c, st);
  int launches = 2;
  if (y) { launch_prior_post(E.headpre, B, h->dec.hpri[1], h->d.nz_y, h->dec.prior_full, loc_y, scale_tril_y, st); ++launches; }
  h->last_launches = launches;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int dpivae_gaussian_sample(const float* loc, const float* scale_tril, const float* eps, int32_t n_mc, int64_t B, int32_t nz,
                           float* z, float* dens, void* stream) {
  if (!loc || !scale_tril || !eps || !z || !dens) return fail("null argument");
  if (n_mc < 1 || B < 1 || nz < 1 || nz > DPIVAE_MAX_Z) return fail("bad sizes");
  launch_gaussian_sample(loc, scale_tril, eps, n_mc, B, nz, z, dens, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int dpivae_mc_mean(const float* v, int32_t n_mc, int64_t B, int32_t d, float* out, void* stream) {
  if (!v || !out || n_mc < 1 || B < 1 || d < 1) return fail("bad argument");
  launch_mc_mean(v, n_mc, B * d, out, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int dpivae_regression_metrics(const float* y_true, const float* y_pred, int64_t N, int32_t d, double* scratch, float* out3,
                              void* stream) {
  if (!y_true || !y_pred || !scratch || !out3 || N < 1 || d < 1 || d > 16) return fail("bad argument");
  launch_regression_metrics(y_true, y_pred, N, d, scratch, out3, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int dpivae_linreg_r2(const float* X_train, const float* y_train, int64_t ldy_train, int64_t N_train, const float* X_test,
                     const float* y_test, int64_t ldy_test, int64_t N_test, int32_t k, double* scratch, float* r2_out,
                     void* stream) {
  if (!X_train || !y_train || !X_test || !y_test || !scratch || !r2_out) return fail("null argument");
  if (k < 1 || k > 8 || N_train < 1 || N_test < 1 || ldy_train < 1 || ldy_test < 1) return fail("bad sizes (1 <= k <= 8)");
  launch_linreg_r2(X_train, y_train, ldy_train, N_train, X_test, y_test, ldy_test, N_test, k, scratch, r2_out, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int datagen_maxw(const dpivae_datagen_desc_t* d) {
  int m = 0;
  for (int l = 0; l <= d->n_layers; ++l) m = d->dims[l] > m ? d->dims[l] : m;
  return m;
}

size_t dpivae_datagen_workspace_bytes(const dpivae_datagen_desc_t* d, int64_t n) {
  if (!d || n < 1) return 0;
  return 2 * align_up((size_t)n * (size_t)datagen_maxw(d) * sizeof(float), 256);
}

int dpivae_sample_response(const dpivae_datagen_desc_t* d, const float* w, const float* b, int64_t n, uint64_t seed,
                           uint64_t offset_in, int32_t sm_count, int32_t max_threads_per_sm, float* z, float* x, float* c,
                           float* y, void* ws, size_t ws_bytes, void* stream, uint64_t* offset_out) {
  if (!d || !w || !b || !z || !x || !c || !y || !ws) return fail("null argument");
  if (n < 1 || d->n_factors < 1 || d->n_factors > 16 || d->n_layers < 1 || d->n_layers > DPIVAE_MAX_PHYS_LAYERS + 1)
    return fail("bad generator sizes");
  if (d->dims[0] != d->n_factors || d->dims[d->n_layers] != d->nd_x) return fail("surrogate dims must run from n_factors to nd_x");
  if (d->nd_c < 1 || d->nd_c > DPIVAE_MAX_NDCY || d->nd_y < 1 || d->nd_y > DPIVAE_MAX_NDCY) return fail("bad nd_c / nd_y");
  if (ws_bytes < dpivae_datagen_workspace_bytes(d, n)) return fail("workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  DataGenParams P;
  memset(&P, 0, sizeof(P));
  P.n = n; P.nf = d->n_factors; P.nd_x = d->nd_x; P.nd_c = d->nd_c; P.nd_y = d->nd_y;
  for (int j = 0; j < d->n_factors; ++j) { P.lo[j] = d->lo[j]; P.hi[j] = d->hi[j]; P.in_mean[j] = d->in_mean[j]; P.in_std[j] = d->in_std[j]; }
  for (int j = 0; j < DPIVAE_MAX_NDCY; ++j) { P.idx_c[j] = d->idx_c[j]; P.idx_y[j] = d->idx_y[j]; }
  for (int j = 0; j < d->nd_c; ++j) if (d->idx_c[j] < 0 || d->idx_c[j] >= d->n_factors) return fail("idx_c out of range");
  for (int j = 0; j < d->nd_y; ++j) if (d->idx_y[j] < 0 || d->idx_y[j] >= d->n_factors) return fail("idx_y out of range");
  P.sigma_x = d->sigma_x; P.sigma_c = d->sigma_c; P.sigma_y = d->sigma_y;
  P.seed = seed;
  // generator bookkeeping of torch's distribution kernels (calc_execution_policy: 256 threads, unroll 4)
  uint64_t cur = offset_in;
  const uint64_t cap = (uint64_t)sm_count * ((uint64_t)max_threads_per_sm / 256);
  auto plan = [&](uint64_t numel, unsigned long long& off, unsigned int& T) {
    uint64_t grid = (numel + 255) / 256;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    off = cur; T = (unsigned int)(256 * grid);
    cur += ((numel - 1) / (256 * grid * 4) + 1) * 4;
  };
  for (int j = 0; j < d->n_factors; ++j) plan((uint64_t)n, P.off_u[j], P.T_u);
  plan((uint64_t)n * d->nd_x, P.off_x, P.T_x);
  plan((uint64_t)n * d->nd_c, P.off_c, P.T_c);
  plan((uint64_t)n * d->nd_y, P.off_y, P.T_y);
  if (offset_out) *offset_out = cur;
  const size_t half = align_up((size_t)n * (size_t)datagen_maxw(d) * sizeof(float), 256);
  float* bufA = (float*)ws;
  float* bufB = (float*)((char*)ws + half);
  P.z = z; P.a0 = bufA; P.x = x; P.c = c; P.y = y;
  launch_datagen_latents(P, st);
  const float* in = bufA;
  long long wo = 0, bo = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    const int K = d->dims[l], N = d->dims[l + 1];
    const bool last = l == d->n_layers - 1;
    float* out = last ? x : (in == bufA ? bufB : bufA);
    launch_mlp_layer(in, w + wo, b + bo, out, n, K, N, !last, st);
    wo += (long long)K * N; bo += N;
    in = out;
  }
  launch_datagen_finish(P, st);
  CUDA_OK(cudaGetLastError());
  return 0;
}

uint64_t dpivae_philox_plan(dpivae_handle_t h, int64_t B_global, int32_t n_mc, int32_t cond, uint64_t offset_in,
                            int32_t sm_count, int32_t max_threads_per_sm, dpivae_rng_t* rng) {
  if (!h || !rng) return offset_in;
  int nz[4] = {0, 0, 0, 0};
  int nt = 0;
  if (h->d.model_type == DPIVAE_MODEL_P) { nz[0] = h->d.nz_x; nz[1] = h->d.nz_c; nz[2] = h->d.nz_y; nt = 3; }
  else { nz[0] = h->d.nz_x + h->d.nz_c + h->d.nz_y; nt = 1; }
  uint64_t cur = offset_in;
  const uint64_t blocks_per_sm = (uint64_t)max_threads_per_sm / 256;
  auto plan_one = [&](int slot, int width) {
    const uint64_t numel = (uint64_t)n_mc * (uint64_t)B_global * (uint64_t)width;
    uint64_t grid = (numel + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * blocks_per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    rng->offset[slot] = cur;
    rng->grid_threads[slot] = (uint32_t)(256 * grid);
    cur += ((numel - 1) / (256 * grid * 4) + 1) * 4;
  };
  for (int k = 0; k < nt; ++k) plan_one(k, nz[k]);
  if (cond) plan_one(3, h->d.nz_c);
  return cur;
}

int dpivae_last_launch_count(dpivae_handle_t h) { return h ? h->last_launches : 0; }

int dpivae_set_math_mode(dpivae_handle_t h, int32_t mode) {
  if (!h) return fail("null handle");
  if (mode != DPIVAE_MATH_FP32 && mode != DPIVAE_MATH_TC_FP16X3 && mode != DPIVAE_MATH_TC_FP16)
    return fail("unknown math mode");
  h->math_mode = mode;
  return 0;
}

int dpivae_last_used_tensor_cores(dpivae_handle_t h) { return h ? h->last_dec_tc : 0; }

int dpivae_set_phase_buffer(dpivae_handle_t h, void* dev_counters16) {
  if (!h) return fail("null handle");
  h->d_phase = (long long*)dev_counters16;
  return 0;
}

int dpivae_set_timing(dpivae_handle_t h, int32_t enable) {
  if (!h) return fail("null handle");
  if (enable && !h->ev[0])
    for (int i = 0; i < 14; ++i) CUDA_OK(cudaEventCreate(&h->ev[i]));
  h->timing = enable ? 1 : 0;
  return 0;
}

int dpivae_last_kernel_ms(dpivae_handle_t h, float* out5) {
  if (!h || !out5) return fail("null argument");
  for (int k = 0; k < 7; ++k) {
    out5[k] = 0.0f;
    if (h->ev[0] && h->ev_used[k]) {
      CUDA_OK(cudaEventSynchronize(h->ev[2 * k + 1]));
      CUDA_OK(cudaEventElapsedTime(&out5[k], h->ev[2 * k], h->ev[2 * k + 1]));
    }
  }
  return 0;
}

int dpivae_ffma_peak_tflops(float* tflops_out, void* stream) {
  if (!tflops_out) return fail("null argument");
  *tflops_out = ffma_peak_tflops((cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
