// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the tensor-core decoder kernel.
//
// Operand storage ("X8" layout, 16-bit elements): a matrix with R rows and C columns (C % 8 == 0) is
// kept in shared memory as uint4 X8[C/8][R], X8[c/8][r] = the 8 halves X[r][c .. c+7].  With NO
// swizzle this single array is a valid tcgen05 canonical layout in BOTH orientations:
//   K-major  operand (MN index = r, K index = c): core matrix = 8 rows x 16 B contiguous (128 B),
//            SBO (next 8 rows) = 128 B, LBO (next 16 B of K) = R * 16 B.
//   MN-major operand (MN index = c, K index = r): core matrix = 8 K-rows x 16 B (8 MN elements),
//            SBO (next 8 MN elements) = R * 16 B, LBO (next 8 K rows) = 128 B.
// So forward (activations x weights^T), dgrad (gradients x weights) and wgrad (activations^T x
// gradients) all read the same buffers; only the descriptor orientation changes.  (32-bit tf32
// operands cannot do this: MN-major tf32 exists only in the SWIZZLE_128B_BASE32B layout -- measured,
// tools/microbench/tc_layout_probe.cu -- which no K-major layout matches.)
//
// FP32 accuracy on the fp16 tensor pipe: every fp32 operand x is split x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi) (22 significant bits) and each GEMM is issued as the three
// products hi*hi + lo*hi + hi*lo into the same fp32 TMEM accumulator.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpv {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (no swizzle, descriptor version 1) ----------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// X8 buffer with R rows, used K-major starting at column chunk `chunk0` (8 columns per chunk)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int R, int chunk0) {
  return make_desc(base + (uint32_t)(chunk0 * R) * 16u, (uint32_t)R * 16u, 128u);
}
// X8 buffer with R rows, used MN-major: MN starts at column chunk `chunk0`, K starts at row `row0`
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int R, int chunk0, int row0) {
  return make_desc(base + (uint32_t)(chunk0 * R + row0) * 16u, 128u, (uint32_t)R * 16u);
}

// instruction descriptor: kind::f16 with fp16 operands, fp32 accumulate, M x N tile, operand majors (0 = K, 1 = MN)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one K = 16 step (32 bytes of fp16 along K)
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// same with the A operand in tensor memory (128 lanes x K/2 packed 32-bit columns: column j = halves (2j, 2j+1))
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// plain arrive (release at CTA scope): pairs with the acquire of mbar_wait on the waiting side
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// transaction-count arrive + 1-D bulk async copy global -> shared (TMA engine, completes on `bar`)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread t of the warp gets row (lane quadrant base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  tmem_wait_st();
}

// asynchronous forms: the registers are valid only after tmem_wait_ld() / the stores are complete after tmem_wait_st()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  tmem_wait_ld();
}
// 8 fp32 values -> packed fp16 hi / lo planes in tensor memory (A operand of a TS-mode MMA): 4 columns each
__device__ __forceinline__ void tmem_put8_packed(uint32_t t_hi, uint32_t t_lo, const float* v);
// and back: value = (hi + lo) * inv_scale
__device__ __forceinline__ void tmem_get8_packed(uint32_t t_hi, uint32_t t_lo, float inv_scale, float* v) {
  uint32_t h[4], l[4];
  tmem_ld4(t_hi, h);
  tmem_ld4(t_lo, l);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
    v[2 * i] = (hf.x + lf.x) * inv_scale;
    v[2 * i + 1] = (hf.y + lf.y) * inv_scale;
  }
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
  tmem_wait_st();
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
  tmem_wait_st();
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  tmem_wait_st();
}

// ---- fp32 -> (hi, lo) fp16 split of 8 consecutive values, packed as two uint4 ------------------
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void tmem_put8_packed(uint32_t t_hi, uint32_t t_lo, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  const float ph[4] = {__uint_as_float(hi.x), __uint_as_float(hi.y), __uint_as_float(hi.z), __uint_as_float(hi.w)};
  const float pl[4] = {__uint_as_float(lo.x), __uint_as_float(lo.y), __uint_as_float(lo.z), __uint_as_float(lo.w)};
  tmem_st4(t_hi, ph);
  tmem_st4(t_lo, pl);
}

// A split operand: hi plane at `base`, lo plane at `base + lo_off` (bytes), both X8 with R rows.
struct Op {
  uint32_t base, lo_off;
  int R;
};

// GEMM issue helpers (ONE thread); k extents in elements, multiples of 16.  `terms` = 3 issues
// hi*hi + lo*hi + hi*lo (fp32-accurate), `terms` = 1 only hi*hi (plain fp16 inputs).  Descriptors are kept
// as (lo, hi) 32-bit words: stepping along K only adds a constant to the 14-bit start-address field.
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
constexpr uint32_t DESC_VERSION_HI = 1u << 14;  // descriptor bit 46

//   fwd   : D[128 rows of A][N]   (+)= sum_k A[r][k] * W[n][k]      A K-major, W K-major (N rows)
static __device__ __noinline__ void issue_fwd(uint32_t d_tmem, Op a, Op w, int N, int K, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  const uint32_t ahi = 8u | DESC_VERSION_HI, whi = 8u | DESC_VERSION_HI;       // SBO = 128 B
  const uint32_t astep = 2u * (uint32_t)a.R, wstep = 2u * (uint32_t)w.R;       // two 16-byte K chunks per MMA
  const int ks = K >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t alo = (((a.base + (t == 1 ? a.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)a.R << 16);  // LBO = R * 16 B
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)w.R << 16);
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
      mma_f16(d_tmem, pack64(alo, ahi), pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      alo += astep;
      wlo += wstep;
    }
  }
}
//   dgrad : D[128 rows of G][Kin] (+)= sum_n G[r][n] * W[n][kin]    G K-major, W (Nout rows) MN-major over kin
static __device__ __noinline__ void issue_dgrad(uint32_t d_tmem, Op g, Op w, int Nout, int Kin, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, Kin, 0, 1);
  const uint32_t ghi = 8u | DESC_VERSION_HI, whi = (uint32_t)w.R | DESC_VERSION_HI;  // MN-major: SBO = R * 16 B
  const uint32_t gstep = 2u * (uint32_t)g.R, wstep = 16u;                            // 16 K rows = 256 B
  const int ks = Nout >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t glo = (((g.base + (t == 1 ? g.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)g.R << 16);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);   // LBO = 128 B
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
      mma_f16(d_tmem, pack64(glo, ghi), pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      glo += gstep;
      wlo += wstep;
    }
  }
}
//   TS forms: A (128 x K) in tensor memory as packed fp16 pairs, hi plane at a_tmem, lo plane at a_tmem + K / 2
static __device__ __noinline__ void issue_fwd_ts(uint32_t d_tmem, uint32_t a_tmem, Op w, int N, int K, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  const uint32_t whi = 8u | DESC_VERSION_HI, wstep = 2u * (uint32_t)w.R;
  const int ks = K >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t a = a_tmem + (t == 1 ? (uint32_t)(K >> 1) : 0u);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)w.R << 16);
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
      mma_f16_ts(d_tmem, a, pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      a += 8u;
      wlo += wstep;
    }
  }
}
static __device__ __noinline__ void issue_dgrad_ts(uint32_t d_tmem, uint32_t g_tmem, Op w, int Nout, int Kin, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, Kin, 0, 1);
  const uint32_t whi = (uint32_t)w.R | DESC_VERSION_HI;
  const int ks = Nout >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t a = g_tmem + (t == 1 ? (uint32_t)(Nout >> 1) : 0u);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
      mma_f16_ts(d_tmem, a, pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      a += 8u;
      wlo += 16u;
    }
  }
}
//   wgrad : D[128 cols of H][N cols of G] (+)= sum_r H[r][m] * G[r][n]   both MN-major, r over R rows
static __device__ __noinline__ void issue_wgrad(uint32_t d_tmem, Op h, Op g, int N, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, N, 1, 1);
  const uint32_t hhi = (uint32_t)h.R | DESC_VERSION_HI, ghi = (uint32_t)g.R | DESC_VERSION_HI;
  const int ks = h.R >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t hlo = (((h.base + (t == 1 ? h.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
    uint32_t glo = (((g.base + (t == 2 ? g.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
      mma_f16(d_tmem, pack64(hlo, hhi), pack64(glo, ghi), idesc, accumulate);
      accumulate = 1;
      hlo += 16u;
      glo += 16u;
    }
  }
}


// ---- warp-wide issue (uniform datapath) ------------------------------------------------------------------------------
// The helpers above are called by ONE thread inside a divergent branch; there ptxas cannot prove the descriptors
// warp-uniform and wraps every tcgen05.mma in an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY "waterfall" loop (~110 cycles
// per MMA measured in the decoder kernel, above the 64-cycle tensor-pipe floor).  The *_w variants are executed by ALL
// 32 converged lanes of the issue warp with operands derived only from kernel parameters, constants and uniform loop
// counters: the descriptor arithmetic runs on the uniform datapath and only the MMA / commit themselves are predicated
// on the lane elected once per stage (`el` != 0 in exactly one lane).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t el;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(el));
  return el;
}
__device__ __forceinline__ void mma_f16_w(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts_w(uint32_t el, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el)
      : "memory");
}
__device__ __forceinline__ void commit_w(uint32_t el, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(el)
      : "memory");
}
__device__ __forceinline__ void issue_fwd_w(uint32_t el, uint32_t d_tmem, Op a, Op w, int N, int K, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  const uint32_t ahi = 8u | DESC_VERSION_HI, whi = 8u | DESC_VERSION_HI;
  const uint32_t astep = 2u * (uint32_t)a.R, wstep = 2u * (uint32_t)w.R;
  const int ks = K >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t alo = (((a.base + (t == 1 ? a.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)a.R << 16);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)w.R << 16);
    for (int k = 0; k < ks; ++k) {
      mma_f16_w(el, d_tmem, pack64(alo, ahi), pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      alo += astep;
      wlo += wstep;
    }
  }
}
// compile-time shapes: both loops unroll completely, every descriptor is base + immediate (the run-time loops above cost
// ~100 cycles of issue per MMA in the encoder kernels -- more than the MMAs themselves)
template <int N, int K, int TERMS, int AR>
__device__ __forceinline__ void issue_fwd_s(uint32_t el, uint32_t d_tmem, Op a, Op w, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
  constexpr uint32_t hi = 8u | DESC_VERSION_HI;
  const uint32_t a0 = ((a.base >> 4) & 0x3FFFu) | ((uint32_t)AR << 16), a1 = (((a.base + a.lo_off) >> 4) & 0x3FFFu) | ((uint32_t)AR << 16);
  const uint32_t w0 = ((w.base >> 4) & 0x3FFFu) | ((uint32_t)N << 16), w1 = (((w.base + w.lo_off) >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
#pragma unroll
    for (int k = 0; k < (K >> 4); ++k) {
      mma_f16_w(el, d_tmem, pack64((t == 1 ? a1 : a0) + (uint32_t)(k * 2 * AR), hi), pack64((t == 2 ? w1 : w0) + (uint32_t)(k * 2 * N), hi), idesc,
                (t == 0 && k == 0) ? accumulate : 1u);
    }
  }
}
template <int N, int K, int TERMS>
__device__ __forceinline__ void issue_fwd_ts_s(uint32_t el, uint32_t d_tmem, uint32_t a_tmem, Op w, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
  constexpr uint32_t hi = 8u | DESC_VERSION_HI;
  const uint32_t w0 = ((w.base >> 4) & 0x3FFFu) | ((uint32_t)N << 16), w1 = (((w.base + w.lo_off) >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
#pragma unroll
    for (int k = 0; k < (K >> 4); ++k) {
      mma_f16_ts_w(el, d_tmem, a_tmem + (uint32_t)((t == 1 ? (K >> 1) : 0) + 8 * k), pack64((t == 2 ? w1 : w0) + (uint32_t)(k * 2 * N), hi), idesc,
                   (t == 0 && k == 0) ? accumulate : 1u);
    }
  }
}
__device__ __forceinline__ void issue_dgrad_w(uint32_t el, uint32_t d_tmem, Op g, Op w, int Nout, int Kin, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, Kin, 0, 1);
  const uint32_t ghi = 8u | DESC_VERSION_HI, whi = (uint32_t)w.R | DESC_VERSION_HI;
  const uint32_t gstep = 2u * (uint32_t)g.R, wstep = 16u;
  const int ks = Nout >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t glo = (((g.base + (t == 1 ? g.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)g.R << 16);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
    for (int k = 0; k < ks; ++k) {
      mma_f16_w(el, d_tmem, pack64(glo, ghi), pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      glo += gstep;
      wlo += wstep;
    }
  }
}
__device__ __forceinline__ void issue_fwd_ts_w(uint32_t el, uint32_t d_tmem, uint32_t a_tmem, Op w, int N, int K, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  const uint32_t whi = 8u | DESC_VERSION_HI, wstep = 2u * (uint32_t)w.R;
  const int ks = K >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t a = a_tmem + (t == 1 ? (uint32_t)(K >> 1) : 0u);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | ((uint32_t)w.R << 16);
    for (int k = 0; k < ks; ++k) {
      mma_f16_ts_w(el, d_tmem, a, pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      a += 8u;
      wlo += wstep;
    }
  }
}
__device__ __forceinline__ void issue_dgrad_ts_w(uint32_t el, uint32_t d_tmem, uint32_t g_tmem, Op w, int Nout, int Kin, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, Kin, 0, 1);
  const uint32_t whi = (uint32_t)w.R | DESC_VERSION_HI;
  const int ks = Nout >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t a = g_tmem + (t == 1 ? (uint32_t)(Nout >> 1) : 0u);
    uint32_t wlo = (((w.base + (t == 2 ? w.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
    for (int k = 0; k < ks; ++k) {
      mma_f16_ts_w(el, d_tmem, a, pack64(wlo, whi), idesc, accumulate);
      accumulate = 1;
      a += 8u;
      wlo += 16u;
    }
  }
}
__device__ __forceinline__ void issue_wgrad_w(uint32_t el, uint32_t d_tmem, Op h, Op g, int N, uint32_t accumulate, int terms) {
  const uint32_t idesc = make_idesc(128, N, 1, 1);
  const uint32_t hhi = (uint32_t)h.R | DESC_VERSION_HI, ghi = (uint32_t)g.R | DESC_VERSION_HI;
  const int ks = h.R >> 4;
  for (int t = 0; t < terms; ++t) {
    uint32_t hlo = (((h.base + (t == 1 ? h.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
    uint32_t glo = (((g.base + (t == 2 ? g.lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
    for (int k = 0; k < ks; ++k) {
      mma_f16_w(el, d_tmem, pack64(hlo, hhi), pack64(glo, ghi), idesc, accumulate);
      accumulate = 1;
      hlo += 16u;
      glo += 16u;
    }
  }
}

// Masked weight gradient, M = 64: D[64 units][N] (+)= sum over R pair rows of A[r][unit] * (B_hi[r][n] + B_lo[r][n]).
// A is an EXACT fp16 plane (0 / 1 ReLU masks: no lo plane, two terms instead of three), both operands MN-major X8
// buffers with R rows.  An M = 64 accumulator row u sits in tensor-memory lane 32 (u / 16) + u % 16
// (tools/microbench/tc_m64_probe.cu).
__device__ __forceinline__ void issue_mask_wgrad64_w(uint32_t el, uint32_t d_tmem, uint32_t a_base, uint32_t b_base, uint32_t b_lo_off,
                                                     int R, int N, uint32_t accumulate) {
  const uint32_t idesc = make_idesc(64, N, 1, 1);
  const uint32_t hi = (uint32_t)R | DESC_VERSION_HI;   // MN-major: SBO = R * 16 B, LBO = 128 B
  const int ks = R >> 4;
  for (int t = 0; t < 2; ++t) {
    uint32_t alo = ((a_base >> 4) & 0x3FFFu) | (8u << 16);
    uint32_t blo = (((b_base + (t == 1 ? b_lo_off : 0u)) >> 4) & 0x3FFFu) | (8u << 16);
    for (int k = 0; k < ks; ++k) {
      mma_f16_w(el, d_tmem, pack64(alo, hi), pack64(blo, hi), idesc, accumulate);
      accumulate = 1;
      alo += 16u;
      blo += 16u;
    }
  }
}

}  // namespace tc
}  // namespace dpv
