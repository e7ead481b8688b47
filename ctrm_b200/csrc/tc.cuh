// tcgen05 building blocks for the small fused MLP kernels (sm_100a only): shared-memory matrix descriptors for
// the un-swizzled K-major canonical layout, fp32-exact BF16x3 operand splitting, MMA issue, TMEM allocation / loads,
// mbarriers.
//
// Operand layout (SWIZZLE_NONE, K-major; 16-bit elements, T = 8 per 16 bytes):
//   a core matrix is 8 rows x 16 bytes stored contiguously (128 B); core matrices adjacent in K are LBO = 128 B apart,
//   8-row groups are SBO = (K/8) * 128 B apart:
//       byte_offset(r, k) = (r / 8) * SBO + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2
//   One tcgen05.mma.kind::f16 consumes K = 16 (two core matrices), so k-step s starts 256 * s bytes further.
// fp32 accuracy ("BF16x3"): every fp32 operand x is split EXACTLY into three bf16 parts x = x1 + x2 + x3
// (x1 = rn_bf16(x), x2 = rn_bf16(x - x1), x3 = x - x1 - x2; 3 x 8 significant bits = fp32's 24), and
//   D += a1*b3 + a3*b1 + a2*b2 + a1*b2 + a2*b1 + a1*b1
// is accumulated in fp32 in TMEM, smallest terms first.  The dropped terms (a2*b3, a3*b2, a3*b3) are <= 2^-24 relative per
// product, i.e. below fp32 rounding.  Inputs that are exact in bf16 (the {0,1} FOV tensor) need only b1, b2, b3.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- operand placement ----
__host__ __device__ constexpr uint32_t sbo_bytes(int K) { return (uint32_t)(K / 8) * 128u; }
__host__ __device__ constexpr uint32_t tile_bytes(int rows, int K) { return (uint32_t)rows * (uint32_t)K * 2u; }
__device__ __forceinline__ uint32_t elem_offset(int r, int k, int K) {  // bytes
  return (uint32_t)(r >> 3) * sbo_bytes(K) + (uint32_t)(k >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 7) * 2u;
}
// exact three-way bf16 split of an fp32 value; returns the three 16-bit patterns
__device__ __forceinline__ void split3(float x, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  const __nv_bfloat16 b1 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(b1);
  const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(b2);
  const __nv_bfloat16 b3 = __float2bfloat16_rn(r2);
  p1 = (uint32_t)__bfloat16_as_ushort(b1);
  p2 = (uint32_t)__bfloat16_as_ushort(b2);
  p3 = (uint32_t)__bfloat16_as_ushort(b3);
}
// store 8 consecutive k (k0 % 8 == 0) of row r as three bf16 parts into the three operand tiles (one 16-byte store each)
__device__ __forceinline__ void store_split8_body(uint8_t* t1, uint8_t* t2, uint8_t* t3, uint32_t off, float v0, float v1, float v2,
                                               float v3, float v4, float v5, float v6, float v7) {
  const float v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
  uint32_t w1[4], w2[4], w3[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t a1, a2, a3, b1, b2, b3;
    split3(v[2 * j], a1, a2, a3);
    split3(v[2 * j + 1], b1, b2, b3);
    w1[j] = a1 | (b1 << 16);
    w2[j] = a2 | (b2 << 16);
    w3[j] = a3 | (b3 << 16);
  }
  *reinterpret_cast<uint4*>(t1 + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  *reinterpret_cast<uint4*>(t2 + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
  *reinterpret_cast<uint4*>(t3 + off) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
}
__device__ __forceinline__ void store_split8(uint8_t* t1, uint8_t* t2, uint8_t* t3, int r, int k0, int K, const float* v) {
  store_split8_body(t1, t2, t3, elem_offset(r, k0, K), v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
}

// The same exact three-way split for ACTIVATIONS, by truncation instead of rounding: each part keeps the next 8 significant
// bits (x = x1 + x2 + x3 still holds exactly: 3 x 8 = 24), the packed words come straight out of the fp32 registers with
// one byte permute per pair and part, and the residuals cost an AND and a subtraction instead of two conversions.  With the
// weights split by rounding (|b2| <= 2^-9 |b|, |b3| <= 2^-17 |b|) the dropped products a2*b3, a3*b2 stay <= 2^-24 |a||b|.
__device__ __forceinline__ void store_act8(uint8_t* t1, uint8_t* t2, uint8_t* t3, int r, int k0, int K, const float* v) {
  uint32_t w1[4], w2[4], w3[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = v[2 * j], b = v[2 * j + 1];
    const uint32_t ab = __float_as_uint(a), bb = __float_as_uint(b);
    w1[j] = __byte_perm(ab, bb, 0x7632);
    const float ra = a - __uint_as_float(ab & 0xFFFF0000u), rb = b - __uint_as_float(bb & 0xFFFF0000u);
    const uint32_t rab = __float_as_uint(ra), rbb = __float_as_uint(rb);
    w2[j] = __byte_perm(rab, rbb, 0x7632);
    const float sa = ra - __uint_as_float(rab & 0xFFFF0000u), sb = rb - __uint_as_float(rbb & 0xFFFF0000u);
    w3[j] = __byte_perm(__float_as_uint(sa), __float_as_uint(sb), 0x7632);
  }
  const uint32_t off = elem_offset(r, k0, K);
  *reinterpret_cast<uint4*>(t1 + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  *reinterpret_cast<uint4*>(t2 + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
  *reinterpret_cast<uint4*>(t3 + off) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
}

// ---- descriptors ----
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, int K) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);                  // start address      [0,14)
  d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 16;             // leading byte off.  [16,30)  K-adjacent core matrices
  d |= (uint64_t)((sbo_bytes(K) >> 4) & 0x3FFFu) << 32;     // stride byte off.   [32,46)  8-row groups
  d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
  return d;                                                 // layout type [61,64) = 0: SWIZZLE_NONE
}
// kind::f16 with BF16 inputs, fp32 accumulate, both operands K-major: c_format F32 (bit 4), a/b_format BF16 (1) at bits
// 7 / 10, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA ----
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[128 x N] = A[128 x K] * B[N x K]^T, fp32-exact products (BF16x3, 6 MMAs per 16-wide k-step); one thread issues.
// a / b are shared-memory byte addresses of the first of three consecutive operand tiles (parts 1, 2, 3), a_part / b_part
// bytes apart.  The tensor core TRUNCATES when it adds into the fp32 accumulator: with all six partial products in one
// accumulator (30 accumulations for K = 80) the CVAE output showed outliers of 2.3e-5.  The main term a1*b1 therefore gets
// an accumulator of its own (K/16 accumulations), every correction term goes to a second one n_cols TMEM columns further,
// and the epilogue adds the two (tmem_ld*_sum2).
// What matters is that nothing small is ever added to the full-magnitude sum inside the tensor core; mixing the 2^-8 and
// 2^-16 correction classes in one accumulator costs 2^-24 * 2^-8 relative (three separate accumulators measured the same).
static __device__ __noinline__ void mma_bf16x3_c2(uint32_t tmem_d, uint32_t n_cols, uint32_t a, uint32_t a_part, uint32_t b,
                                              uint32_t b_part, int K, int N) {
  const uint32_t id = idesc_bf16(128, N);
  uint32_t acc = 0u;
  for (int s = 0; s < K / 16; ++s) {
    const uint32_t o = 256u * s;
    const uint64_t a1 = smem_desc(a + o, K), a2 = smem_desc(a + a_part + o, K), a3 = smem_desc(a + 2 * a_part + o, K);
    const uint64_t b1 = smem_desc(b + o, K), b2 = smem_desc(b + b_part + o, K), b3 = smem_desc(b + 2 * b_part + o, K);
    mma_bf16(tmem_d + n_cols, a1, b3, id, acc);
    mma_bf16(tmem_d + n_cols, a3, b1, id, 1u);
    mma_bf16(tmem_d + n_cols, a2, b2, id, 1u);
    mma_bf16(tmem_d + n_cols, a1, b2, id, 1u);
    mma_bf16(tmem_d + n_cols, a2, b1, id, 1u);
    mma_bf16(tmem_d, a1, b1, id, acc);
    acc = 1u;
  }
}
// A exactly representable in bf16 (e.g. {0,1} inputs): only B is split (3 MMAs per k-step)
__device__ __forceinline__ void mma_bf16_exactA(uint32_t tmem_d, uint32_t a, uint32_t b, uint32_t b_part, int K, int N,
                                                bool accumulate) {
  const uint32_t id = idesc_bf16(128, N);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int s = 0; s < K / 16; ++s) {
    const uint32_t o = 256u * s;
    const uint64_t a1 = smem_desc(a + o, K);
    mma_bf16(tmem_d, a1, smem_desc(b + 2 * b_part + o, K), id, acc);
    mma_bf16(tmem_d, a1, smem_desc(b + b_part + o, K), id, 1u);
    mma_bf16(tmem_d, a1, smem_desc(b + o, K), id, 1u);
    acc = 1u;
  }
}
// ---- 8-bit integer operands (kind::i8; sm_100a has it, sm_103a does not) ----
// Same canonical K-major layout with 16 elements per 16-byte row of a core matrix:
//     byte_offset(r, k) = (r / 8) * SBO + (k / 16) * 128 + (r % 8) * 16 + (k % 16),   SBO = (K / 16) * 128
// one MMA consumes K = 32 (two core matrices, 256 bytes further per k-step); accumulators are int32, exact.
__host__ __device__ constexpr uint32_t sbo_bytes8(int K) { return (uint32_t)(K / 16) * 128u; }
__host__ __device__ constexpr uint32_t tile_bytes8(int rows, int K) { return (uint32_t)rows * (uint32_t)K; }
__host__ __device__ __forceinline__ uint32_t elem_offset8(int r, int k, int K) {
  return (uint32_t)(r >> 3) * sbo_bytes8(K) + (uint32_t)(k >> 4) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 15);
}
__device__ __forceinline__ uint64_t smem_desc8(uint32_t saddr, int K) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes8(K) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// A unsigned 8-bit (a_format 0), B signed 8-bit (b_format 1), int32 accumulate (c_format 2)
__host__ __device__ constexpr uint32_t idesc_u8s8(int M, int N) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// four signed base-128 digits: b * scale in [-64, 64] (scale a power of two), b = sum_j q_j 2^(-7 j) / scale up to
// 2^-28 / scale; every q_j is an integer in [-64, 64]
__device__ __forceinline__ void digits4_s7(float b, float scale, int* q) {
  float t = b * scale;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float d = rintf(t);
    q[j] = (int)d;
    t = (t - d) * 128.f;  // exact
  }
}

__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

// ---- exact fixed-point digit products ("Ozaki" splitting) ----
// When fp32 accumulation itself is the error source (a well-scaled input against weights of magnitude ~50 cancelling to an
// O(1) result: the CVAE decoder's first layer), both operands are cut into 8-bit FIXED-POINT digits, each exactly a bf16:
//     a = sum_i qa_i 2^(-8 i)          (a in [0, 1],   qa_i integers in [0, 256],  i = 1..4)
//     b = sum_j qb_j 2^(e + 1 - 8 j)   (|b| < 2^e,     qb_j integers in [-128, 128], j = 1..4)
// Every digit product is an integer below 2^15; a K <= 64 dot product of up to four digit pairs stays below 2^24, so the
// fp32 accumulator holds it EXACTLY.  Pairs with equal i + j share one accumulator g = i + j in {2, 3, 4, 5} with scale
// 2^(e + 1 - 8 g); the epilogue combines the four accumulators in fp64.  Truncation: 2^-32 absolute in a, 2^(e-32) in b.
__device__ __forceinline__ void digits4_signed(float b, float scale, uint32_t* q) {  // |b * scale| <= 128
  float t = b * scale;  // power-of-two scale: exact
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float d = rintf(t);
    q[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(d));
    t = (t - d) * 256.f;
  }
}
// 10 MMAs per 16-wide k-step into four accumulators (32-bit columns tmem_d + n_cols * (g - 2)); a / b point at four
// consecutive digit tiles
static __device__ __noinline__ void mma_digits4(uint32_t tmem_d, uint32_t n_cols, uint32_t a, uint32_t a_part, uint32_t b,
                                            uint32_t b_part, int K, int N) {
  const uint32_t id = idesc_bf16(128, N);
  for (int s = 0; s < K / 16; ++s) {
    const uint32_t o = 256u * s;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i + j > 3) continue;  // digits i+1, j+1 with (i+1)+(j+1) <= 5
        const bool first = (s == 0) && (i == 0);  // the first MMA into accumulator g = i + j is (i = 0, j = g) at s = 0
        mma_bf16(tmem_d + n_cols * (uint32_t)(i + j), smem_desc(a + i * a_part + o, K), smem_desc(b + j * b_part + o, K), id,
                 first ? 0u : 1u);
      }
  }
}

// ---- fences ----
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(mbar),
      "r"(parity), "r"(0x989680u)  // suspend-time hint: the waiting warp sleeps in hardware instead of polling
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion counted on the mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

// one thread: bulk-copy a prebuilt operand image (16-byte aligned, size % 16 == 0) into shared memory; completion is counted on
// `mbar` (initialised with count 1: this call is its one arrival)
__device__ __forceinline__ void tma_load_image(void* dst, const uint8_t* src, uint32_t bytes, uint32_t mbar) {
  mbar_expect_tx(mbar, bytes);
  for (uint32_t o = 0; o < bytes; o += 32768u)
    tma_bulk_g2s(smem_u32(reinterpret_cast<uint8_t*>(dst) + o), src + o, bytes - o < 32768u ? bytes - o : 32768u, mbar);
}
// the reverse, for the one-off kernels that build such an image in their own shared memory
__device__ __forceinline__ void dump_image(const uint8_t* smem_src, uint8_t* __restrict__ dst, uint32_t bytes, int tid, int nthreads) {
  for (uint32_t i = (uint32_t)tid * 16u; i < bytes; i += (uint32_t)nthreads * 16u)
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(smem_src + i);
}

// same, with an L2 evict-first hint for data that is consumed exactly once
__device__ __forceinline__ void tma_bulk_g2s_evict_first(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   dst_smem),
               "l"(src), "r"(bytes), "r"(mbar), "l"(policy)
               : "memory");
}

// ---- TMEM ----
// whole warp executes; ncols power of two >= 32; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// lane i of warp w reads TMEM lane 32*(w%4)+i, 32 consecutive fp32 columns starting at column `col` of `tbase`
__device__ __forceinline__ void tmem_ld32(uint32_t tbase, int warp, int col, float* v) {
  const uint32_t addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-column variant
__device__ __forceinline__ void tmem_ld16(uint32_t tbase, int warp, int col, float* v) {
  const uint32_t addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// two-accumulator epilogues of mma_bf16x3_c2: v = corrections + main
__device__ __forceinline__ void tmem_ld32_sum2(uint32_t tbase, int warp, int col, int n_cols, float* v) {
  float t[32];
  tmem_ld32(tbase, warp, col + n_cols, v);
  tmem_ld32(tbase, warp, col, t);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] += t[i];
}
__device__ __forceinline__ void tmem_ld16_sum2(uint32_t tbase, int warp, int col, int n_cols, float* v) {
  float t[16];
  tmem_ld16(tbase, warp, col + n_cols, v);
  tmem_ld16(tbase, warp, col, t);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += t[i];
}


// copy rows [k_begin, k_begin + K_real) of a blob weight matrix stored [in][ld] (fp32, global) into three consecutive
// bf16 B-operand tiles (N_tile x K, K-major, tile_bytes(N_tile, K) apart) at tile rows [n_off, n_off + N):
// B[n_off + n][k] = w[(k_begin + k) * ld + n]; k >= K_real and n >= N_real are zero padding
static __device__ __noinline__ void stage_weight_split_at(const float* __restrict__ w, int ld, int k_begin, int K_real, int N_real,
                                                      int N, int n_off, int N_tile, int K, uint8_t* tiles, int tid, int nthreads,
                                                      int k_off = 0, int K_blk = -1) {
  const uint32_t part = tile_bytes(N_tile, K);
  if (K_blk < 0) K_blk = K;  // columns [k_off, k_off + K_blk) of the tile are written (K_blk % 8 == 0)
  for (int i = tid; i < N * (K_blk / 8); i += nthreads) {
    const int k0 = (i / N) * 8, n = i - (i / N) * N;  // consecutive threads -> consecutive n (coalesced reads)
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (n < N_real && k0 + j < K_real) ? __ldg(w + (size_t)(k_begin + k0 + j) * ld + n) : 0.f;
    store_split8(tiles, tiles + part, tiles + 2 * part, n_off + n, k_off + k0, K, v);
  }
}
__device__ __forceinline__ void stage_weight_split(const float* __restrict__ w, int ld, int k_begin, int K_real, int N_real,
                                                   int N, int K, uint8_t* tiles, int tid, int nthreads) {
  stage_weight_split_at(w, ld, k_begin, K_real, N_real, N, 0, N, K, tiles, tid, nthreads);
}

// four signed fixed-point digit tiles of rows [k_begin, k_begin + K_real) of a blob weight matrix [in][ld]
static __device__ __noinline__ void stage_weight_digits(const float* __restrict__ w, int ld, int k_begin, int K_real, int N_real,
                                                    int N, int K, float scale, uint8_t* tiles, int tid, int nthreads) {
  const uint32_t part = tile_bytes(N, K);
  for (int i = tid; i < N * (K / 8); i += nthreads) {
    const int k0 = (i / N) * 8, n = i - (i / N) * N;
    uint32_t dg[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = (n < N_real && k0 + j < K_real) ? __ldg(w + (size_t)(k_begin + k0 + j) * ld + n) : 0.f;
      digits4_signed(v, scale, dg[j]);
    }
    const uint32_t off = elem_offset(n, k0, K);
#pragma unroll
    for (int d = 0; d < 4; ++d)
      *reinterpret_cast<uint4*>(tiles + d * part + off) =
          make_uint4(dg[0][d] | (dg[1][d] << 16), dg[2][d] | (dg[3][d] << 16), dg[4][d] | (dg[5][d] << 16),
                     dg[6][d] | (dg[7][d] << 16));
  }
}
// row r, 8 consecutive k of an input in [0, 1] as four digit tiles
__device__ __forceinline__ void store_digits8_unit_body(uint8_t* tiles, uint32_t part, uint32_t off, float v0, float v1, float v2,
                                                     float v3, float v4, float v5, float v6, float v7) {
  const float v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
  // a digit is an integer in [0, 256]: at most 8 significant bits, so the upper half of its fp32 pattern IS its bf16 pattern
  // and two digits pack with one byte permute (no conversion instruction)
  float t[8];
  uint32_t w[4][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = v[j] * 256.f;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float da = floorf(t[2 * j]), db = floorf(t[2 * j + 1]);
      w[d][j] = __byte_perm(__float_as_uint(da), __float_as_uint(db), 0x7632);
      t[2 * j] = (t[2 * j] - da) * 256.f;        // exact
      t[2 * j + 1] = (t[2 * j + 1] - db) * 256.f;
    }
  }
#pragma unroll
  for (int d = 0; d < 4; ++d)
    *reinterpret_cast<uint4*>(tiles + d * part + off) = make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]);
}
__device__ __forceinline__ void store_digits8_unit(uint8_t* tiles, uint32_t part, int r, int k0, int K, const float* v) {
  store_digits8_unit_body(tiles, part, elem_offset(r, k0, K), v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
}

}  // namespace tc
