// FOV encoders on the tensor cores.  reference: FOVEncoder.forward x 2 (learning/model/fov_encoder.py:34-59;
// learning/formats/format_input.py:311-312) + the fov_vain rows of the AgentEncoder's first layer
// (learning/model/agent_encoder.py:57-60; format_input.py:328-331).
//
// The FOV tensor is binary, so it crosses HBM as BITS: [tile of 128 agents][chunk of 96 inputs][row][12 bytes] (1,536 B per
// chunk, written by the raster kernel, raster.cu).  The TMA producer drops a chunk's bits into the tail of the stage's A area
// and four otherwise idle warps expand them in place to the UNSIGNED 8-BIT canonical UMMA tile [128 x 96] (12,288 B) while the
// previous chunk's MMAs run.
// First layers of both encoders side by side: D[128 x 64] = FOV[128 x 2F^2] . W0[2F^2 x 64] on the INTEGER tensor cores
// (tcgen05.mma kind::i8): the weights are cut into four signed base-128 fixed-point digits (tc.cuh), every digit product
// is an integer and the int32 accumulators are exact; the four digit accumulators are combined in fp64.  One byte per
// operand element instead of two halves the bytes a tile pulls through its TMA (cp.async.bulk) + mbarrier pipeline, which
// is what bounds this kernel (one CTA per SM waiting on L2 round trips), and lets the ring hold four stages instead of two:
// one thread produces, one thread issues the MMAs, tcgen05.commit releases the stage.
// Tails (32->32->32 per encoder as block-diagonal 64x64, then the 32x32 vain projection) are BF16x3 rounds on the same tile.
#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

#define FT_THREADS 256
#define FT_CHUNK 96
#define FT_A_CHUNK 12288u   // tile_bytes8(128, 96)
#define FT_A_BITS 1536u     // the same chunk as bits: 128 rows x 12 bytes
#define FT_W_DIGIT 6144u    // tile_bytes8(64, 96)
#define FT_W_CHUNK 24576u   // four digits
#define FT_STAGES 3
#define FT_STAGE (FT_A_CHUNK + FT_W_CHUNK)
#define FT_H_PART 16384u    // tile_bytes(128, 64)
#define FT_H32_PART 8192u   // tile_bytes(128, 32)
#define FT_W64_PART 8192u   // tile_bytes(64, 64)
#define FT_W32_PART 2048u   // tile_bytes(32, 32)
#define FT_NF 256
#define FT_TAIL_OFF 49152u  // after layer 0 the ring memory holds the hidden tiles [0, 48 KB) and the tail-layer weights behind them
#define FT_TAILW_BYTES (2 * 3 * FT_W64_PART + 3 * FT_W32_PART)
// 109 KB: TWO CTAs per SM.  The kernel waits on L2 round trips (operand ring) and on its three dependent tail rounds, so a
// second resident CTA nearly doubles what an SM gets done; to fit, the 54 KB of tail-layer weights are not resident but
// pulled again (TMA, from L2) for every tile into ring memory that is dead once layer 0 has completed.
#define FT_SMEM (FT_STAGES * FT_STAGE + FT_NF * 4 + 128)

// ---- weight preparation (once per ctrm_load_weights): digit tiles of W0 in global memory, chunk-major ----
__global__ void fov_w0_prep_kernel(const float* __restrict__ w0 /*[KD][64]*/, int KD, int n_chunks, float scale,
                                   uint8_t* __restrict__ out /*[n_chunks][4][12288]*/) {
  const int total = n_chunks * 64 * (FT_CHUNK / 16);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i / (64 * (FT_CHUNK / 16));
    const int rem = i - c * 64 * (FT_CHUNK / 16);
    const int k0 = (rem / 64) * 16, n = rem % 64;
    uint32_t w[4][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = c * FT_CHUNK + k0 + j;
      int q[4];
      tc::digits4_s7(k < KD ? w0[(size_t)k * 64 + n] : 0.f, scale, q);
#pragma unroll
      for (int d = 0; d < 4; ++d) w[d][j >> 2] |= ((uint32_t)q[d] & 0xFFu) << (8 * (j & 3));
    }
    const uint32_t off = tc::elem_offset8(n, k0, FT_CHUNK);
#pragma unroll
    for (int d = 0; d < 4; ++d)
      *reinterpret_cast<uint4*>(out + (size_t)c * FT_W_CHUNK + d * FT_W_DIGIT + off) = make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]);
  }
}

// ---- f32 FOV tensor [A][KD] -> bit tiles (stage entry point ctrm_encode_features only) ----
__global__ void fov_f32_to_tiles_kernel(const float* __restrict__ fov, int A, int KD, int n_chunks, uint8_t* __restrict__ tiles) {
  const int nw = n_chunks * 3;  // 32-bit words per row
  const long long total = (long long)((A + 127) / 128) * 128 * nw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i / nw), w = (int)(i - (long long)a * nw);
    uint32_t bits = 0u;
    if (a < A) {
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int k = 32 * w + j;
        if (k < KD && fov[(size_t)a * KD + k] != 0.f) bits |= 1u << j;  // the FOV tensor is {0, 1}
      }
    }
    const int c = w / 3, jw = w - 3 * c;
    *reinterpret_cast<uint32_t*>(tiles + ((size_t)((a >> 7) * n_chunks + c) * 128 + (a & 127)) * 12 + 4 * jw) = bits;
  }
}

// tail-layer weight operands of fov_encode exactly as they sit in its shared memory from W1c on
#define FT_W_BYTES (2 * 3 * FT_W64_PART + 3 * FT_W32_PART + 224 * 4)
__global__ void __launch_bounds__(FT_THREADS) fov_encode_image_kernel(const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                                    uint8_t* __restrict__ image) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* W1c = smem;
  uint8_t* W2c = W1c + 3 * FT_W64_PART;
  uint8_t* Wp = W2c + 3 * FT_W64_PART;
  float* cf = reinterpret_cast<float*>(Wp + 3 * FT_W32_PART);
  float *b0 = cf, *b1 = cf + 64, *b2 = cf + 128, *bp = cf + 192;
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(6 * FT_W64_PART) / 16; i += FT_THREADS) reinterpret_cast<uint4*>(W1c)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  tc::stage_weight_split_at(blob + off.fovs_w1, 32, 0, 32, 32, 32, 0, 64, 64, W1c, tid, FT_THREADS, 0, 32);
  tc::stage_weight_split_at(blob + off.fovv_w1, 32, 0, 32, 32, 32, 32, 64, 64, W1c, tid, FT_THREADS, 32, 32);
  tc::stage_weight_split_at(blob + off.fovs_w2, 32, 0, 32, 32, 32, 0, 64, 64, W2c, tid, FT_THREADS, 0, 32);
  tc::stage_weight_split_at(blob + off.fovv_w2, 32, 0, 32, 32, 32, 32, 64, 64, W2c, tid, FT_THREADS, 32, 32);
  tc::stage_weight_split(blob + off.ag_w0v, 32, 0, 32, 32, 32, 32, Wp, tid, FT_THREADS);
  if (tid < 64) b0[tid] = __ldg(&blob[off.fov_b0 + tid]);
  if (tid < 32) {
    b1[tid] = __ldg(&blob[off.fovs_b1 + tid]);
    b1[32 + tid] = __ldg(&blob[off.fovv_b1 + tid]);
    b2[tid] = __ldg(&blob[off.fovs_b2 + tid]);
    b2[32 + tid] = __ldg(&blob[off.fovv_b2 + tid]);
    bp[tid] = __ldg(&blob[off.ag_b0 + tid]);
  }
  __syncthreads();
  tc::dump_image(smem, image, FT_W_BYTES, tid, FT_THREADS);
}

__global__ void __launch_bounds__(FT_THREADS, 2) fov_encode_tc_kernel(const uint8_t* __restrict__ fov_tiles,
                                                                   const uint8_t* __restrict__ w0_tiles, int n_chunks, int w0_exp,
                                                                   const int32_t* __restrict__ agent_inst,
                                                                   const uint8_t* __restrict__ skip_inst, int A,
                                                                   const uint8_t* __restrict__ w_image,
                                                                   float* __restrict__ e_self, float* __restrict__ e_vain,
                                                                   float* __restrict__ vain_proj, int n_tiles,
                                                                   const int32_t* __restrict__ live_agent,
                                                                   const int32_t* __restrict__ n_live) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage0 = smem;                              // stage s: A chunk | W chunk (4 digits)
  uint8_t* W1c = smem + FT_TAIL_OFF;                   // blockdiag(self.mlp.2, vain.mlp.2)  [64 x 64] x 3   (per tile)
  uint8_t* W2c = W1c + 3 * FT_W64_PART;                // blockdiag(self.mlp.4, vain.mlp.4)
  uint8_t* Wp = W2c + 3 * FT_W64_PART;                 // agent.mlp.0 rows 11..42  [32 x 32] x 3
  float* cf = reinterpret_cast<float*>(smem + FT_STAGES * FT_STAGE);
  float* b0 = cf;          // 64
  float* b1 = cf + 64;     // 64
  float* b2 = cf + 128;    // 64
  float* bp = cf + 192;    // 32
  // full[FT_STAGES], empty[FT_STAGES], done, tail weights landed, ready[FT_STAGES] (A chunk expanded: 128 arrivals)
  uint64_t* bars = reinterpret_cast<uint64_t*>(cf + FT_NF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * FT_STAGES + 2);
  uint8_t* Ht = stage0;    // hidden tiles reuse the operand stages once layer 0 has completed
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;  // two threads per row; TMEM lane group = warp % 4 for both halves

  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    for (int i = 0; i < 2 * FT_STAGES + 2; ++i) tc::mbar_init(tc::smem_u32(&bars[i]), 1);
    for (int i = 0; i < FT_STAGES; ++i) tc::mbar_init(tc::smem_u32(&bars[2 * FT_STAGES + 2 + i]), 128);
    // biases: the constant tail of the prebuilt image (fov_encode_image_kernel), through the TMA engine
    tc::tma_load_image(cf, w_image + FT_TAILW_BYTES, 224 * 4, tc::smem_u32(&bars[2 * FT_STAGES]));
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t full0 = tc::smem_u32(&bars[0]), empty0 = tc::smem_u32(&bars[FT_STAGES]),
                 done = tc::smem_u32(&bars[2 * FT_STAGES]);
  const uint32_t hb = tc::smem_u32(&bars[2 * FT_STAGES + 1]);
  const uint32_t ready0 = tc::smem_u32(&bars[2 * FT_STAGES + 2]);
  uint32_t hphase = 0;
  tc::mbar_wait(done, 0);  // biases have landed; the tail MMA rounds continue on the same barrier with phase 1
  const uint32_t sStage = tc::smem_u32(stage0), sH = tc::smem_u32(Ht), sW1 = tc::smem_u32(W1c), sW2 = tc::smem_u32(W2c),
                 sWp = tc::smem_u32(Wp);
  const uint32_t id64 = tc::idesc_u8s8(128, 64);
  uint32_t it = 0;          // chunks processed so far by this CTA (ring position)
  uint32_t done_phase = 1;
  double dscale[4];         // digit j (0-based) weighs 2^(w0_exp - 6 - 7 j)
#pragma unroll
  for (int d = 0; d < 4; ++d) dscale[d] = exp2((double)(w0_exp - 6 - 7 * d));

  auto mma_round = [&](uint32_t tcol, uint32_t a, uint32_t a_part, uint32_t b, uint32_t b_part, int Kk, int Nn) {
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::mbar_wait(hb, hphase);  // this tile's tail weights (only the issuing thread needs to know)
      tc::fence_after_sync();
      tc::mma_bf16x3_c2(tbase + tcol, (uint32_t)Nn, a, a_part, b, b_part, Kk, Nn);
      tc::commit(done);
    }
    tc::mbar_wait(done, done_phase);
    done_phase ^= 1u;
    tc::fence_after_sync();
  };

  const int n_rows = live_agent != nullptr ? *n_live : A;
  n_tiles = min(n_tiles, (n_rows + 127) >> 7);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row = tile * 128 + r;  // row of the operand tile; its agent comes from the live list when there is one
    bool live = row < n_rows;
    const int a = live ? (live_agent != nullptr ? live_agent[row] : row) : 0;
    if (live && skip_inst != nullptr) live = skip_inst[agent_inst[a]] == 0;
    if (!__syncthreads_or(live)) continue;
    // ---- layer 0: streamed exact digit products
    if (tid == 0) {  // producer
      const uint8_t* gA = fov_tiles + (size_t)tile * n_chunks * FT_A_BITS;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t u = it + c, s = u % FT_STAGES;
        if (u >= FT_STAGES) tc::mbar_wait(empty0 + 8 * s, ((u / FT_STAGES) - 1) & 1u);
        tc::mbar_expect_tx(full0 + 8 * s, FT_A_BITS + FT_W_CHUNK);
        // the chunk's bits land in the LAST 1,536 bytes of the stage's A area and are expanded in place below
        tc::tma_bulk_g2s_evict_first(sStage + s * FT_STAGE + (FT_A_CHUNK - FT_A_BITS), gA + (size_t)c * FT_A_BITS, FT_A_BITS,
                                     full0 + 8 * s);
        tc::tma_bulk_g2s(sStage + s * FT_STAGE + FT_A_CHUNK, w0_tiles + (size_t)c * FT_W_CHUNK, FT_W_CHUNK, full0 + 8 * s);
      }
    } else if (tid >= 128) {  // expanders: thread = row; 96 bits -> 96 bytes {0, 1} in the canonical 8-bit UMMA layout
      const int er = tid - 128;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t u = it + c, s = u % FT_STAGES;
        uint8_t* sa = stage0 + (size_t)s * FT_STAGE;
        tc::mbar_wait(full0 + 8 * s, (u / FT_STAGES) & 1u);
        const uint32_t* bp = reinterpret_cast<const uint32_t*>(sa + (FT_A_CHUNK - FT_A_BITS) + er * 12);
        const uint32_t bw[3] = {bp[0], bp[1], bp[2]};
        // rows 112..127 expand INTO the region the bits arrived in: everybody reads before anybody writes
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint8_t* dst = sa + (size_t)(er >> 3) * tc::sbo_bytes8(FT_CHUNK) + (er & 7) * 16;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          const uint32_t b16 = (bw[g >> 1] >> (16 * (g & 1))) & 0xFFFFu;
          // a nibble times 0x00204081 puts bit i at bit 8 i: four bytes of 0 / 1
          *reinterpret_cast<uint4*>(dst + g * 128) =
              make_uint4(((b16 & 0xFu) * 0x00204081u) & 0x01010101u, (((b16 >> 4) & 0xFu) * 0x00204081u) & 0x01010101u,
                         (((b16 >> 8) & 0xFu) * 0x00204081u) & 0x01010101u, ((b16 >> 12) * 0x00204081u) & 0x01010101u);
        }
        tc::fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        tc::mbar_arrive(ready0 + 8 * s);
      }
    } else if (tid == 32) {  // MMA issuer
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t u = it + c, s = u % FT_STAGES;
        tc::mbar_wait(ready0 + 8 * s, (u / FT_STAGES) & 1u);  // the weights landed (full) and the A chunk is expanded
        tc::fence_after_sync();
        const uint32_t sa = sStage + s * FT_STAGE, sw = sa + FT_A_CHUNK;
#pragma unroll
        for (int ks = 0; ks < FT_CHUNK / 32; ++ks) {
          const uint64_t ad = tc::smem_desc8(sa + 256u * ks, FT_CHUNK);
          const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
          for (int d = 0; d < 4; ++d)
            tc::mma_i8(tbase + 64 * d, ad, tc::smem_desc8(sw + d * FT_W_DIGIT + 256u * ks, FT_CHUNK), id64, acc);
        }
        tc::commit(empty0 + 8 * s);
      }
      tc::commit(done);
    }
    it += (uint32_t)n_chunks;
    tc::mbar_wait(done, done_phase);
    done_phase ^= 1u;
    tc::fence_after_sync();
    // layer 0 has completed: nothing reads the ring any more.  This tile's tail-layer weights go behind the hidden tiles
    // while the first epilogue runs.
    if (tid == 0) tc::tma_load_image(W1c, w_image, FT_TAILW_BYTES, hb);
    // ---- epilogue 0: combine digits in fp64, + bias, ReLU -> hidden tile [128 x 64].  Thread (row r, half) owns the
    //      32 columns of ONE encoder: half 0 = fov_encoder_self, half 1 = fov_encoder_vain (16 columns at a time)
    float h[32];
#pragma unroll
    for (int p16 = 0; p16 < 2; ++p16) {
      double pre[16];
      float hv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pre[j] = 0.0;
#pragma unroll
      for (int d = 3; d >= 0; --d) {
        tc::tmem_ld16(tbase, warp, 64 * d + 32 * half + 16 * p16, hv);
#pragma unroll
        for (int j = 0; j < 16; ++j) pre[j] = fma((double)__float_as_int(hv[j]), dscale[d], pre[j]);  // int32 accumulators
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) hv[j] = fmaxf((float)pre[j] + b0[32 * half + 16 * p16 + j], 0.f);
      // the hidden tile lies in ring memory that layer 0 has finished with (waited on above)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        tc::store_act8(Ht, Ht + FT_H_PART, Ht + 2 * FT_H_PART, r, 32 * half + 16 * p16 + q * 8, 64, hv + q * 8);
    }
    // ---- tail 1 (the layer-0 accumulators have been read by everyone before the round's barrier)
    mma_round(0, sH, FT_H_PART, sW1, FT_W64_PART, 64, 64);
    tc::tmem_ld32_sum2(tbase, warp, 0 + 32 * half, 64, h);
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j] + b1[32 * half + j], 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) tc::store_act8(Ht, Ht + FT_H_PART, Ht + 2 * FT_H_PART, r, 32 * half + q * 8, 64, h + q * 8);
    // ---- tail 2 -> e_self (half 0) | e_vain (half 1)
    mma_round(128, sH, FT_H_PART, sW2, FT_W64_PART, 64, 64);
    tc::tmem_ld32_sum2(tbase, warp, 128 + 32 * half, 64, h);
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] += b2[32 * half + j];
    {
      float* dstp = half == 0 ? e_self : e_vain;
      if (live && dstp != nullptr) {
        float4* dst = reinterpret_cast<float4*>(dstp + (size_t)a * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) dst[j4] = make_float4(h[j4 * 4], h[j4 * 4 + 1], h[j4 * 4 + 2], h[j4 * 4 + 3]);
      }
    }
    // ---- fov_vain part of AgentEncoder layer 0: vain_proj = b0 + W0[11:43]^T e_vain, shared by every pair (i, j)
    tc::fence_before_sync();
    __syncthreads();  // tail-2 accumulators read by everyone before the hidden tile is rewritten
    if (half == 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) tc::store_act8(Ht, Ht + FT_H32_PART, Ht + 2 * FT_H32_PART, r, q * 8, 32, h + q * 8);
    }
    mma_round(0, sH, FT_H32_PART, sWp, FT_W32_PART, 32, 32);
    tc::tmem_ld16_sum2(tbase, warp, 16 * half, 32, h);
    if (live) {
      float4* dst = reinterpret_cast<float4*>(vain_proj + (size_t)a * 32 + 16 * half);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        dst[j4] = make_float4(h[j4 * 4] + bp[16 * half + j4 * 4], h[j4 * 4 + 1] + bp[16 * half + j4 * 4 + 1],
                              h[j4 * 4 + 2] + bp[16 * half + j4 * 4 + 2], h[j4 * 4 + 3] + bp[16 * half + j4 * 4 + 3]);
    }
    hphase ^= 1u;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 256);
}

// ---- host side ----
int fov_n_chunks(int F) { return (2 * F * F + FT_CHUNK - 1) / FT_CHUNK; }
size_t fov_tiles_bytes(int A, int F) { return (size_t)((A + 127) / 128) * fov_n_chunks(F) * FT_A_BITS; }
size_t fov_w0_tiles_bytes(int F) { return (size_t)fov_n_chunks(F) * FT_W_CHUNK; }

int launch_fov_w0_prep(const float* d_w0, int F, int w0_exp, uint8_t* d_out, cudaStream_t st) {
  const int n_chunks = fov_n_chunks(F);
  fov_w0_prep_kernel<<<64, 256, 0, st>>>(d_w0, 2 * F * F, n_chunks, exp2f((float)(6 - w0_exp)), d_out);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_fov_f32_to_tiles(const float* d_fov, int A, int F, uint8_t* d_tiles, cudaStream_t st) {
  if (A <= 0) return 0;
  fov_f32_to_tiles_kernel<<<148 * 4, 256, 0, st>>>(d_fov, A, 2 * F * F, fov_n_chunks(F), d_tiles);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t fov_encode_image_bytes() { return FT_W_BYTES; }
int launch_fov_encode_image(const float* d_blob, const ctrm_weight_offsets_t& off, uint8_t* d_image, cudaStream_t st) {
  CTRM_CUDA_OK(cudaFuncSetAttribute(fov_encode_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_W_BYTES));
  fov_encode_image_kernel<<<1, FT_THREADS, FT_W_BYTES, st>>>(d_blob, off, d_image);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_fov_encode(const DevModel& m, const uint8_t* d_fov_tiles, const int32_t* d_agent_inst, const uint8_t* d_skip_inst, int A,
                      float* d_e_self, float* d_e_vain, float* d_vain_proj, const int32_t* d_live_agent,
                      const int32_t* d_n_live, cudaStream_t st) {
  if (A <= 0) return 0;
  const int n_tiles = (A + 127) / 128;
  static SmemAttrCache attr;
  CTRM_CUDA_OK(attr.ensure(fov_encode_tc_kernel, FT_SMEM));
  const int grid = n_tiles < 2 * 148 ? n_tiles : 2 * 148;  // persistent, two CTAs per SM
  fov_encode_tc_kernel<<<grid, FT_THREADS, FT_SMEM, st>>>(d_fov_tiles, m.fov_w0_tiles, fov_n_chunks(m.desc.fov_size), m.fov_w0_exp,
                                                         d_agent_inst, d_skip_inst, A, m.fov_img, d_e_self, d_e_vain,
                                                         d_vain_proj, n_tiles, d_live_agent, d_n_live);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
