// CVAE decoder on the 5th-generation tensor cores (tcgen05 + TMEM), fused with the Gumbel-softmax that produces its
// input.  reference: RelaxedOneHotCategorical(temp).rsample() + CTRMNet.decoder (learning/model/cvae.py:72-76, :133-137).
//
// One CTA = 256 threads = one 128-row tile at a time (row = (agent, copy)); two threads share row r / TMEM lane r.
//   1. thread r: Philox uniforms -> Gumbel noise -> z = softmax((logits + g)/temp)   (64 values, fp32)
//      z is split exactly into three bf16 parts and written to shared memory in the UMMA K-major canonical layout
//   2. one thread issues 4 k-steps x 6 tcgen05.mma (BF16x3): D0[128x32] = z[128x64] . W0z[64x32]  -> TMEM cols 0..31
//   3. thread r: tcgen05.ld its 32 accumulators, + dec0[agent] (x-part of layer 0 and folded BatchNorm bias, from
//      cvae_prior), ReLU, split, back to shared memory as the next A operand
//   4. D1[128x32] = h1 . W1 (2 k-steps x 6 MMAs) -> TMEM cols 32..63
//   5. thread r: ld, + b1, ReLU, the 32x3 output layer on CUDA cores, store (dx, dy, mag)
// Weights are split and staged once per (persistent) CTA.  fp32-accurate: see tc.cuh.
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"
#include "tc.cuh"

// ------------------------------------------------------------------------------------------------
// cvae_prior: everything the K copies of an agent share, one 128-agent tile per CTA, all dense layers on tcgen05 (BF16x3):
//   indicator 72->32->32->3 + argmax (roadmap_learned/learned_sampler.py:99-104), encoder_input [X, onehot] 75->32->32->64 +
//   LogSoftmax + Categorical normalisation + clamp(p) of torch.distributions' probs_to_logits (learning/model/cvae.py:57-68,
//   :132-136; emitted as sqrt(clamp(p)), see cvae_decode), and the [X, onehot] rows of the decoder's first layer (cvae.py:72-76, :137).
// The three first layers share the A operand X[128 x 72]: ONE N = 96 MMA chain computes them side by side
// (TMEM columns 0..31 indicator, 32..63 encoder, 64..95 decoder, times two accumulator classes); the one-hot indicator columns are added as a weight row
// in the epilogue.  outputs per agent: sqrtp[64], dec0[32], ind.
// ------------------------------------------------------------------------------------------------
#define PR_THREADS 256
#define PR_K0 80                      // 72 padded to a multiple of 16
#define PR_X_PART 20480u              // tile_bytes(128, 80)
#define PR_H_PART 8192u               // tile_bytes(128, 32)
#define PR_W0_PART 15360u             // tile_bytes(96, 80)
#define PR_W1_PART 2048u              // tile_bytes(32, 32)
#define PR_W2_PART 4096u              // tile_bytes(64, 32)
#define PR_NW 608                     // fp32 constants (biases, output layer, one-hot rows) that follow the W0 tiles
#define PR_W0_BYTES (3 * 15360 + PR_NW * 4)         // resident part of the weight image: W0 tiles | constants
#define PR_HEAD_BYTES (2 * 3 * 2048 + 3 * 4096)     // per-tile part: ind_w1 | enc_w1 | enc_w2 tiles
#define PR_W_BYTES (PR_W0_BYTES + PR_HEAD_BYTES)    // the prebuilt weight image
#define PR_HEAD_OFF 24576u            // the head-layer weights live in the X tile's memory behind the hidden tiles
#define PR_NF 1664                    // fp32 constants + the [2][128][4] exchange buffer of the two halves of a row
// 114 KB: TWO CTAs per SM.  The kernel is a chain of seven dependent MMA round trips per tile, so a second resident CTA
// nearly doubles what an SM gets done; to fit, only W0 stays resident and the 24 KB of head-layer weights are pulled
// again (TMA, from L2) for every tile into the part of the X tile that is dead once layer 0 has completed.
#define PR_SMEM (3 * 20480 + 3 * 15360 + PR_NF * 4 + 32)
#define PR_TMEM_COLS 256  // layer 0: 2 x 96 at 0 | head layers: 2 x 32 at 192 | encoder output 2 x 64 reuses 0..127

// weight operands of cvae_prior exactly as they sit in its shared memory from W0t on: split bf16 tiles, then fp32 constants
__device__ __forceinline__ void prior_stage_weights(uint8_t* W0t, const float* __restrict__ blob, const ctrm_weight_offsets_t& off,
                                                    int tid, int nthreads) {
  float* cf = reinterpret_cast<float*>(W0t + 3 * PR_W0_PART);
  uint8_t* Wi1 = W0t + PR_W0_BYTES;
  uint8_t* We1 = Wi1 + 3 * PR_W1_PART;
  uint8_t* We2 = We1 + 3 * PR_W1_PART;
  tc::stage_weight_split_at(blob + off.ind_w0, 32, 0, 72, 32, 32, 0, 96, PR_K0, W0t, tid, nthreads);
  tc::stage_weight_split_at(blob + off.enc_w0, 32, 0, 72, 32, 32, 32, 96, PR_K0, W0t, tid, nthreads);
  tc::stage_weight_split_at(blob + off.dec_w0, 32, 64, 72, 32, 32, 64, 96, PR_K0, W0t, tid, nthreads);
  tc::stage_weight_split(blob + off.ind_w1, 32, 0, 32, 32, 32, 32, Wi1, tid, nthreads);
  tc::stage_weight_split(blob + off.enc_w1, 32, 0, 32, 32, 32, 32, We1, tid, nthreads);
  tc::stage_weight_split(blob + off.enc_w2, 64, 0, 32, 64, 64, 32, We2, tid, nthreads);
  for (int i = tid; i < 32; i += nthreads) {
    cf[i] = __ldg(&blob[off.ind_b0 + i]);
    cf[32 + i] = __ldg(&blob[off.ind_b1 + i]);
    cf[196 + i] = __ldg(&blob[off.enc_b0 + i]);
    cf[228 + i] = __ldg(&blob[off.enc_b1 + i]);
    cf[324 + i] = __ldg(&blob[off.dec_b0 + i]);
  }
  for (int i = tid; i < 64; i += nthreads) cf[260 + i] = __ldg(&blob[off.enc_b2 + i]);
  for (int i = tid; i < 4; i += nthreads) cf[192 + i] = __ldg(&blob[off.ind_b2 + i]);
  for (int i = tid; i < 128; i += nthreads) cf[64 + i] = __ldg(&blob[off.ind_w2 + i]);
  for (int i = tid; i < 96; i += nthreads) {
    cf[356 + i] = __ldg(&blob[off.enc_w0 + CTRM_DIM_X * 32 + i]);
    cf[452 + i] = __ldg(&blob[off.dec_w0 + (64 + CTRM_DIM_X) * 32 + i]);
  }
  for (int i = 548 + tid; i < PR_NW; i += nthreads) cf[i] = 0.f;
}
__global__ void __launch_bounds__(PR_THREADS) cvae_prior_image_kernel(const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                                    uint8_t* __restrict__ image) {
  extern __shared__ __align__(1024) uint8_t smem[];
  prior_stage_weights(smem, blob, off, threadIdx.x, PR_THREADS);
  __syncthreads();
  tc::dump_image(smem, image, PR_W_BYTES, threadIdx.x, PR_THREADS);
}

__global__ void __launch_bounds__(PR_THREADS, 2) cvae_prior_tc_kernel(Batch bt, const float* __restrict__ X,
                                                                   const uint8_t* __restrict__ w_image,
                                                                   float* __restrict__ sqrtp_out, float* __restrict__ dec0_out,
                                                                   int32_t* __restrict__ ind_out, int n_tiles, int force_ind) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Xt = smem;                        // X parts 1..3 [128 x 80]; after layer 0: hidden tiles [128 x 32] x 3 | head weights
  uint8_t* W0t = Xt + 3 * PR_X_PART;         // [96 x 80] x 3 : ind_w0 | enc_w0 | dec_w0 rows 64..135
  uint8_t* Wi1 = Xt + PR_HEAD_OFF;           // ind_w1 [32 x 32] x 3   (re-loaded for every tile)
  uint8_t* We1 = Wi1 + 3 * PR_W1_PART;       // enc_w1 [32 x 32] x 3
  uint8_t* We2 = We1 + 3 * PR_W1_PART;       // enc_w2 [64 x 32] x 3
  float* cf = reinterpret_cast<float*>(W0t + 3 * PR_W0_PART);
  float* ind_b0 = cf;            // 32
  float* ind_b1 = cf + 32;       // 32
  float* ind_w2 = cf + 64;       // [32][4]
  float* ind_b2 = cf + 192;      // 4
  float* enc_b0 = cf + 196;      // 32
  float* enc_b1 = cf + 228;      // 32
  float* enc_b2 = cf + 260;      // 64
  float* dec_b0 = cf + 324;      // 32
  float* enc_oh = cf + 356;      // [3][32] rows 72..74 of enc_w0
  float* dec_oh = cf + 452;      // [3][32] rows 136..138 of dec_w0
  float* red = cf + 608;         // [2][128][4]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(cf + PR_NF);
  uint64_t* hbar = mbar + 1;                 // head weights of the current tile have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;  // two threads per row: TMEM lane group = warp % 4 for both halves

  if (tid == 0) {  // the weight operands arrive as a prebuilt image (cvae_prior_image_kernel), through the TMA engine
    tc::mbar_init(tc::smem_u32(mbar), 1);
    tc::mbar_init(tc::smem_u32(hbar), 1);
    tc::tma_load_image(W0t, w_image, PR_W0_BYTES, tc::smem_u32(mbar));
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, PR_TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t mb = tc::smem_u32(mbar);
  tc::mbar_wait(mb, 0);  // weights have landed; the MMA rounds continue on the same barrier with phase 1
  const uint32_t sX = tc::smem_u32(Xt), sW0 = tc::smem_u32(W0t), sWi1 = tc::smem_u32(Wi1), sWe1 = tc::smem_u32(We1),
                 sWe2 = tc::smem_u32(We2);
  uint32_t phase = 1, hphase = 0;
  const uint32_t hb = tc::smem_u32(hbar);
  const float eps = 1.1920928955078125e-07f;

  // one MMA round: the A tile in shared memory is complete -> issue, commit, wait.  heads = the round reads the per-tile
  // head weights, whose TMA must have landed (only the issuing thread needs to know)
  auto mma_round = [&](uint32_t tcol, uint32_t a, uint32_t a_part, uint32_t b, uint32_t b_part, int Kk, int Nn, bool heads) {
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (heads) tc::mbar_wait(hb, hphase);
      tc::fence_after_sync();
      tc::mma_bf16x3_c2(tbase + tcol, (uint32_t)Nn, a, a_part, b, b_part, Kk, Nn);
      tc::commit(mb);
    }
    tc::mbar_wait(mb, phase);
    phase ^= 1u;
    tc::fence_after_sync();
  };
  // thread (row r, half): hidden columns [16 half, 16 half + 16) -> three bf16 parts of the [128 x 32] hidden tile
  auto store_hidden16 = [&](const float* h) {
#pragma unroll
    for (int q = 0; q < 2; ++q) tc::store_act8(Xt, Xt + PR_H_PART, Xt + 2 * PR_H_PART, r, 16 * half + q * 8, 32, h + q * 8);
  };

  const int n_arows = bt.live_agent != nullptr ? *bt.n_live : bt.A;  // agents of the unfinished instances, or all
  n_tiles = min(n_tiles, (n_arows + 127) >> 7);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int arow = tile * 128 + r;
    bool live = arow < n_arows;
    const int a = live ? (bt.live_agent != nullptr ? bt.live_agent[arow] : arow) : 0;
    if (live) live = !(bt.done != nullptr && bt.done[bt.agent_inst[a]]);
    if (!__syncthreads_or(live)) continue;
    {  // X row -> three bf16 parts (K padded 72 -> 80 with zeros); each half stages five of the ten 8-wide chunks
      const float4* xr = reinterpret_cast<const float4*>(X + (size_t)a * CTRM_DIM_X);
#pragma unroll
      for (int qq = 0; qq < 5; ++qq) {
        const int q = 5 * half + qq;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (live && q < 9) {
          const float4 p0 = __ldg(&xr[2 * q]), p1 = __ldg(&xr[2 * q + 1]);
          v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
        }
        tc::store_act8(Xt, Xt + PR_X_PART, Xt + 2 * PR_X_PART, r, q * 8, PR_K0, v);
      }
    }
    mma_round(0, sX, PR_X_PART, sW0, PR_W0_PART, PR_K0, 96, false);
    // layer 0 has completed: nothing reads the X tile any more.  Its tail receives this tile's head-layer weights while the
    // first epilogue runs (the generic-proxy writes of the X staging were fenced before the MMA that consumed them)
    if (tid == 0) tc::tma_load_image(Wi1, w_image + PR_W0_BYTES, PR_HEAD_BYTES, hb);
    float h[16];
    // ---- indicator: hidden 1 -> hidden 2 -> 3 logits -> argmax
    tc::tmem_ld16_sum2(tbase, warp, 16 * half, 96, h);
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = fmaxf(h[j] + ind_b0[16 * half + j], 0.f);
    tc::fence_before_sync();
    __syncthreads();  // the X tile is dead (layer 0 done) and every thread has issued its first TMEM reads
    store_hidden16(h);
    mma_round(192, sX, PR_H_PART, sWi1, PR_W1_PART, 32, 32, true);
    tc::tmem_ld16_sum2(tbase, warp, 192 + 16 * half, 32, h);
    int ind = 0;
    {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = fmaxf(h[j] + ind_b1[16 * half + j], 0.f);
        const float4 w = *reinterpret_cast<const float4*>(&ind_w2[(16 * half + j) * 4]);
        o0 = fmaf(v, w.x, o0);
        o1 = fmaf(v, w.y, o1);
        o2 = fmaf(v, w.z, o2);
      }
      red[(half * 128 + r) * 4 + 0] = o0;
      red[(half * 128 + r) * 4 + 1] = o1;
      red[(half * 128 + r) * 4 + 2] = o2;
      __syncthreads();
      // both halves form the same sums in the same order: identical argmax
      o0 = (ind_b2[0] + red[r * 4 + 0]) + red[(128 + r) * 4 + 0];
      o1 = (ind_b2[1] + red[r * 4 + 1]) + red[(128 + r) * 4 + 1];
      o2 = (ind_b2[2] + red[r * 4 + 2]) + red[(128 + r) * 4 + 2];
      float best = o0;
      if (o1 > best) { best = o1; ind = 1; }
      if (o2 > best) { best = o2; ind = 2; }
      if (force_ind >= 0) ind = force_ind;  // one launch per indicator value (randomize_indicator, learned_sampler.py:93-97)
      else if (live && half == 0 && ind_out != nullptr) ind_out[a] = ind;
    }
    // ---- decoder first layer, [X, onehot] part (TMEM columns 64..95 of the shared chain)
    tc::tmem_ld16_sum2(tbase, warp, 64 + 16 * half, 96, h);
    if (live) {
      float4* d0 = reinterpret_cast<float4*>(dec0_out + (size_t)a * 32 + 16 * half);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int c0 = 16 * half + j4 * 4;
        d0[j4] = make_float4(h[j4 * 4 + 0] + dec_b0[c0 + 0] + dec_oh[ind * 32 + c0 + 0],
                             h[j4 * 4 + 1] + dec_b0[c0 + 1] + dec_oh[ind * 32 + c0 + 1],
                             h[j4 * 4 + 2] + dec_b0[c0 + 2] + dec_oh[ind * 32 + c0 + 2],
                             h[j4 * 4 + 3] + dec_b0[c0 + 3] + dec_oh[ind * 32 + c0 + 3]);
      }
    }
    // ---- encoder_input: hidden 1 (+ one-hot row) -> hidden 2 -> 64 logits -> log_softmax -> sqrt of the clamped probabilities
    tc::tmem_ld16_sum2(tbase, warp, 32 + 16 * half, 96, h);
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = fmaxf(h[j] + enc_b0[16 * half + j] + enc_oh[ind * 32 + 16 * half + j], 0.f);
    store_hidden16(h);  // the indicator's hidden tile was consumed by its MMA round (waited on above)
    mma_round(192, sX, PR_H_PART, sWe1, PR_W1_PART, 32, 32, true);
    tc::tmem_ld16_sum2(tbase, warp, 192 + 16 * half, 32, h);
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = fmaxf(h[j] + enc_b1[16 * half + j], 0.f);
    store_hidden16(h);
    mma_round(0, sX, PR_H_PART, sWe2, PR_W2_PART, 32, 64, true);  // every layer-0 column has been consumed
    {
      float lg[32];  // this half's 32 of the 64 logits
      tc::tmem_ld32_sum2(tbase, warp, 32 * half, 64, lg);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        lg[j] += enc_b2[32 * half + j];
        mx = fmaxf(mx, lg[j]);
      }
      red[half * 128 + r] = mx;
      __syncthreads();
      mx = fmaxf(red[r], red[128 + r]);
      __syncthreads();
      float se = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) se += expf(lg[j] - mx);
      red[half * 128 + r] = se;
      __syncthreads();
      const float lse = logf(red[r] + red[128 + r]);
      __syncthreads();
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        lg[j] = expf((lg[j] - mx) - lse);  // probs = exp(log_softmax)
        ps += lg[j];
      }
      red[half * 128 + r] = ps;
      __syncthreads();
      ps = red[r] + red[128 + r];
      if (live) {
        // sqrt of the clamped, renormalised probabilities (see cvae_decode: for the trained temperature 2 the relaxed one-hot
        // sample is z_i = sqrt(p_i / E_i) / sum_j sqrt(p_j / E_j) with E = -log u; the K copies share sqrt(p))
        float4* lo = reinterpret_cast<float4*>(sqrtp_out + (size_t)a * 64 + 32 * half);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          lo[j4] = make_float4(sqrtf(fminf(fmaxf(lg[j4 * 4 + 0] / ps, eps), 1.f - eps)),
                               sqrtf(fminf(fmaxf(lg[j4 * 4 + 1] / ps, eps), 1.f - eps)),
                               sqrtf(fminf(fmaxf(lg[j4 * 4 + 2] / ps, eps), 1.f - eps)),
                               sqrtf(fminf(fmaxf(lg[j4 * 4 + 3] / ps, eps), 1.f - eps)));
      }
    }
    hphase ^= 1u;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, PR_TMEM_COLS);
}

size_t cvae_prior_image_bytes() { return PR_W_BYTES; }
int launch_cvae_prior_image(const float* d_blob, const ctrm_weight_offsets_t& off, uint8_t* d_image, cudaStream_t st) {
  CTRM_CUDA_OK(cudaFuncSetAttribute(cvae_prior_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_W_BYTES));
  cvae_prior_image_kernel<<<1, PR_THREADS, PR_W_BYTES, st>>>(d_blob, off, d_image);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_cvae_prior(const DevModel& m, const Batch& bt, const float* d_X, float* d_sqrtp, float* d_dec0, int32_t* d_ind,
                      cudaStream_t st, int force_ind) {
  if (bt.A <= 0) return 0;
  const int n_tiles = (bt.A + 127) / 128;
  static SmemAttrCache attr;
  CTRM_CUDA_OK(attr.ensure(cvae_prior_tc_kernel, PR_SMEM));
  const int grid = n_tiles < 2 * 148 ? n_tiles : 2 * 148;  // persistent, two CTAs per SM (114 KB of shared memory each)
  cvae_prior_tc_kernel<<<grid, PR_THREADS, PR_SMEM, st>>>(bt, d_X, m.prior_img, d_sqrtp, d_dec0, d_ind, n_tiles, force_ind);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

#define DT_THREADS 256
// shared memory: z digit tiles 4 x 16 KB (K = 64; tiles 2 and 3 double as the fp32 scratch of the softmax, and the whole
// region is reused for the three bf16 parts of h1, K = 32) | W0 digit tiles 4 x 4 KB | W1 parts 3 x 2 KB | W2, b1, b2 |
// mbarrier, TMEM slot
#define DT_A_PART 16384u
#define DT_H_PART 8192u
#define DT_W0_PART 4096u
#define DT_W1_PART 2048u
#define DT_SMEM (4 * 16384 + 4 * 4096 + 3 * 2048 + (128 + 32 + 4 + 256 + 512) * 4 + 16)
#define DT_TMEM_COLS 256  // 4 x 32 layer-0 digit accumulators + 2 x 32 layer-1 accumulator classes at 128

// 8 fp32 values of row r, columns k0..k0+7, parked as two 16-byte halves in the slots of digit tiles 2 and 3
__device__ __forceinline__ void park8(uint8_t* At, int r, int k0, const float* v) {
  const uint32_t off = tc::elem_offset(r, k0, 64);
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = __float_as_uint(v[2 * j]), b = __float_as_uint(v[2 * j + 1]);
    lo[j] = (a & 0xFFFFu) | (b << 16);
    hi[j] = (a >> 16) | (b & 0xFFFF0000u);
  }
  *reinterpret_cast<uint4*>(At + 2 * DT_A_PART + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(At + 3 * DT_A_PART + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}
__device__ __forceinline__ void unpark8(const uint8_t* At, int r, int k0, float* v) {
  const uint32_t off = tc::elem_offset(r, k0, 64);
  const uint4 l4 = *reinterpret_cast<const uint4*>(At + 2 * DT_A_PART + off);
  const uint4 h4 = *reinterpret_cast<const uint4*>(At + 3 * DT_A_PART + off);
  const uint32_t lo[4] = {l4.x, l4.y, l4.z, l4.w}, hi[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float((lo[j] & 0xFFFFu) | (hi[j] << 16));
    v[2 * j + 1] = __uint_as_float((lo[j] >> 16) | (hi[j] & 0xFFFF0000u));
  }
}

// Two threads per row: thread (row, half) owns latent columns [32 half, 32 half + 32) in the Gumbel-softmax phase and
// hidden columns [16 half, 16 half + 16) in the epilogues; warps 0-3 and 4-7 address the same TMEM lanes (lane group =
// warp % 4) at different columns.  The kernel is bound by the latency of the logf / expf chains, so twice the warps per
// tile is worth more than anything else here.
// weight operands of cvae_decode exactly as they sit in its shared memory from W0t on
#define DT_W_BYTES (4 * DT_W0_PART + 3 * DT_W1_PART + (128 + 32 + 4) * 4)
__global__ void __launch_bounds__(DT_THREADS) cvae_decode_image_kernel(const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                                     int w0_exp, uint8_t* __restrict__ image) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* W0t = smem;
  uint8_t* W1t = W0t + 4 * DT_W0_PART;
  float* W2 = reinterpret_cast<float*>(W1t + 3 * DT_W1_PART);
  float* B1 = W2 + 128;
  float* B2 = B1 + 32;
  const int tid = threadIdx.x;
  // |W0z| < 2^w0_exp (host): digit scale 2^(7 - w0_exp) puts every weight in [-128, 128]
  tc::stage_weight_digits(blob + off.dec_w0, 32, 0, 64, 32, 32, 64, exp2f((float)(7 - w0_exp)), W0t, tid, DT_THREADS);
  tc::stage_weight_split(blob + off.dec_w1, 32, 0, 32, 32, 32, 32, W1t, tid, DT_THREADS);
  for (int i = tid; i < 128; i += DT_THREADS) W2[i] = __ldg(&blob[off.dec_w2 + i]);
  if (tid < 32) B1[tid] = __ldg(&blob[off.dec_b1 + tid]);
  if (tid < 4) B2[tid] = __ldg(&blob[off.dec_b2 + tid]);
  __syncthreads();
  tc::dump_image(smem, image, DT_W_BYTES, tid, DT_THREADS);
}

__global__ void __launch_bounds__(DT_THREADS, 2) cvae_decode_tc_kernel(Batch bt, int K, const float* __restrict__ sqrtp,
                                                                       const float* __restrict__ dec0,
                                                                       const float* __restrict__ uniforms, int use_state_kt,
                                                                       int kt_fixed, int t_fixed, float temp, int w0_exp,
                                                                       const uint8_t* __restrict__ w_image,
                                                                       float* __restrict__ samples, int n_tiles, int variants,
                                                                       int k_first, int k_count, int emit_retry) {
  // This launch computes copies k_first .. k_first + k_count - 1 of every agent row (K = the agent-major stride of `samples`).
  // emit_retry: the copy is the FIRST candidate of the scan of learned_sampler.py:159-166; its validity (the valid_edge closure
  // :127-132) is tested here with the step kernels' own predicates and the agents whose first candidate fails are appended to
  // Batch::retry_agent — only those need the remaining copies (a second, small launch): 94 % of the learned agent-steps of
  // the benchmark take their first candidate.
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* At = smem;                       // z digits 1..4, [128 x 64] bf16 each
  uint8_t* W0t = At + 4 * DT_A_PART;        // W0z digits 1..4, [32 x 64]
  uint8_t* W1t = W0t + 4 * DT_W0_PART;      // W1 parts 1..3, [32 x 32]
  float* W2 = reinterpret_cast<float*>(W1t + 3 * DT_W1_PART);  // [32][4]
  float* B1 = W2 + 128;
  float* B2 = B1 + 32;
  float* red = B2 + 4;                      // [2][128] max / sum exchange between the two halves of a row
  float* red3 = red + 256;                  // [128][4] partial outputs of half 1
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red3 + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;

  if (tid == 0) {  // the weight operands arrive as one prebuilt image (cvae_decode_image_kernel), through the TMA engine
    tc::mbar_init(tc::smem_u32(mbar), 1);
    tc::tma_load_image(W0t, w_image, DT_W_BYTES, tc::smem_u32(mbar));
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, DT_TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t mb = tc::smem_u32(mbar);
  tc::mbar_wait(mb, 0);  // weights have landed; the MMA rounds continue on the same barrier with phase 1
  const uint32_t sA = tc::smem_u32(At), sW0 = tc::smem_u32(W0t), sW1 = tc::smem_u32(W1t);
  uint32_t phase = 1;
  const float eps = 1.1920928955078125e-07f;
  int temp_exp;
  const float inv_temp = (frexpf(temp, &temp_exp) == 0.5f && temp_exp > -100 && temp_exp < 100) ? 1.f / temp : 0.f;
  const bool sqrt_path = temp == 2.0f;
  const int n_arows = bt.live_agent != nullptr ? *bt.n_live : bt.A;  // agents of the unfinished instances, or all
  const long long n_rows = (long long)n_arows * k_count;
  n_tiles = (int)min((long long)n_tiles, (n_rows + 127) >> 7);
  // accumulator g (digit index sum, 0-based) carries weight 256^-(g+1) * 2^(w0_exp - 7)
  double acc_scale[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) acc_scale[g] = exp2((double)(w0_exp - 7 - 8 * (g + 1)));

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row = (long long)tile * 128 + r;
    bool live = row < n_rows;
    int a = 0, k = 0, b = 0;
    long long orow = 0;  // row of this (agent, copy) in the agent-major output / uniform tables
    if (live) {
      const int arow = (int)(row / k_count);
      k = k_first + (int)(row - (long long)arow * k_count);
      a = bt.live_agent != nullptr ? bt.live_agent[arow] : arow;
      orow = (long long)a * K + k;
      b = bt.agent_inst[a];
      live = !(bt.done != nullptr && bt.done[b]);
    }
    // indicator variant of this row (randomize_indicator, learned_sampler.py:93-97): one draw per (instance, step) decides
    // whether the step's indicators are random; then every row k*N + i draws its own, else all copies take the argmax
    size_t voff = 0;
    if (variants && live) {
      const int kt = use_state_kt ? bt.k_traj[b] : kt_fixed, tt = use_state_kt ? bt.t[b] : t_fixed;
      const uint32_t c0v = bt.inst_base + (uint32_t)b, c1v = rng_ctr1(kt, tt);
      const uint4 ws = philox4x32(c0v, c1v, 0u, rng_ctr3(STREAM_STEP, 0), bt.seed_lo, bt.seed_hi);
      const int mk = bt.makespan[b], D = mk ? mk : bt.max_T;
      int v;
      if (u64_to_f64(ws.x, ws.y) > bt.p_rw[(size_t)tt * (bt.max_T + 1) + D]) {
        const int N = bt.agent_off[b + 1] - bt.agent_off[b], il = a - bt.agent_off[b];
        v = (int)(philox4x32(c0v, c1v, (uint32_t)(k * N + il), rng_ctr3(STREAM_RANDIND, 0), bt.seed_lo, bt.seed_hi).x %
                  (uint32_t)bt.dim_ind);
      } else {
        v = bt.ind[a];
      }
      voff = (size_t)v * bt.A;
    }
    // a tile whose rows all belong to finished instances is skipped entirely
    if (!__syncthreads_or(live)) continue;
    // ---- z = RelaxedOneHotCategorical(temp, probs).rsample() (cvae.py:133-136): softmax((log p + g) / temp) with
    //      g = -log(-log u).  For temp = 2 (the trained value) exp((log p - log E) / 2) = sqrt(p / E), E = -log u, so
    //      z_i = sqrt(p_i) rsqrt(E_i) / sum_j (...): ONE accurate logarithm per latent instead of two plus an exponential, no
    //      max pass (sqrt(p / E) lies in [1e-4, 3e3]), and every factor carries <= 2 ulp — closer to the float64 value than
    //      the reference's own fp32 chain, whose log / exp round at magnitude 16.  sqrt(p) comes from cvae_prior.
    //      Any other temperature takes the literal formula below.
    float se = 0.f;
    if (sqrt_path) {
      float rv[32];
      if (live) {
        const float* sp = sqrtp + (voff + (size_t)a) * 64;
        const float* urow = nullptr;
        uint32_t c0 = 0, c1 = 0, c2 = 0;
        if (uniforms != nullptr) {
          urow = uniforms + (size_t)orow * 64;
        } else {
          const int N = bt.agent_off[b + 1] - bt.agent_off[b], il = a - bt.agent_off[b];
          const int kt = use_state_kt ? bt.k_traj[b] : kt_fixed, tt = use_state_kt ? bt.t[b] : t_fixed;
          c0 = bt.inst_base + (uint32_t)b;
          c1 = rng_ctr1(kt, tt);
          c2 = (uint32_t)(k * N + il);  // the reference's row order k*N + i (learned_sampler.py:82)
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {  // 8 x 4 latents of this half
          const int c = 32 * half + 4 * qq;
          float4 u4;
          if (urow != nullptr) {
            u4 = __ldg(reinterpret_cast<const float4*>(urow + c));
          } else {
            const uint4 w = philox4x32(c0, c1, c2, rng_ctr3(STREAM_GUMBEL, (uint32_t)(c >> 2)), bt.seed_lo, bt.seed_hi);
            u4 = make_float4(u32_to_f32(w.x), u32_to_f32(w.y), u32_to_f32(w.z), u32_to_f32(w.w));
          }
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(sp + c));
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float E = -logf(fminf(fmaxf(uu[j], eps), 1.f - eps));  // in [1.19e-7, 15.95]
            rv[4 * qq + j] = pp[j] * rsqrtf(E);
            se += rv[4 * qq + j];
          }
        }
      }
      red[half * 128 + r] = se;
      __syncthreads();
      se = red[r] + red[128 + r];
      if (live) {
        const float inv = 1.f / se;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float s8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) s8[j] = fminf(rv[8 * q + j] * inv, 1.f);
          tc::store_digits8_unit(At, DT_A_PART, r, (4 * half + q) * 8, 64, s8);
        }
      } else {
        const float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q = 4 * half; q < 4 * half + 4; ++q) tc::store_digits8_unit(At, DT_A_PART, r, q * 8, 64, v);
      }
    } else {
    float mx = -INFINITY;
    if (live) {
      const float* sp = sqrtp + (voff + (size_t)a) * 64;
      const float* urow = nullptr;
      uint32_t c0 = 0, c1 = 0, c2 = 0;
      if (uniforms != nullptr) {
        urow = uniforms + (size_t)orow * 64;
      } else {
        const int N = bt.agent_off[b + 1] - bt.agent_off[b], il = a - bt.agent_off[b];
        const int kt = use_state_kt ? bt.k_traj[b] : kt_fixed, tt = use_state_kt ? bt.t[b] : t_fixed;
        c0 = bt.inst_base + (uint32_t)b;
        c1 = rng_ctr1(kt, tt);
        c2 = (uint32_t)(k * N + il);  // the reference's row order k*N + i (learned_sampler.py:82)
      }
      for (int q = 4 * half; q < 4 * half + 4; ++q) {
        float u8[8], s[8];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          if (urow != nullptr) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(urow + q * 8 + h2 * 4));
            u8[h2 * 4 + 0] = v.x; u8[h2 * 4 + 1] = v.y; u8[h2 * 4 + 2] = v.z; u8[h2 * 4 + 3] = v.w;
          } else {
            const uint4 w = philox4x32(c0, c1, c2, rng_ctr3(STREAM_GUMBEL, (uint32_t)(q * 2 + h2)), bt.seed_lo, bt.seed_hi);
            u8[h2 * 4 + 0] = u32_to_f32(w.x); u8[h2 * 4 + 1] = u32_to_f32(w.y);
            u8[h2 * 4 + 2] = u32_to_f32(w.z); u8[h2 * 4 + 3] = u32_to_f32(w.w);
          }
        }
        const float4 la = __ldg(reinterpret_cast<const float4*>(sp + q * 8)), lb = __ldg(reinterpret_cast<const float4*>(sp + q * 8 + 4));
        // logits = log(clamp(p)) = 2 log(sqrt(clamp(p)))  (torch.distributions probs_to_logits)
        const float l[8] = {2.f * logf(la.x), 2.f * logf(la.y), 2.f * logf(la.z), 2.f * logf(la.w),
                            2.f * logf(lb.x), 2.f * logf(lb.y), 2.f * logf(lb.z), 2.f * logf(lb.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float u = fminf(fmaxf(u8[j], eps), 1.f - eps);
          const float g = -logf(-logf(u));
          // torch divides by the temperature; for a power of two (the trained 2.0) the reciprocal multiply is the same
          // float and saves an IEEE division per latent
          s[j] = inv_temp != 0.f ? (l[j] + g) * inv_temp : (l[j] + g) / temp;
          mx = fmaxf(mx, s[j]);
        }
        park8(At, r, q * 8, s);
      }
    }
    red[half * 128 + r] = mx;
    __syncthreads();
    mx = fmaxf(red[r], red[128 + r]);
    __syncthreads();
    // z = softmax(s): e = exp(s - max) is computed once and parked; z = e / sum(e) (torch evaluates
    // exp(s - logsumexp(s)); the two agree to an ulp and this one saves 64 expf per row)
    if (live) {
      for (int q = 4 * half; q < 4 * half + 4; ++q) {
        float s[8];
        unpark8(At, r, q * 8, s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] = expf(s[j] - mx);
          se += s[j];
        }
        park8(At, r, q * 8, s);
      }
    }
    red[half * 128 + r] = se;
    __syncthreads();
    se = red[r] + red[128 + r];
    if (live) {
      const float inv = 1.f / se;
      for (int q = 4 * half; q < 4 * half + 4; ++q) {
        float s[8];
        unpark8(At, r, q * 8, s);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = fminf(s[j] * inv, 1.f);
        tc::store_digits8_unit(At, DT_A_PART, r, q * 8, 64, s);
      }
    } else {
      const float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int q = 4 * half; q < 4 * half + 4; ++q) tc::store_digits8_unit(At, DT_A_PART, r, q * 8, 64, v);
    }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::mma_digits4(tbase, 32, sA, DT_A_PART, sW0, DT_W0_PART, 64, 32);
      tc::commit(mb);
    }
    tc::mbar_wait(mb, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    // combine the four exact digit accumulators in fp64, add the x-part / bias, round once to fp32 (16 columns per thread)
    float h[16];
    {
      double pre[16];
      const float4* d0 = reinterpret_cast<const float4*>(dec0 + (voff + (size_t)a) * 32 + 16 * half);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 v = live ? __ldg(&d0[j4]) : make_float4(0.f, 0.f, 0.f, 0.f);
        pre[j4 * 4 + 0] = (double)v.x; pre[j4 * 4 + 1] = (double)v.y; pre[j4 * 4 + 2] = (double)v.z; pre[j4 * 4 + 3] = (double)v.w;
      }
#pragma unroll
      for (int g = 3; g >= 0; --g) {
        tc::tmem_ld16(tbase, warp, 32 * g + 16 * half, h);
#pragma unroll
        for (int j = 0; j < 16; ++j) pre[j] = fma((double)h[j], acc_scale[g], pre[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = fmaxf((float)pre[j], 0.f);
    }
    // h1 becomes the A operand of layer 1 (K = 32), reusing the digit tiles (layer 0 has completed)
    tc::fence_before_sync();
    __syncthreads();  // every thread has read its layer-0 accumulators before the tile memory is rewritten
#pragma unroll
    for (int q = 0; q < 2; ++q) tc::store_act8(At, At + DT_H_PART, At + 2 * DT_H_PART, r, 16 * half + q * 8, 32, h + q * 8);
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::mma_bf16x3_c2(tbase + 128, 32, sA, DT_H_PART, sW1, DT_W1_PART, 32, 32);
      tc::commit(mb);
    }
    tc::mbar_wait(mb, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    tc::tmem_ld16_sum2(tbase, warp, 128 + 16 * half, 32, h);
    {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = fmaxf(h[j] + B1[16 * half + j], 0.f);
        const float4 w = *reinterpret_cast<const float4*>(&W2[(16 * half + j) * 4]);
        o0 = fmaf(v, w.x, o0);
        o1 = fmaf(v, w.y, o1);
        o2 = fmaf(v, w.z, o2);
      }
      if (half == 1) {
        red3[r * 4 + 0] = o0;
        red3[r * 4 + 1] = o1;
        red3[r * 4 + 2] = o2;
      }
      tc::fence_before_sync();
      __syncthreads();  // also: every TMEM / shared-memory read of this tile is done before the next tile overwrites them
      tc::fence_after_sync();
      if (half == 0 && live) {
        float* out = samples + (size_t)orow * 3;
        const float s0 = (B2[0] + o0) + red3[r * 4 + 0], s1 = (B2[1] + o1) + red3[r * 4 + 1], s2 = (B2[2] + o2) + red3[r * 4 + 2];
        out[0] = s0;
        out[1] = s1;
        out[2] = s2;
        if (emit_retry) {
          // valid_edge(cur, Y) exactly as select evaluates it (rollout.cu): Y = cur + f32(dx * mag), bounds, speed, obstacles
          const double2 cp = *reinterpret_cast<const double2*>(bt.cur + 2 * (size_t)a);
          const double2 q1 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 2);  // speed, speed^2
          const double rad = bt.agent_pack[6 * (size_t)a + 5];
          const double x = f64add(cp.x, (double)__fmul_rn(s0, s2)), y = f64add(cp.y, (double)__fmul_rn(s1, s2));
          bool ok = bounds_ok(x, y, rad) && speed_ok(cp.x, cp.y, x, y, q1.x, q1.y);
          if (ok) {
            const int o0b = bt.obs_off[b], o1b = bt.obs_off[b + 1];
            ok = !any_obstacle_hit_f4(cp.x, cp.y, x, y, rad, bt.obs_xyr, bt.obs_f4, o0b, o1b);
          }
          if (!ok) {
            const int slot = atomicAdd(bt.n_retry, 1);
            bt.retry_agent[slot] = a;
            bt.retry_inst[slot] = b;
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, DT_TMEM_COLS);
}

size_t cvae_decode_image_bytes() { return DT_W_BYTES; }
int launch_cvae_decode_image(const float* d_blob, const ctrm_weight_offsets_t& off, int w0_exp, uint8_t* d_image, cudaStream_t st) {
  cvae_decode_image_kernel<<<1, DT_THREADS, DT_W_BYTES, st>>>(d_blob, off, w0_exp, d_image);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_cvae_decode(const DevModel& m, const Batch& bt, const float* d_sqrtp, const float* d_dec0, const float* d_uniforms,
                       int use_state_kt, int kt_fixed, int t_fixed, float* d_samples, cudaStream_t st, int variants, int k_first,
                       int k_count, int emit_retry) {
  if (bt.A <= 0) return 0;
  const int K = bt.K;
  if (k_count < 0) k_count = K - k_first;  // default: all remaining copies
  if (k_count <= 0) return 0;
  const long long rows = (long long)bt.A * k_count;
  const int n_tiles = (int)((rows + 127) / 128);
  static SmemAttrCache attr;
  CTRM_CUDA_OK(attr.ensure(cvae_decode_tc_kernel, DT_SMEM));
  const int grid = n_tiles < 2 * 148 ? n_tiles : 2 * 148;  // persistent: two CTAs per SM
  cvae_decode_tc_kernel<<<grid, DT_THREADS, DT_SMEM, st>>>(bt, K, d_sqrtp, d_dec0, d_uniforms, use_state_kt, kt_fixed, t_fixed,
                                                          m.desc.temp, m.dec_w0z_exp, m.decode_img, d_samples, n_tiles, variants, k_first, k_count,
                                                          emit_retry);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
