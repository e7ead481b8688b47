// Indicator + CVAE sampling.
//   reference: indicator/argmax/one_hot (roadmap_learned/learned_sampler.py:99-104), CTRMNet.sample
//   (learning/model/cvae.py:128-138) = encoder_input (cvae.py:57-68) -> RelaxedOneHotCategorical(temp).rsample()
//   (torch.distributions: probs normalised, logits = log(clamp(probs)), Gumbel noise from clamped uniforms,
//   softmax((logits+g)/temp)) -> decoder (cvae.py:72-76).  BatchNorm runs in eval mode (model/__init__.py:53) and
//   is folded into the preceding Linear by the host.
// Two kernels: cvae_prior (one thread per agent: the parts shared by the K copies) and cvae_decode (one thread
// per (agent, copy) row: Gumbel-softmax + decoder MLP).
#include "common.cuh"
#include "kernels.h"

#define CV_T 64  // threads (rows) per block

template <int OUT, int LDW>
__device__ __forceinline__ void fma_row(float xv, const float* __restrict__ wrow, float* acc) {
#pragma unroll
  for (int j4 = 0; j4 < (OUT + 3) / 4; ++j4) {
    float4 w = *reinterpret_cast<const float4*>(&wrow[j4 * 4]);
    acc[j4 * 4 + 0] = fmaf(xv, w.x, acc[j4 * 4 + 0]);
    if (j4 * 4 + 1 < OUT) acc[j4 * 4 + 1] = fmaf(xv, w.y, acc[j4 * 4 + 1]);
    if (j4 * 4 + 2 < OUT) acc[j4 * 4 + 2] = fmaf(xv, w.z, acc[j4 * 4 + 2]);
    if (j4 * 4 + 3 < OUT) acc[j4 * 4 + 3] = fmaf(xv, w.w, acc[j4 * 4 + 3]);
  }
}

// acc[OUT] = b + sum_k relu?(h[k]) * W[k][:]   (h in registers, fully unrolled)
template <int OUT, int LDW>
__device__ __forceinline__ void dense32(const float* h, const float* __restrict__ Ws, const float* __restrict__ bs, float* acc) {
#pragma unroll
  for (int j = 0; j < OUT; ++j) acc[j] = bs[j];
#pragma unroll
  for (int k = 0; k < 32; ++k) fma_row<OUT, LDW>(h[k], Ws + k * LDW, acc);
}

// ------------------------------------------------------------------------------------------------
// prior: indicator (72->32->32->3, argmax), encoder_input ([X, onehot] 75->32->32->64, log_softmax),
//        logits = log(clamp(p / sum p, eps, 1-eps)), and the X/indicator part of the decoder's first layer.
// outputs per agent: logits[64], dec0[32], ind
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_T) cvae_prior_kernel(Batch bt, const float* __restrict__ X, const float* __restrict__ blob,
                                                          ctrm_weight_offsets_t off, int w_lo, int w_hi,
                                                          float* __restrict__ logits_out, float* __restrict__ dec0_out,
                                                          int32_t* __restrict__ ind_out, const int32_t* __restrict__ ind_override) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                    // blob[w_lo, w_hi)
  float* xs = smem + (w_hi - w_lo);    // [72][CV_T]
  for (int i = threadIdx.x; i < w_hi - w_lo; i += blockDim.x) Ws[i] = __ldg(&blob[w_lo + i]);
  const int a0 = blockIdx.x * CV_T;
  const int nrow = min(CV_T, bt.A - a0);
  for (int i = threadIdx.x; i < nrow * CTRM_DIM_X; i += blockDim.x) {
    int r = i / CTRM_DIM_X, k = i - r * CTRM_DIM_X;
    xs[k * CV_T + r] = X[(size_t)(a0 + r) * CTRM_DIM_X + k];
  }
  __syncthreads();
  const int tid = threadIdx.x, a = a0 + tid;
  if (a >= bt.A) return;
  if (bt.done != nullptr && bt.done[bt.agent_inst[a]]) return;
#define WP(o) (Ws + ((o)-w_lo))
  float h1[32], h2[32];
  // ---- indicator
  int ind;
  {
#pragma unroll
    for (int j = 0; j < 32; ++j) h1[j] = WP(off.ind_b0)[j];
    for (int k = 0; k < CTRM_DIM_X; ++k) fma_row<32, 32>(xs[k * CV_T + tid], WP(off.ind_w0) + k * 32, h1);
#pragma unroll
    for (int j = 0; j < 32; ++j) h1[j] = fmaxf(h1[j], 0.f);
    dense32<32, 32>(h1, WP(off.ind_w1), WP(off.ind_b1), h2);
#pragma unroll
    for (int j = 0; j < 32; ++j) h2[j] = fmaxf(h2[j], 0.f);
    float o[4];
    dense32<3, 4>(h2, WP(off.ind_w2), WP(off.ind_b2), o);
    ind = 0;
    float best = o[0];
    for (int j = 1; j < bt.dim_ind && j < 3; ++j)
      if (o[j] > best) { best = o[j]; ind = j; }
    if (ind_override != nullptr && ind_override[a] >= 0) ind = ind_override[a];
    if (ind_out != nullptr) ind_out[a] = ind;
  }
  // ---- encoder_input on [X, onehot(ind)]
  {
#pragma unroll
    for (int j = 0; j < 32; ++j) h1[j] = WP(off.enc_b0)[j] + WP(off.enc_w0)[(CTRM_DIM_X + ind) * 32 + j];
    for (int k = 0; k < CTRM_DIM_X; ++k) fma_row<32, 32>(xs[k * CV_T + tid], WP(off.enc_w0) + k * 32, h1);
#pragma unroll
    for (int j = 0; j < 32; ++j) h1[j] = fmaxf(h1[j], 0.f);
    dense32<32, 32>(h1, WP(off.enc_w1), WP(off.enc_b1), h2);
#pragma unroll
    for (int j = 0; j < 32; ++j) h2[j] = fmaxf(h2[j], 0.f);
    float lg[64];
    dense32<64, 64>(h2, WP(off.enc_w2), WP(off.enc_b2), lg);
    // log_softmax (cvae.py:67 LogSoftmax) then probs = exp(.) (cvae.py:132), Categorical normalisation,
    // logits = log(clamp(probs, eps, 1 - eps))   (torch.distributions.utils.probs_to_logits)
    float mx = lg[0];
#pragma unroll
    for (int j = 1; j < 64; ++j) mx = fmaxf(mx, lg[j]);
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) se += expf(lg[j] - mx);
    float lse = logf(se);
    float ps = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      lg[j] = expf((lg[j] - mx) - lse);
      ps += lg[j];
    }
    const float eps = 1.1920928955078125e-07f;
    float* lo = logits_out + (size_t)a * 64;
#pragma unroll
    for (int j4 = 0; j4 < 16; ++j4) {
      float4 v;
      v.x = logf(fminf(fmaxf(lg[j4 * 4 + 0] / ps, eps), 1.f - eps));
      v.y = logf(fminf(fmaxf(lg[j4 * 4 + 1] / ps, eps), 1.f - eps));
      v.z = logf(fminf(fmaxf(lg[j4 * 4 + 2] / ps, eps), 1.f - eps));
      v.w = logf(fminf(fmaxf(lg[j4 * 4 + 3] / ps, eps), 1.f - eps));
      *reinterpret_cast<float4*>(lo + j4 * 4) = v;
    }
  }
  // ---- decoder first layer, the [X, onehot] rows 64..138 of W0 (input = [z 64 | X 72 | ind 3], cvae.py:137)
  {
#pragma unroll
    for (int j = 0; j < 32; ++j) h1[j] = WP(off.dec_b0)[j] + WP(off.dec_w0)[(64 + CTRM_DIM_X + ind) * 32 + j];
    for (int k = 0; k < CTRM_DIM_X; ++k) fma_row<32, 32>(xs[k * CV_T + tid], WP(off.dec_w0) + (64 + k) * 32, h1);
    float* d0 = dec0_out + (size_t)a * 32;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
      *reinterpret_cast<float4*>(d0 + j4 * 4) = make_float4(h1[j4 * 4], h1[j4 * 4 + 1], h1[j4 * 4 + 2], h1[j4 * 4 + 3]);
  }
#undef WP
}

// ------------------------------------------------------------------------------------------------
// decode: one thread per (agent a, copy k).  u -> Gumbel -> z = softmax((logits + g)/temp) -> decoder.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_T) cvae_decode_kernel(Batch bt, int K, const float* __restrict__ logits,
                                                           const float* __restrict__ dec0, const float* __restrict__ uniforms,
                                                           int use_state_kt, int kt_fixed, int t_fixed, float temp, const float* __restrict__ blob,
                                                           ctrm_weight_offsets_t off, float* __restrict__ samples) {
  extern __shared__ __align__(16) float smem[];
  float* W0 = smem;               // [64][32] z rows of the decoder's first layer
  float* W1 = W0 + 64 * 32;       // [32][32]
  float* W2 = W1 + 32 * 32;       // [32][4]
  float* B1 = W2 + 32 * 4;        // [32]
  float* B2 = B1 + 32;            // [4]
  float* zs = B2 + 4;             // [64][CV_T]
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) W0[i] = __ldg(&blob[off.dec_w0 + i]);
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) W1[i] = __ldg(&blob[off.dec_w1 + i]);
  for (int i = threadIdx.x; i < 32 * 4; i += blockDim.x) W2[i] = __ldg(&blob[off.dec_w2 + i]);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) B1[i] = __ldg(&blob[off.dec_b1 + i]);
  for (int i = threadIdx.x; i < 4; i += blockDim.x) B2[i] = __ldg(&blob[off.dec_b2 + i]);
  __syncthreads();
  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * CV_T + tid;
  if (row >= (long long)bt.A * K) return;
  const int a = (int)(row / K), k = (int)(row - (long long)a * K);
  const int b = bt.agent_inst[a];
  if (bt.done != nullptr && bt.done[b]) return;
  const float eps = 1.1920928955078125e-07f;
  const float* lg = logits + (size_t)a * 64;
  // uniforms: explicit [A][K][64] or Philox stream 8 keyed (instance, k_traj|t, reference row k*N+i, column/4)
  const float* urow = nullptr;
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  if (uniforms != nullptr) {
    urow = uniforms + (size_t)row * 64;
  } else {
    int N = bt.agent_off[b + 1] - bt.agent_off[b], il = a - bt.agent_off[b];
    int kt = use_state_kt ? bt.k_traj[b] : kt_fixed, tt = use_state_kt ? bt.t[b] : t_fixed;
    c0 = bt.inst_base + (uint32_t)b;
    c1 = rng_ctr1(kt, tt);
    c2 = (uint32_t)(k * N + il);
  }
  float mx = -INFINITY;
#pragma unroll 4
  for (int q = 0; q < 16; ++q) {
    float u4[4];
    if (urow != nullptr) {
      float4 v = __ldg(reinterpret_cast<const float4*>(urow + q * 4));
      u4[0] = v.x; u4[1] = v.y; u4[2] = v.z; u4[3] = v.w;
    } else {
      uint4 w = philox4x32(c0, c1, c2, rng_ctr3(STREAM_GUMBEL, (uint32_t)q), bt.seed_lo, bt.seed_hi);
      u4[0] = u32_to_f32(w.x); u4[1] = u32_to_f32(w.y); u4[2] = u32_to_f32(w.z); u4[3] = u32_to_f32(w.w);
    }
    float4 l4 = __ldg(reinterpret_cast<const float4*>(lg + q * 4));
    float l[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float u = fminf(fmaxf(u4[j], eps), 1.f - eps);
      float g = -logf(-logf(u));
      float s = (l[j] + g) / temp;
      zs[(q * 4 + j) * CV_T + tid] = s;
      mx = fmaxf(mx, s);
    }
  }
  float se = 0.f;
  for (int c = 0; c < 64; ++c) se += expf(zs[c * CV_T + tid] - mx);
  const float lse = mx + logf(se);
  float h1[32], h2[32];
  {
    const float4* d0 = reinterpret_cast<const float4*>(dec0 + (size_t)a * 32);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 v = __ldg(&d0[j4]);
      h1[j4 * 4] = v.x; h1[j4 * 4 + 1] = v.y; h1[j4 * 4 + 2] = v.z; h1[j4 * 4 + 3] = v.w;
    }
  }
  for (int c = 0; c < 64; ++c) {
    float z = expf(zs[c * CV_T + tid] - lse);
    fma_row<32, 32>(z, W0 + c * 32, h1);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) h1[j] = fmaxf(h1[j], 0.f);
  dense32<32, 32>(h1, W1, B1, h2);
#pragma unroll
  for (int j = 0; j < 32; ++j) h2[j] = fmaxf(h2[j], 0.f);
  float o[4];
  dense32<3, 4>(h2, W2, B2, o);
  float* out = samples + (size_t)row * 3;
  out[0] = o[0];
  out[1] = o[1];
  out[2] = o[2];
}

// scratch owned by the caller: logits [A][64], dec0 [A][32]
int launch_cvae_prior(const DevModel& m, const Batch& bt, const float* d_X, float* d_logits, float* d_dec0, int32_t* d_ind,
                      cudaStream_t st) {
  if (bt.A <= 0) return 0;
  int w_lo = m.off.ind_w0, w_hi = m.off.total;
  size_t smem1 = (size_t)(w_hi - w_lo + CTRM_DIM_X * CV_T) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    CTRM_CUDA_OK(cudaFuncSetAttribute(cvae_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    attr = true;
  }
  cvae_prior_kernel<<<(bt.A + CV_T - 1) / CV_T, CV_T, smem1, st>>>(bt, d_X, m.blob, m.off, w_lo, w_hi, d_logits, d_dec0, d_ind,
                                                                  nullptr);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_cvae_decode(const DevModel& m, const Batch& bt, const float* d_logits, const float* d_dec0, const float* d_uniforms,
                       int use_state_kt, int kt_fixed, int t_fixed, float* d_samples, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  const int K = bt.K;
  size_t smem2 = (size_t)(64 * 32 + 32 * 32 + 32 * 4 + 32 + 4 + 64 * CV_T) * sizeof(float);
  long long rows = (long long)bt.A * K;
  cvae_decode_kernel<<<(unsigned)((rows + CV_T - 1) / CV_T), CV_T, smem2, st>>>(bt, K, d_logits, d_dec0, d_uniforms, use_state_kt,
                                                                              kt_fixed, t_fixed, m.desc.temp, m.blob, m.off, d_samples);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
