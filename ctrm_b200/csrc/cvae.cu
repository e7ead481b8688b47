// Indicator + CVAE sampling.
//   reference: indicator/argmax/one_hot (roadmap_learned/learned_sampler.py:99-104), CTRMNet.sample
//   (learning/model/cvae.py:128-138) = encoder_input (cvae.py:57-68) -> RelaxedOneHotCategorical(temp).rsample()
//   (torch.distributions: probs normalised, logits = log(clamp(probs)), Gumbel noise from clamped uniforms,
//   softmax((logits+g)/temp)) -> decoder (cvae.py:72-76).  BatchNorm runs in eval mode (model/__init__.py:53) and
//   is folded into the preceding Linear by the host.
// cvae_prior: one thread per agent, the parts shared by the K copies.  The per-copy Gumbel-softmax + decoder MLP runs on
// the tensor cores: cvae_tc.cu.
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
