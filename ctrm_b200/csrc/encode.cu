// Feature encoders of Format2D_CTRM_Input.get_feature_one_step (learning/formats/format_input.py:279-357):
//   fov_encode : FOVEncoder.forward x2 (learning/model/fov_encoder.py:34-59; format_input.py:311-312) as one
//                [A x 2F^2] x [2F^2 x 64] GEMM (self | vain first layers side by side) with the two 32->32->32
//                tails and the fov_vain part of the AgentEncoder's first layer fused into the epilogue.
//   agent_comm : numba get_arr_others_info (learning/compiled.py:30-97), AgentEncoder.forward
//                (learning/model/agent_encoder.py:34-65; format_input.py:328-334), get_comm
//                (format_input.py:212-277), get_self_info (compiled.py:100-125) and the final concat
//                (format_input.py:354-355), one warp per agent i, never materialising the N^2 tensors.
// fp32 with FMA; the summation order differs from ATen's blocked sgemm, parity is 1e-5 relative (BASELINE.json).
#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// FOV encoders
// ------------------------------------------------------------------------------------------------
#define FE_BM 64
#define FE_BN 64
#define FE_BK 32
#define FE_ALD 36  // As[row][k] leading dim
#define FE_HLD 65

__global__ void __launch_bounds__(256) fov_encode_kernel(const float* __restrict__ fov, const int32_t* __restrict__ agent_inst,
                                                         const uint8_t* __restrict__ skip_inst, int A, int KD,
                                                         const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                         float* __restrict__ e_self, float* __restrict__ e_vain,
                                                         float* __restrict__ vain_proj) {
  __shared__ __align__(16) float ABs[FE_BM * FE_ALD + FE_BK * FE_BN];  // GEMM tiles; reused as H2s by the tails
  __shared__ float Hs[FE_BM * FE_HLD];
  float* As = ABs;
  float* Bs = ABs + FE_BM * FE_ALD;
  float* H2s = ABs;
  static_assert(FE_BM * FE_HLD <= FE_BM * FE_ALD + FE_BK * FE_BN, "H2s must fit in the tile buffer");
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * FE_BM;
  // skip tiles whose agents all belong to finished instances
  if (skip_inst != nullptr) {
    int r = row0 + (tid & 63);
    bool live = (r < A) && (skip_inst[agent_inst[r]] == 0);
    if (!__syncthreads_or(live)) return;
  }
  const float* W0 = blob + off.fov_w0;
  const int tx = tid & 15, ty = tid >> 4;
  // global->register staging indices
  const int a_row = tid >> 2, a_kq = tid & 3;       // A: row a_row, k = a_kq*8 .. +7
  const int b_k = tid >> 3, b_c = (tid & 7) * 8;    // B: k = b_k, cols b_c .. +7
  const bool a_in = (row0 + a_row) < A;
  const float* a_ptr = fov + (size_t)(row0 + a_row) * KD + a_kq * 8;
  float2 ra[4];
  float4 rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + a_kq * 8 + 2 * j;
      ra[j] = (a_in && k + 1 < KD) ? *reinterpret_cast<const float2*>(a_ptr + k0 + 2 * j)
                                   : make_float2((a_in && k < KD) ? a_ptr[k0 + 2 * j] : 0.f, 0.f);
    }
    int kb = k0 + b_k;
    if (kb < KD) {
      rb[0] = __ldg(reinterpret_cast<const float4*>(W0 + (size_t)kb * FE_BN + b_c));
      rb[1] = __ldg(reinterpret_cast<const float4*>(W0 + (size_t)kb * FE_BN + b_c + 4));
    } else {
      rb[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[1] = rb[0];
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(&As[a_row * FE_ALD + a_kq * 8 + 2 * j]) = ra[j];
    *reinterpret_cast<float4*>(&Bs[b_k * FE_BN + b_c]) = rb[0];
    *reinterpret_cast<float4*>(&Bs[b_k * FE_BN + b_c + 4]) = rb[1];
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  load_tiles(0);
  for (int k0 = 0; k0 < KD; k0 += FE_BK) {
    store_tiles();
    __syncthreads();
    if (k0 + FE_BK < KD) load_tiles(k0 + FE_BK);
#pragma unroll
    for (int k = 0; k < FE_BK; ++k) {
      float4 b = *reinterpret_cast<const float4*>(&Bs[k * FE_BN + tx * 4]);
      float a0 = As[(ty * 4 + 0) * FE_ALD + k], a1 = As[(ty * 4 + 1) * FE_ALD + k];
      float a2 = As[(ty * 4 + 2) * FE_ALD + k], a3 = As[(ty * 4 + 3) * FE_ALD + k];
      acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
      acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
      acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
      acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
      acc[2][0] = fmaf(a2, b.x, acc[2][0]); acc[2][1] = fmaf(a2, b.y, acc[2][1]);
      acc[2][2] = fmaf(a2, b.z, acc[2][2]); acc[2][3] = fmaf(a2, b.w, acc[2][3]);
      acc[3][0] = fmaf(a3, b.x, acc[3][0]); acc[3][1] = fmaf(a3, b.y, acc[3][1]);
      acc[3][2] = fmaf(a3, b.z, acc[3][2]); acc[3][3] = fmaf(a3, b.w, acc[3][3]);
    }
    __syncthreads();
  }
  // epilogue 1: + bias, ReLU -> Hs[row][64]
  {
    const float* b0 = blob + off.fov_b0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        Hs[(ty * 4 + i) * FE_HLD + tx * 4 + j] = fmaxf(acc[i][j] + __ldg(&b0[tx * 4 + j]), 0.f);
  }
  __syncthreads();
  // tails: thread -> (row, encoder e, 16-output half)
  const int row = tid >> 2, q = tid & 3, e = q >> 1, half = q & 1;
  auto dense16 = [&](const float* in_s, const float* W, const float* bias, float* o16, bool relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o16[j] = __ldg(&bias[half * 16 + j]);
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      float x = in_s[row * FE_HLD + e * 32 + k];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        float4 w = __ldg(reinterpret_cast<const float4*>(W + k * 32 + half * 16 + j4 * 4));
        o16[j4 * 4 + 0] = fmaf(x, w.x, o16[j4 * 4 + 0]);
        o16[j4 * 4 + 1] = fmaf(x, w.y, o16[j4 * 4 + 1]);
        o16[j4 * 4 + 2] = fmaf(x, w.z, o16[j4 * 4 + 2]);
        o16[j4 * 4 + 3] = fmaf(x, w.w, o16[j4 * 4 + 3]);
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) o16[j] = fmaxf(o16[j], 0.f);
    }
  };
  float o16[16];
  dense16(Hs, blob + (e ? off.fovv_w1 : off.fovs_w1), blob + (e ? off.fovv_b1 : off.fovs_b1), o16, true);
#pragma unroll
  for (int j = 0; j < 16; ++j) H2s[row * FE_HLD + e * 32 + half * 16 + j] = o16[j];
  __syncthreads();
  dense16(H2s, blob + (e ? off.fovv_w2 : off.fovs_w2), blob + (e ? off.fovv_b2 : off.fovs_b2), o16, false);
  const int grow = row0 + row;
  if (grow < A) {
    float* dst = (e ? e_vain : e_self);
    if (dst != nullptr) {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        *reinterpret_cast<float4*>(dst + (size_t)grow * 32 + half * 16 + j4 * 4) =
            make_float4(o16[j4 * 4], o16[j4 * 4 + 1], o16[j4 * 4 + 2], o16[j4 * 4 + 3]);
    }
  }
  // fov_vain part of AgentEncoder layer 0 (agent_encoder.py:57-60 input = [others_info | fov_vain]):
  // vain_proj[j] = b0 + W0[11:43]^T e_vain, shared by every pair (i, j)
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; ++j) Hs[row * FE_HLD + e * 32 + half * 16 + j] = o16[j];
  __syncthreads();
  if (e == 1) {
    dense16(Hs, blob + off.ag_w0v, blob + off.ag_b0, o16, false);
    if (grow < A) {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        *reinterpret_cast<float4*>(vain_proj + (size_t)grow * 32 + half * 16 + j4 * 4) =
            make_float4(o16[j4 * 4], o16[j4 * 4 + 1], o16[j4 * 4 + 2], o16[j4 * 4 + 3]);
    }
  }
}

int launch_fov_encode(const DevModel& m, const float* d_fov, const int32_t* d_agent_inst, const uint8_t* d_skip_inst, int A,
                      float* d_e_self, float* d_e_vain, float* d_vain_proj, cudaStream_t st) {
  if (A <= 0) return 0;
  int KD = 2 * m.desc.fov_size * m.desc.fov_size;
  fov_encode_kernel<<<(A + FE_BM - 1) / FE_BM, 256, 0, st>>>(d_fov, d_agent_inst, d_skip_inst, A, KD, m.blob, m.off, d_e_self,
                                                            d_e_vain, d_vain_proj);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pairwise features + agent encoder + neighbour attention, one warp per agent i
// ------------------------------------------------------------------------------------------------
#define AC_WARPS 4
#define AC_ROW 45  // per-pair smem row: msg 32 | att 10 | dist 1, padded to an odd stride

// get_normed_vec_mag (compiled.py:11-27) in float64, cast to float32 like `.float()` (format_input.py:327)
__device__ __forceinline__ void nvm(double vx, double vy, float* o) {
  double mag = sqrt(f64add(f64mul(vx, vx), f64mul(vy, vy)));
  double d = (mag == 0.0) ? 1.0 : mag;
  o[0] = (float)(vx / d);
  o[1] = (float)(vy / d);
  o[2] = (float)mag;
}

template <int IN, int OUT, int LDW>
__device__ __forceinline__ void dense_reg(const float* x, const float* __restrict__ Ws, float* acc) {
  // acc[OUT] += x[IN] . Ws[IN][LDW]; Ws in shared memory (warp-broadcast float4 reads)
#pragma unroll
  for (int k = 0; k < IN; ++k) {
    float xv = x[k];
#pragma unroll
    for (int j4 = 0; j4 < (OUT + 3) / 4; ++j4) {
      float4 w = *reinterpret_cast<const float4*>(&Ws[k * LDW + j4 * 4]);
      acc[j4 * 4 + 0] = fmaf(xv, w.x, acc[j4 * 4 + 0]);
      if (j4 * 4 + 1 < OUT) acc[j4 * 4 + 1] = fmaf(xv, w.y, acc[j4 * 4 + 1]);
      if (j4 * 4 + 2 < OUT) acc[j4 * 4 + 2] = fmaf(xv, w.z, acc[j4 * 4 + 2]);
      if (j4 * 4 + 3 < OUT) acc[j4 * 4 + 3] = fmaf(xv, w.w, acc[j4 * 4 + 3]);
    }
  }
}

__global__ void __launch_bounds__(AC_WARPS * 32) agent_comm_kernel(Batch bt, const double* __restrict__ cur,
                                                                   const double* __restrict__ prev,
                                                                   const float* __restrict__ e_self,
                                                                   const float* __restrict__ vain_proj,
                                                                   const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                                   int num_neighbors, int use_k_neighbor, int include_self,
                                                                   int n_max, float* __restrict__ X) {
  extern __shared__ __align__(16) float smem[];
  // weights: w0i [11][32] | w1 [32][32] | w2 [32][44] | b1 [32] | b2 [44]
  float* Ww0 = smem;
  float* Ww1 = Ww0 + 11 * 32;
  float* Ww2 = Ww1 + 32 * 32;
  float* Wb1 = Ww2 + 32 * 44;
  float* Wb2 = Wb1 + 32;
  float* pair_base = Wb2 + 44;  // [AC_WARPS][n_max][AC_ROW]
  int* order_base = reinterpret_cast<int*>(pair_base + (size_t)AC_WARPS * n_max * AC_ROW);  // [AC_WARPS][n_max]
  for (int i = threadIdx.x; i < 11 * 32; i += blockDim.x) Ww0[i] = __ldg(&blob[off.ag_w0i + i]);
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) Ww1[i] = __ldg(&blob[off.ag_w1 + i]);
  for (int i = threadIdx.x; i < 32 * 44; i += blockDim.x) Ww2[i] = __ldg(&blob[off.ag_w2 + i]);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) Wb1[i] = __ldg(&blob[off.ag_b1 + i]);
  for (int i = threadIdx.x; i < 44; i += blockDim.x) Wb2[i] = __ldg(&blob[off.ag_b2 + i]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a = blockIdx.x * AC_WARPS + warp;
  if (a >= bt.A) return;
  const int b = bt.agent_inst[a];
  if (bt.done != nullptr && bt.done[b]) return;
  const int a0 = bt.agent_off[b], N = bt.agent_off[b + 1] - a0, il = a - a0;
  float* prow = pair_base + (size_t)warp * n_max * AC_ROW;
  int* order = order_base + (size_t)warp * n_max;
  const double cix = cur[2 * a], ciy = cur[2 * a + 1];
  for (int j = lane; j < N; j += 32) {
    const int aj = a0 + j;
    float info[CTRM_DIM_INFO];
    nvm(f64sub(cur[2 * aj], cix), f64sub(cur[2 * aj + 1], ciy), info + 0);
    nvm(f64sub(bt.goals[2 * aj], cix), f64sub(bt.goals[2 * aj + 1], ciy), info + 3);
    nvm(f64sub(prev[2 * aj], cix), f64sub(prev[2 * aj + 1], ciy), info + 6);
    info[9] = (float)bt.rads[aj];
    info[10] = (float)bt.speeds[aj];
    float h1[32], h2[32], o[44];
    {
      const float4* vp = reinterpret_cast<const float4*>(vain_proj + (size_t)aj * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = __ldg(&vp[q]);
        h1[q * 4] = v.x; h1[q * 4 + 1] = v.y; h1[q * 4 + 2] = v.z; h1[q * 4 + 3] = v.w;
      }
    }
    dense_reg<CTRM_DIM_INFO, 32, 32>(info, Ww0, h1);
#pragma unroll
    for (int k = 0; k < 32; ++k) { h1[k] = fmaxf(h1[k], 0.f); h2[k] = Wb1[k]; }
    dense_reg<32, 32, 32>(h1, Ww1, h2);
#pragma unroll
    for (int k = 0; k < 32; ++k) h2[k] = fmaxf(h2[k], 0.f);
#pragma unroll
    for (int k = 0; k < 44; ++k) o[k] = Wb2[k];
    dense_reg<32, 42, 44>(h2, Ww2, o);
    float* pr = prow + (size_t)j * AC_ROW;
#pragma unroll
    for (int k = 0; k < 42; ++k) pr[k] = o[k];
    pr[42] = info[2];
    if (j == il) {  // get_self_info (compiled.py:100-125): row (i,i), columns 3..10
      float* xr = X + (size_t)a * CTRM_DIM_X;
#pragma unroll
      for (int k = 0; k < 8; ++k) xr[k] = info[3 + k];
    }
  }
  __syncwarp();
  // rank neighbours by distance (argsort of column 2, format_input.py:239-243); self (distance 0) first,
  // ties broken by index
  int k_nb = N - 1;
  if (use_k_neighbor) k_nb = min(num_neighbors, k_nb);
  for (int j = lane; j < N; j += 32) {
    float dj = prow[(size_t)j * AC_ROW + 42];
    int rank = 0;
    if (j != il) {
      rank = 1;
      for (int j2 = 0; j2 < N; ++j2) {
        if (j2 == il || j2 == j) continue;
        float d2 = prow[(size_t)j2 * AC_ROW + 42];
        rank += (d2 < dj || (d2 == dj && j2 < j)) ? 1 : 0;
      }
    }
    order[rank] = j;
  }
  __syncwarp();
  // similarity to own attention vector, softmax over the k nearest (self excluded unless include_self_attention)
  const float* self_row = prow + (size_t)il * AC_ROW;
  const int first = include_self ? 0 : 1;
  float mx = -INFINITY;
  for (int r = first + lane; r <= k_nb; r += 32) {
    const float* pr = prow + (size_t)order[r] * AC_ROW;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      float d = pr[32 + c] - self_row[32 + c];
      s += d * d;
    }
    s = -s;
    pr = nullptr;
    prow[(size_t)order[r] * AC_ROW + 43] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
  float sum = 0.f;
  for (int r = first + lane; r <= k_nb; r += 32) {
    float* pr = prow + (size_t)order[r] * AC_ROW;
    float ev = expf(pr[43] - mx);
    pr[43] = ev;
    sum += ev;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
  __syncwarp();
  // comm_i[c] = sum over the k neighbours (rank order) of w_j * msg_ij[c]; lane = channel c
  float comm = 0.f;
  for (int r = 1; r <= k_nb; ++r) {
    const float* pr = prow + (size_t)order[r] * AC_ROW;
    comm += pr[lane] * (pr[43] / sum);
  }
  float* xr = X + (size_t)a * CTRM_DIM_X;
  xr[8 + lane] = __ldg(&e_self[(size_t)a * 32 + lane]);
  xr[40 + lane] = comm;
}

int launch_agent_comm(const DevModel& m, const Batch& bt, const double* d_cur, const double* d_prev, const float* d_e_self,
                      const float* d_vain_proj, float* d_X, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  const int n_max = bt.n_max;
  size_t smem = (size_t)(11 * 32 + 32 * 32 + 32 * 44 + 32 + 44) * sizeof(float) +
                (size_t)AC_WARPS * n_max * AC_ROW * sizeof(float) + (size_t)AC_WARPS * n_max * sizeof(int);
  if (smem > 200 * 1024) return ctrm_fail(-1, "instance too large for agent_comm shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    CTRM_CUDA_OK(cudaFuncSetAttribute(agent_comm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  agent_comm_kernel<<<(bt.A + AC_WARPS - 1) / AC_WARPS, AC_WARPS * 32, smem, st>>>(
      bt, d_cur, d_prev, d_e_self, d_vain_proj, m.blob, m.off, m.desc.num_neighbors, m.desc.use_k_neighbor,
      m.desc.include_self_attention, n_max, d_X);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
