// Feature encoders of Format2D_CTRM_Input.get_feature_one_step (learning/formats/format_input.py:279-357):
//   (the FOV encoders themselves run on the tensor cores: fov_tc.cu)
//   agent_comm : numba get_arr_others_info (learning/compiled.py:30-97), AgentEncoder.forward
//                (learning/model/agent_encoder.py:34-65; format_input.py:328-334), get_comm
//                (format_input.py:212-277), get_self_info (compiled.py:100-125) and the final concat
//                (format_input.py:354-355), one warp per agent i, never materialising the N^2 tensors.
// fp32 with FMA; the summation order differs from ATen's blocked sgemm, parity is 1e-5 relative (BASELINE.json).
#include "common.cuh"
#include "kernels.h"

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
