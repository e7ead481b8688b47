// Feature encoders of Format2D_CTRM_Input.get_feature_one_step (learning/formats/format_input.py:279-357):
//   (the FOV encoders themselves run on the tensor cores: fov_tc.cu)
//   agent_comm : numba get_arr_others_info (learning/compiled.py:30-97), AgentEncoder.forward
//                (learning/model/agent_encoder.py:34-65; format_input.py:328-334), get_comm
//                (format_input.py:212-277), get_self_info (compiled.py:100-125) and the final concat
//                (format_input.py:354-355), one warp per agent i, never materialising the N^2 tensors.
// fp32 with FMA; the summation order differs from ATen's blocked sgemm, parity is 1e-5 relative (BASELINE.json).
#include <cstring>

#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// pairwise features + agent encoder + neighbour attention, one warp per agent i
// ------------------------------------------------------------------------------------------------
#define AC_WARPS 4

// get_normed_vec_mag (compiled.py:11-27).  The reference works in float64 and casts to float32 (format_input.py:327);
// here the difference vector is formed in float64 (no cancellation error), rounded once to float32 and normalised in
// float32 — within 2e-7 relative of the reference's values, far inside the 1e-5 bar.  Three fp64 sqrt + six fp64
// divisions per pair were ~20% of this kernel's issue slots.
__device__ __forceinline__ void nvm(double vx, double vy, float* o) {
  const float x = (float)vx, y = (float)vy;
  const float mag = sqrtf(fmaf(x, x, y * y));
  const float d = (mag == 0.f) ? 1.f : mag;
  o[0] = x / d;
  o[1] = y / d;
  o[2] = mag;
}

// AgentEncoder weights in constant memory: with fully unrolled loops every weight is an immediate constant-bank operand of
// its FFMA (c[bank][offset]), so the MLP issues no load instructions at all — the shared-memory version spent its time
// waiting on one broadcast LDS.128 per four FMAs (short-scoreboard stalls, profiles/r1_full_step_kernels_b1024_tc.md).
// layout: w0i [11][32] | w1 [32][32] | w2 [32][44] | b1 [32] | b2 [44]
#define AG_W0 0
#define AG_W1 (11 * 32)
#define AG_W2 (AG_W1 + 32 * 32)
#define AG_B1 (AG_W2 + 32 * 44)
#define AG_B2 (AG_B1 + 32)
#define AG_TOTAL (AG_B2 + 44)
__constant__ float c_ag[AG_TOTAL];

template <int IN, int OUT, int LDW, int BASE>
__device__ __forceinline__ void dense_const(const float* x, float* acc) {
  // acc[OUT] += x[IN] . c_ag[BASE + k * LDW + j]
#pragma unroll
  for (int k = 0; k < IN; ++k) {
    const float xv = x[k];
#pragma unroll
    for (int j = 0; j < OUT; ++j) acc[j] = fmaf(xv, c_ag[BASE + k * LDW + j], acc[j]);
  }
}

int upload_agent_constants(const float* h_blob, const ctrm_weight_offsets_t& off) {
  float h[AG_TOTAL];
  memcpy(h + AG_W0, h_blob + off.ag_w0i, 11 * 32 * sizeof(float));
  memcpy(h + AG_W1, h_blob + off.ag_w1, 32 * 32 * sizeof(float));
  memcpy(h + AG_W2, h_blob + off.ag_w2, 32 * 44 * sizeof(float));
  memcpy(h + AG_B1, h_blob + off.ag_b1, 32 * sizeof(float));
  memcpy(h + AG_B2, h_blob + off.ag_b2, 44 * sizeof(float));
  CTRM_CUDA_OK(cudaMemcpyToSymbol(c_ag, h, sizeof(h)));
  return 0;
}

// Only the k nearest neighbours of an agent ever reach its communication vector (format_input.py:250-275), so the
// 43->32->32->42 MLP runs for P = k + 1 pairs per agent (self + k nearest) instead of all N.  A warp serves G = 32 / P
// agents (two for the trained k = 15): lane = (agent g, slot s), slot 0 = the agent itself, slot s = its s-th nearest.
__global__ void __launch_bounds__(AC_WARPS * 32, 6) agent_comm_kernel(Batch bt, const double* __restrict__ cur,
                                                                   const double* __restrict__ prev,
                                                                   const float* __restrict__ e_self,
                                                                   const float* __restrict__ vain_proj,
                                                                   const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                                   int num_neighbors, int use_k_neighbor, int include_self,
                                                                   int n_max, int P, int G, float* __restrict__ X) {
  extern __shared__ __align__(16) float smem[];
  float* dist_base = smem;                                          // [AC_WARPS][G][n_max] fp32 distances
  int* order_base = reinterpret_cast<int*>(dist_base + (size_t)AC_WARPS * G * n_max);  // [AC_WARPS][G][P]
  float* msg_base = reinterpret_cast<float*>(order_base + (size_t)AC_WARPS * G * P);   // [AC_WARPS][G][P][33]: w*msg | w
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / P, s = lane - g * P;
  const int a = (blockIdx.x * AC_WARPS + warp) * G + g;
  bool active = g < G && a < bt.A;
  int b = 0;
  if (active) {
    b = bt.agent_inst[a];
    if (bt.done != nullptr && bt.done[b]) active = false;
  }
  if (!__any_sync(0xFFFFFFFFu, active)) return;
  const int a0 = active ? bt.agent_off[b] : 0, N = active ? bt.agent_off[b + 1] - a0 : 0, il = a - a0;
  float* dist = dist_base + (size_t)(warp * G + (g < G ? g : 0)) * n_max;
  int* order = order_base + (size_t)(warp * G + (g < G ? g : 0)) * P;
  float* mrow = msg_base + (size_t)(warp * G + (g < G ? g : 0)) * P * 33;
  int k_nb = N - 1;
  if (use_k_neighbor) k_nb = min(num_neighbors, k_nb);
  const double cix = active ? cur[2 * a] : 0.0, ciy = active ? cur[2 * a + 1] : 0.0;
  // ---- distances to every agent of the instance (column 2 of others_info, the fp32 value the reference argsorts)
  if (active)
    for (int j = s; j < N; j += P) {
      const float dx = (float)f64sub(cur[2 * (a0 + j)], cix), dy = (float)f64sub(cur[2 * (a0 + j) + 1], ciy);
      dist[j] = sqrtf(fmaf(dx, dx, dy * dy));
    }
  __syncwarp();
  // ---- rank (argsort, format_input.py:239-243): self first, ties broken by index; keep ranks 0..k
  if (active)
    for (int j = s; j < N; j += P) {
      int rank = 0;
      if (j != il) {
        const float dj = dist[j];
        rank = 1;
        for (int j2 = 0; j2 < N; ++j2) {
          if (j2 == il || j2 == j) continue;
          const float d2 = dist[j2];
          rank += (d2 < dj || (d2 == dj && j2 < j)) ? 1 : 0;
        }
      }
      if (rank <= k_nb) order[rank] = j;
    }
  __syncwarp();
  // ---- pair MLP for slot s (pair i -> order[s])
  const bool has_pair = active && s <= k_nb;
  float o[44];
  if (has_pair) {
    const int aj = a0 + order[s];
    float info[CTRM_DIM_INFO];
    nvm(f64sub(cur[2 * aj], cix), f64sub(cur[2 * aj + 1], ciy), info + 0);
    nvm(f64sub(bt.goals[2 * aj], cix), f64sub(bt.goals[2 * aj + 1], ciy), info + 3);
    nvm(f64sub(prev[2 * aj], cix), f64sub(prev[2 * aj + 1], ciy), info + 6);
    info[9] = (float)bt.rads[aj];
    info[10] = (float)bt.speeds[aj];
    float h1[32], h2[32];
    {
      const float4* vp = reinterpret_cast<const float4*>(vain_proj + (size_t)aj * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = __ldg(&vp[q]);
        h1[q * 4] = v.x; h1[q * 4 + 1] = v.y; h1[q * 4 + 2] = v.z; h1[q * 4 + 3] = v.w;
      }
    }
    dense_const<CTRM_DIM_INFO, 32, 32, AG_W0>(info, h1);
#pragma unroll
    for (int k = 0; k < 32; ++k) { h1[k] = fmaxf(h1[k], 0.f); h2[k] = c_ag[AG_B1 + k]; }
    dense_const<32, 32, 32, AG_W1>(h1, h2);
#pragma unroll
    for (int k = 0; k < 32; ++k) h2[k] = fmaxf(h2[k], 0.f);
#pragma unroll
    for (int k = 0; k < 42; ++k) o[k] = c_ag[AG_B2 + k];
    o[42] = 0.f; o[43] = 0.f;
    dense_const<32, 42, 44, AG_W2>(h2, o);
    if (s == 0) {  // get_self_info (compiled.py:100-125): row (i,i), columns 3..10
      float* xr = X + (size_t)a * CTRM_DIM_X;
#pragma unroll
      for (int k = 0; k < 8; ++k) xr[k] = info[3 + k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 44; ++k) o[k] = 0.f;
  }
  // ---- similarity to the agent's own attention vector (slot 0 of the group), softmax over the neighbours
  float sim = 0.f;
  {
    const int src = (g < G ? g : 0) * P;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      const float self_att = __shfl_sync(0xFFFFFFFFu, o[32 + c], src);
      const float d = o[32 + c] - self_att;
      sim += d * d;
    }
    sim = -sim;
  }
  const int first = include_self ? 0 : 1;
  const bool in_softmax = has_pair && s >= first;
  // softmax over the group's slots first..k_nb: max and sum by register shuffles (one expf per lane)
  const int gbase = (g < G ? g : 0) * P;
  float mx = -INFINITY;
  for (int r = 0; r < P; ++r) {
    const float v = __shfl_sync(0xFFFFFFFFu, sim, gbase + r);
    if (r >= first && r <= k_nb) mx = fmaxf(mx, v);
  }
  const float ev = in_softmax ? expf(sim - mx) : 0.f;
  float sum = 0.f;
  for (int r = 0; r < P; ++r) {
    const float v = __shfl_sync(0xFFFFFFFFu, ev, gbase + r);
    if (r >= first && r <= k_nb) sum += v;
  }
  const float w = in_softmax ? ev / sum : 0.f;
  __syncwarp();
  if (has_pair) {
#pragma unroll
    for (int c = 0; c < 32; ++c) mrow[s * 33 + c] = o[c] * w;
  }
  __syncwarp();
  // ---- comm_i[c] = sum over the neighbours in rank order of w_j * msg_ij[c]; X row = [self_info | e_self | comm]
  if (active) {
    float* xr = X + (size_t)a * CTRM_DIM_X;
    for (int c = s; c < 32; c += P) {
      float comm = 0.f;
      for (int r = 1; r <= k_nb; ++r) comm += mrow[r * 33 + c];
      xr[40 + c] = comm;
      xr[8 + c] = __ldg(&e_self[(size_t)a * 32 + c]);
    }
  }
}

int launch_agent_comm(const DevModel& m, const Batch& bt, const double* d_cur, const double* d_prev, const float* d_e_self,
                      const float* d_vain_proj, float* d_X, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  const int n_max = bt.n_max;
  int P = (m.desc.use_k_neighbor ? (m.desc.num_neighbors < n_max - 1 ? m.desc.num_neighbors : n_max - 1) : n_max - 1) + 1;
  if (P < 1) P = 1;
  if (P > 32) return ctrm_fail(-1, "agent_comm: more than 31 neighbours per agent are not supported");
  const int G = 32 / P;
  size_t smem = (size_t)AC_WARPS * G * ((size_t)n_max * sizeof(float) + (size_t)P * sizeof(int) + (size_t)P * 33 * sizeof(float));
  if (smem > 200 * 1024) return ctrm_fail(-1, "instance too large for agent_comm shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    CTRM_CUDA_OK(cudaFuncSetAttribute(agent_comm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int agents_per_block = AC_WARPS * G;
  agent_comm_kernel<<<(bt.A + agents_per_block - 1) / agents_per_block, AC_WARPS * 32, smem, st>>>(
      bt, d_cur, d_prev, d_e_self, d_vain_proj, m.blob, m.off, m.desc.num_neighbors, m.desc.use_k_neighbor,
      m.desc.include_self_attention, n_max, P, G, d_X);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
