// Agent communication of Format2D_CTRM_Input.get_feature_one_step (learning/formats/format_input.py:279-357) on the
// tensor cores:
//   numba get_arr_others_info (learning/compiled.py:30-97), AgentEncoder.forward (learning/model/agent_encoder.py:34-65;
//   format_input.py:328-334), get_comm (format_input.py:212-277), get_self_info (compiled.py:100-125) and the final
//   concat (format_input.py:354-355), never materialising the N^2 tensors.
//
// Only the k nearest neighbours of an agent ever reach its communication vector (format_input.py:250-275), so the
// 43->32->32->42 MLP runs for P = k + 1 pairs per agent (self + k nearest) instead of all N.  A tile is 128 pair rows:
// warp w serves G = 32 / P agents (two for the trained k = 15), lane = (agent g, slot s), slot 0 = the agent itself, slot
// s = its s-th nearest; row = thread = TMEM lane.  The three layers are three rounds of fp32-exact BF16x3 MMAs
// (tc.cuh, two accumulator classes); layer 0's FOV part (32 of its 43 inputs) depends only on agent j and arrives
// pre-multiplied from fov_encode (vain_proj), so its MMA is K = 16 over the 11 pair features.  Each epilogue adds the
// bias, applies ReLU, splits the row into three bf16 parts and writes it back as the next A operand.
#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

#define AT_THREADS 256
#define AT_TMEM_COLS 128
#define AT_A_PART tc::tile_bytes(128, 32)
#define AT_A0_PART tc::tile_bytes(128, 16)
#define AT_W0_PART tc::tile_bytes(32, 16)
#define AT_W1_PART tc::tile_bytes(32, 32)
#define AT_W2_PART tc::tile_bytes(48, 32)
#define AT_FIXED_SMEM (3 * AT_A_PART + 3 * AT_W0_PART + 3 * AT_W1_PART + 3 * AT_W2_PART + (32 + 48) * 4 + 16)

// get_normed_vec_mag (compiled.py:11-27).  The reference works in float64 and casts to float32 (format_input.py:327);
// here the difference vector is formed in float64 (no cancellation error), rounded once to float32 and normalised in
// float32 — within 3e-7 relative of the reference's values, far inside the 1e-5 bar.
__device__ __forceinline__ void nvm(double vx, double vy, float* o) {
  const float x = (float)vx, y = (float)vy;
  const float mag = sqrtf(fmaf(x, x, y * y));
  const float inv = __frcp_rn((mag == 0.f) ? 1.f : mag);  // one correctly rounded reciprocal, two multiplies: <= 1.5 ulp
  o[0] = x * inv;
  o[1] = y * inv;
  o[2] = mag;
}

#define AT_W_BYTES (3 * AT_W0_PART + 3 * AT_W1_PART + 3 * AT_W2_PART + (32 + 48) * 4)
// weight operands of agent_comm exactly as they sit in its shared memory from W0t on
__global__ void __launch_bounds__(AT_THREADS) agent_comm_image_kernel(const float* __restrict__ blob, ctrm_weight_offsets_t off,
                                                                    uint8_t* __restrict__ image) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* W0t = smem;
  uint8_t* W1t = W0t + 3 * AT_W0_PART;
  uint8_t* W2t = W1t + 3 * AT_W1_PART;
  float* B1 = reinterpret_cast<float*>(W2t + 3 * AT_W2_PART);
  float* B2 = B1 + 32;
  const int tid = threadIdx.x;
  {
    // layer-0 input order: the half-0 thread of a row produces k 0..7 = [cur(3) goal(3) rad speed], the half-1 thread
    // k 8..10 = prev(3); others_info columns (compiled.py:69-97) are cur 0-2, goal 3-5, prev 6-8, rad 9, speed 10
    const int kperm[16] = {0, 1, 2, 3, 4, 5, 9, 10, 6, 7, 8, -1, -1, -1, -1, -1};
    const float* w0 = blob + off.ag_w0i;
    for (int i = tid; i < 32 * 2; i += AT_THREADS) {
      const int k0 = (i >> 5) * 8, n = i & 31;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = kperm[k0 + j] >= 0 ? __ldg(w0 + kperm[k0 + j] * 32 + n) : 0.f;
      tc::store_split8(W0t, W0t + AT_W0_PART, W0t + 2 * AT_W0_PART, n, k0, 16, v);
    }
  }
  tc::stage_weight_split(blob + off.ag_w1, 32, 0, 32, 32, 32, 32, W1t, tid, AT_THREADS);
  tc::stage_weight_split(blob + off.ag_w2, 44, 0, 32, 42, 48, 32, W2t, tid, AT_THREADS);
  if (tid < 32) B1[tid] = __ldg(&blob[off.ag_b1 + tid]);
  if (tid < 48) B2[tid] = tid < 42 ? __ldg(&blob[off.ag_b2 + tid]) : 0.f;
  __syncthreads();
  tc::dump_image(smem, image, AT_W_BYTES, tid, AT_THREADS);
}

// warps w and w + 4 share the rows (TMEM lanes) 32 (w % 4) .. +31; they meet on their own named barrier
__device__ __forceinline__ void pair_sync(int wq) { asm volatile("bar.sync %0, 64;" ::"r"(wq + 1) : "memory"); }

// Two threads per pair row: thread (row, half) ranks every second candidate, computes half of the pair features and owns
// hidden columns [16 half, 16 half + 16) in the epilogues.  The kernel is bound by the latency of its dependent loads
// (positions -> ranks -> neighbour ids -> neighbour state) and of the three MMA round trips, so warps per SM matter most.
__global__ void __launch_bounds__(AT_THREADS, 4) agent_comm_tc_kernel(Batch bt, const double* __restrict__ cur,
                                                                     const double* __restrict__ prev,
                                                                     const float* __restrict__ e_self,
                                                                     const float* __restrict__ vain_proj,
                                                                     const uint8_t* __restrict__ w_image, int num_neighbors, int use_k_neighbor, int include_self,
                                                                     int n_max, int P, int G, float* __restrict__ X, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* At = smem;                  // A operand, 3 parts: [128 x 16] (layer 0) / [128 x 32] (layers 1, 2); later w * msg rows
  uint8_t* W0t = At + 3 * AT_A_PART;   // W0[0:11] parts 1..3, [32 x 16], input order permuted (see kperm)
  uint8_t* W1t = W0t + 3 * AT_W0_PART; // W1, [32 x 32]
  uint8_t* W2t = W1t + 3 * AT_W1_PART; // W2, [48 x 32] (42 outputs + zero padding)
  float* B1 = reinterpret_cast<float*>(W2t + 3 * AT_W2_PART);
  float* B2 = B1 + 32;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(B2 + 48);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  unsigned long long* key_base = reinterpret_cast<unsigned long long*>(smem + AT_FIXED_SMEM);  // [4][G][n_max] sort keys
  int* order_base = reinterpret_cast<int*>(key_base + (size_t)4 * G * n_max);                  // [4][G][P]
  float* msg_base = reinterpret_cast<float*>(At);                                // [4][G][P][33]: w * msg (A is dead by then)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, half = tid >> 7, wq = warp & 3;
  const int g = lane / P, s = lane - g * P, gg = g < G ? g : 0;

  if (tid == 0) {  // the weight operands arrive as one prebuilt image (agent_comm_image_kernel), through the TMA engine
    tc::mbar_init(tc::smem_u32(mbar), 1);
    tc::tma_load_image(W0t, w_image, AT_W_BYTES, tc::smem_u32(mbar));
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, AT_TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t mb = tc::smem_u32(mbar);
  tc::mbar_wait(mb, 0);  // weights have landed; the MMA rounds continue on the same barrier with phase 1
  const uint32_t sA = tc::smem_u32(At), sW0 = tc::smem_u32(W0t), sW1 = tc::smem_u32(W1t), sW2 = tc::smem_u32(W2t);
  uint32_t phase = 1;
  unsigned long long* key = key_base + (size_t)(wq * G + gg) * n_max;
  int* order = order_base + (size_t)(wq * G + gg) * P;
  float* mrow = msg_base + (size_t)(wq * G + gg) * P * 33;
  const bool p_pow2 = (P & (P - 1)) == 0;

  const int n_rows = bt.live_agent != nullptr ? *bt.n_live : bt.A;  // agents of the unfinished instances, or all
  n_tiles = min(n_tiles, (n_rows + 4 * G - 1) / (4 * G));
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int arow = (tile * 4 + wq) * G + g;
    bool active = g < G && arow < n_rows;
    const int a = active ? (bt.live_agent != nullptr ? bt.live_agent[arow] : arow) : 0;
    int b = 0;
    if (active) {
      b = bt.live_agent != nullptr ? bt.live_inst[arow] : bt.agent_inst[a];
      if (bt.done != nullptr && bt.done[b]) active = false;
    }
    // also: every shared-memory / TMEM read of the previous tile is done before this one overwrites them
    tc::fence_before_sync();
    if (!__syncthreads_or(active)) continue;
    const int a0 = active ? bt.agent_off[b] : 0, N = active ? bt.agent_off[b + 1] - a0 : 0, il = a - a0;
    int k_nb = N - 1;
    if (use_k_neighbor) k_nb = min(num_neighbors, k_nb);
    const double cix = active ? cur[2 * a] : 0.0, ciy = active ? cur[2 * a + 1] : 0.0;
    // ---- distances to every agent of the instance (column 2 of others_info, the fp32 value the reference argsorts) as
    //      sort keys (distance bits + 1, index); the agent itself gets key 0: self first, ties broken by index
    if (active)
      for (int j = s + P * half; j < N; j += 2 * P) {
        // exactly the value the reference argsorts: float32(sqrt_f64(dx^2 + dy^2)) (compiled.py:24, format_input.py:327) —
        // an fp32 evaluation can swap two neighbours whose distances differ by an fp32 ulp on the k-th boundary
        const double dx = f64sub(cur[2 * (a0 + j)], cix), dy = f64sub(cur[2 * (a0 + j) + 1], ciy);
        const float d = __double2float_rn(__dsqrt_rn(f64add(f64mul(dx, dx), f64mul(dy, dy))));
        key[j] = (j == il) ? 0ull : (((unsigned long long)(__float_as_uint(d) + 1u) << 32) | (unsigned)j);
      }
    pair_sync(wq);
    // ---- rank (argsort, format_input.py:239-243); keep ranks 0..k
    if (active)
      for (int j = s + P * half; j < N; j += 2 * P) {
        const unsigned long long kj = key[j];
        int rank = 0;
        for (int j2 = 0; j2 < N; ++j2) rank += key[j2] < kj ? 1 : 0;
        if (rank <= k_nb) order[rank] = j;
      }
    pair_sync(wq);
    // ---- pair features of slot s (pair i -> order[s]) become row `row` of the layer-0 A operand
    const bool has_pair = active && s <= k_nb;
    int aj = 0;
    {
      float info[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) info[k] = 0.f;
      if (has_pair) {
        aj = a0 + order[s];
        float* xr = X + (size_t)a * CTRM_DIM_X;  // get_self_info (compiled.py:100-125): row (i,i), columns 3..10
        if (half == 0) {
          nvm(f64sub(cur[2 * aj], cix), f64sub(cur[2 * aj + 1], ciy), info + 0);
          nvm(f64sub(bt.goals[2 * aj], cix), f64sub(bt.goals[2 * aj + 1], ciy), info + 3);
          info[6] = (float)bt.rads[aj];
          info[7] = (float)bt.speeds[aj];
          if (s == 0) {
            xr[0] = info[3]; xr[1] = info[4]; xr[2] = info[5];
            xr[6] = info[6]; xr[7] = info[7];
          }
        } else {
          nvm(f64sub(prev[2 * aj], cix), f64sub(prev[2 * aj + 1], ciy), info + 0);
          if (s == 0) {
            xr[3] = info[0]; xr[4] = info[1]; xr[5] = info[2];
          }
        }
      }
      tc::store_act8(At, At + AT_A0_PART, At + 2 * AT_A0_PART, row, 8 * half, 16, info);
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::mma_bf16x3_c2(tbase, 32, sA, AT_A0_PART, sW0, AT_W0_PART, 16, 32);
      tc::commit(mb);
    }
    float h[16], acc[16];
    {  // FOV part of layer 0 (+ bias), gathered while the MMA runs
      const float4* vp = reinterpret_cast<const float4*>(vain_proj + (size_t)aj * 32 + 16 * half);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = has_pair ? __ldg(&vp[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
        h[q * 4] = v.x; h[q * 4 + 1] = v.y; h[q * 4 + 2] = v.z; h[q * 4 + 3] = v.w;
      }
    }
    tc::mbar_wait(mb, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    tc::tmem_ld16_sum2(tbase, warp, 16 * half, 32, acc);
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = fmaxf(h[k] + acc[k], 0.f);
#pragma unroll
    for (int q = 0; q < 2; ++q) tc::store_act8(At, At + AT_A_PART, At + 2 * AT_A_PART, row, 16 * half + q * 8, 32, h + q * 8);
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::mma_bf16x3_c2(tbase, 32, sA, AT_A_PART, sW1, AT_W1_PART, 32, 32);
      tc::commit(mb);
    }
    tc::mbar_wait(mb, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    tc::tmem_ld16_sum2(tbase, warp, 16 * half, 32, acc);
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = fmaxf(acc[k] + B1[16 * half + k], 0.f);
#pragma unroll
    for (int q = 0; q < 2; ++q) tc::store_act8(At, At + AT_A_PART, At + 2 * AT_A_PART, row, 16 * half + q * 8, 32, h + q * 8);
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::mma_bf16x3_c2(tbase, 48, sA, AT_A_PART, sW2, AT_W2_PART, 32, 48);
      tc::commit(mb);
    }
    tc::mbar_wait(mb, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    // this thread's 16 message columns in acc, the attention head (10 of 16 columns; read by both halves) in h
    tc::tmem_ld16_sum2(tbase, warp, 16 * half, 48, acc);
    tc::tmem_ld16_sum2(tbase, warp, 32, 48, h);
    // ---- similarity to the agent's own attention vector (slot 0 of the group), softmax over the neighbours
    float sim = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      const float mine = has_pair ? h[c] + B2[32 + c] : 0.f;
      const float self_att = __shfl_sync(0xFFFFFFFFu, mine, gg * P);
      const float d = mine - self_att;
      sim += d * d;
    }
    sim = -sim;
    const int first = include_self ? 0 : 1;
    const bool in_softmax = has_pair && s >= first;
    float mx, sum, ev;
    if (p_pow2) {  // aligned power-of-two groups: butterflies
      mx = in_softmax ? sim : -INFINITY;
      for (int d = 1; d < P; d <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
      ev = in_softmax ? expf(sim - mx) : 0.f;
      sum = ev;
      for (int d = 1; d < P; d <<= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    } else {
      const int gbase = gg * P;
      mx = -INFINITY;
      for (int r = 0; r < P; ++r) {
        const float v = __shfl_sync(0xFFFFFFFFu, sim, gbase + r);
        if (r >= first && r <= k_nb) mx = fmaxf(mx, v);
      }
      ev = in_softmax ? expf(sim - mx) : 0.f;
      sum = 0.f;
      for (int r = 0; r < P; ++r) {
        const float v = __shfl_sync(0xFFFFFFFFu, ev, gbase + r);
        if (r >= first && r <= k_nb) sum += v;
      }
    }
    const float w = in_softmax ? ev / sum : 0.f;
    // the A operand is dead (layer 2 has completed for the whole tile): its memory now carries w * msg, pair-local rows
    if (has_pair) {
#pragma unroll
      for (int c = 0; c < 16; ++c) mrow[s * 33 + 16 * half + c] = (acc[c] + B2[16 * half + c]) * w;
    }
    pair_sync(wq);
    // ---- comm_i[c] = sum over the neighbours in rank order of w_j * msg_ij[c]; X row = [self_info | e_self | comm]
    if (active) {
      float* xr = X + (size_t)a * CTRM_DIM_X;
      for (int c = s + P * half; c < 32; c += 2 * P) {
        float comm = 0.f;
        for (int r = 1; r <= k_nb; ++r) comm += mrow[r * 33 + c];
        xr[40 + c] = comm;
        xr[8 + c] = __ldg(&e_self[(size_t)a * 32 + c]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, AT_TMEM_COLS);
}

size_t agent_comm_image_bytes() { return AT_W_BYTES; }
int launch_agent_comm_image(const float* d_blob, const ctrm_weight_offsets_t& off, uint8_t* d_image, cudaStream_t st) {
  agent_comm_image_kernel<<<1, AT_THREADS, AT_W_BYTES, st>>>(d_blob, off, d_image);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_agent_comm(const DevModel& m, const Batch& bt, const double* d_cur, const double* d_prev, const float* d_e_self,
                      const float* d_vain_proj, float* d_X, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  const int n_max = bt.n_max;
  int P = (m.desc.use_k_neighbor ? (m.desc.num_neighbors < n_max - 1 ? m.desc.num_neighbors : n_max - 1) : n_max - 1) + 1;
  if (P < 1) P = 1;
  if (P > 32) return ctrm_fail(-1, "agent_comm: more than 31 neighbours per agent are not supported");
  const int G = 32 / P;
  const size_t smem = (size_t)AT_FIXED_SMEM + (size_t)4 * G * ((size_t)n_max * sizeof(unsigned long long) + (size_t)P * sizeof(int));
  if (smem > 200 * 1024) return ctrm_fail(-1, "instance too large for agent_comm shared memory");
  static SmemAttrCache attr;
  CTRM_CUDA_OK(attr.ensure(agent_comm_tc_kernel, smem));
  const int agents_per_tile = 4 * G;
  const int n_tiles = (bt.A + agents_per_tile - 1) / agents_per_tile;
  const int grid = n_tiles < 4 * 148 ? n_tiles : 4 * 148;  // persistent: four CTAs per SM (128 TMEM columns each)
  agent_comm_tc_kernel<<<grid, AT_THREADS, smem, st>>>(bt, d_cur, d_prev, d_e_self, d_vain_proj, m.agent_img,
                                                       m.desc.num_neighbors, m.desc.use_k_neighbor,
                                                       m.desc.include_self_attention, n_max, P, G, d_X, n_tiles);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
