// ONE THREAD-BLOCK CLUSTER PER INSTANCE: the whole roll-out loop of
// get_timed_roadmaps_multiple_paths_with_learned_indicator (roadmap_learned/learned_sampler.py:65-196) inside one kernel.
//
// The batch path (api.cu enqueue_step) runs every roll-out step as seven kernels over all instances; with a handful of
// instances each of those kernels is one 128-row tensor-core tile that is 80 % padding and costs a launch plus three to seven
// dependent MMA round trips (12-18 us per kernel, 63 us per step, 96 ms for one 28-agent instance).  Here the N_traj x max_T
// steps of an instance never leave the chip: a cluster of C = 1, 2, 4 or 8 CTAs owns the instance, every CTA keeps all network
// weights but the FOV encoders' first layer resident in shared memory (92 KB fp32), the agents are split over the CTAs, and the
// per-step phases are separated by __syncthreads / barrier.cluster instead of kernel boundaries:
//
//   raster   FOV crops as bits (cost_to_go.cpp:109-186), every CTA for all agents but only its K-slices of the 2 F^2 inputs
//   fov L0   the 2F^2 -> 64 first layers of both FOV encoders (fov_encoder.py:34-59): K is cut into 8 slices of 96 inputs, CTA r
//            accumulates the slices it owns for ALL agents (its part of W0 stays in its L1), partial sums go through L2
//   -- cluster barrier --
//   fov tail own agents: sum of the 8 partials in slice order (the result does not depend on C), 32->32->32 tails, vain_proj
//   -- cluster barrier --
//   comm     own agents on the learned branch: k nearest neighbours, pair features (compiled.py:30-97), AgentEncoder
//            43->32->32->42 on the k+1 pairs, attention softmax and message sum (format_input.py:212-277)
//   cvae     indicator + argmax, prior, K relaxed one-hot samples, decoder (cvae.py:57-79, :128-138; learned_sampler.py:99-121)
//   step     select + merge_samples, one warp per own agent: the device functions of the batch kernels (rollout.cuh), hence
//            the same bits for every fp64 operation, verdict and counter
//   -- cluster barrier --
//   advance  every CTA reads all agents' arrival flags / next positions and advances its own copy of (k_traj, t, makespan)
//   -- cluster barrier --
//
// The networks run on the CUDA cores in plain fp32 (FMA, fixed summation orders): at 28 rows per step a 128-row MMA tile with a
// TMEM round trip per layer is slower than 512 threads walking 32-wide layers out of shared memory.  They agree with the
// tensor-core kernels of the batch path (fp32-exact BF16x3 / integer digits) and with the reference to ~1e-6 relative; all
// geometry, draws and roadmap updates are the same code as the batch path.  Under teacher forcing the two paths are bit-identical.
#include <cooperative_groups.h>

#include "rollout.cuh"

namespace cg = cooperative_groups;

#define IK_THREADS 512
#define IK_WARPS 16
#define IK_NMAX 64     // agents per instance
#define IK_SLICES 8    // K-slices of the FOV first layer (96 inputs each)
#define IK_WORDS 24    // 32-bit words of one agent's crop bits
#define IK_CHUNK 8     // learned agents per pass of the comm / cvae phases
#define IK_PMAX 16     // pairs per agent: itself + at most 15 neighbours
#define IK_OBS_MAX 32  // obstacles staged per warp in shared memory (more: read through L1)
#define IK_KMAX 8      // CVAE candidates per agent

struct InstArgs {
  const float* blob;
  ctrm_weight_offsets_t off;
  int F, num_neighbors, use_k_neighbor, include_self;
  int nn;        // run the networks (0: teacher-forced run that does not record the CVAE output)
  int C;         // CTAs per instance
  int n_w;       // floats of the blob from fov_b0 on (resident in shared memory)
  float* part;   // [B][IK_SLICES][n_max][64] partial sums of the FOV first layers
};

// get_normed_vec_mag (compiled.py:11-27), as in agent_tc.cu: difference in fp64, rounded once, normalised in fp32
__device__ __forceinline__ void ik_nvm(double vx, double vy, float* o) {
  const float x = (float)vx, y = (float)vy;
  const float mag = sqrtf(fmaf(x, x, y * y));
  const float inv = __frcp_rn((mag == 0.f) ? 1.f : mag);
  o[0] = x * inv;
  o[1] = y * inv;
  o[2] = mag;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
  return v;
}

// out[lane] = bias[lane] + sum_k in_k * W[k][lane] for a 32-wide input held one value per lane (W: [32][ld] in shared memory)
__device__ __forceinline__ float warp_layer32(float in, const float* __restrict__ W, int ld, int col, float bias) {
  float acc = bias;
#pragma unroll
  for (int k = 0; k < 32; ++k) acc = fmaf(__shfl_sync(0xFFFFFFFFu, in, k), W[k * ld + col], acc);
  return acc;
}

// rows x 32 -> rows x (4 nq) layer out of shared memory: item = (row, four output columns); in: [rows][ld_in], W: [K][ld_w]
template <bool RELU>
__device__ __forceinline__ void block_layer(const float* __restrict__ in, int ld_in, int K, const float* __restrict__ W, int ld_w,
                                            const float* __restrict__ bias, int rows, int nq, float* __restrict__ out, int ld_out,
                                            int tid) {
  for (int it = tid; it < rows * nq; it += IK_THREADS) {
    const int row = it / nq, q = it - row * nq;
    const float* x = in + (size_t)row * ld_in;
    float4 acc = *reinterpret_cast<const float4*>(bias + 4 * q);
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(W + k * ld_w + 4 * q);
      const float v = x[k];
      acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
    }
    float* o = out + (size_t)row * ld_out + 4 * q;
    if (RELU) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
  }
}

template <int W>
__global__ void __launch_bounds__(IK_THREADS, 1) instance_rollout_kernel(Batch bt, InstArgs ia) {
  extern __shared__ __align__(16) uint8_t ik_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = ia.C;
  const int b = blockIdx.x / C, r = blockIdx.x - b * C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a0 = bt.agent_off[b], n = bt.agent_off[b + 1] - a0;
  const int o0 = bt.obs_off[b], nO = bt.obs_off[b + 1] - o0;
  const int M = bt.M, F = ia.F, FF = F * F, KD = 2 * FF, buf = (F - 1) / 2;
  const int q_own = (n + C - 1) / C;
  const int own_lo = min(n, r * q_own), own_hi = min(n, own_lo + q_own), n_own = own_hi - own_lo;
  const int spc = IK_SLICES / C, s_lo = r * spc, s_hi = s_lo + spc;  // K-slices of this CTA
  const ctrm_weight_offsets_t& off = ia.off;

  // ---- shared memory
  float* Wsm = reinterpret_cast<float*>(ik_smem);  // blob[fov_b0 .. total)
  uint8_t* p = ik_smem + (((size_t)ia.n_w * 4 + 15) & ~(size_t)15);
  auto carve = [&](size_t bytes) { uint8_t* q = p; p += (bytes + 15) & ~(size_t)15; return q; };
  double* s_cur = reinterpret_cast<double*>(carve(IK_NMAX * 2 * 8));
  double* s_prev = reinterpret_cast<double*>(carve(IK_NMAX * 2 * 8));
  double* s_obs = reinterpret_cast<double*>(carve(IK_WARPS * 3 * IK_OBS_MAX * 8));
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(carve(IK_WARPS * IK_NMAX * 8));
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(carve(IK_NMAX * IK_WORDS * 4));
  float* s_vp = reinterpret_cast<float*>(carve(IK_NMAX * 32 * 4));
  float* s_eself = reinterpret_cast<float*>(carve(IK_NMAX * 32 * 4));
  float* s_X = reinterpret_cast<float*>(carve(IK_CHUNK * CTRM_DIM_X * 4));
  int* s_order = reinterpret_cast<int*>(carve(IK_CHUNK * IK_PMAX * 4));
  int* s_pj = reinterpret_cast<int*>(carve(IK_CHUNK * IK_PMAX * 4));
  int* s_cx = reinterpret_cast<int*>(carve(IK_NMAX * 4));
  int* s_cy = reinterpret_cast<int*>(carve(IK_NMAX * 4));
  int* s_cc = reinterpret_cast<int*>(carve(IK_NMAX * 4));
  int* s_lrn = reinterpret_cast<int*>(carve(IK_NMAX * 4));
  int* s_ind = reinterpret_cast<int*>(carve(IK_CHUNK * 4));
  uint8_t* s_arrived = carve(IK_NMAX);
  uint8_t* s_learned = carve(IK_NMAX);
  int* s_misc = reinterpret_cast<int*>(carve(16 * 4));  // [0] n_lrn  [1] any learned
  uint8_t* U = carve(62976);
  // union scratch U
  float* u_h0 = reinterpret_cast<float*>(U);            // fov: [IK_NMAX][64]
  float* u_h1 = u_h0 + IK_NMAX * 64;                    //      [IK_NMAX][64]
  float* u_info = reinterpret_cast<float*>(U);          // comm: [128][12]
  float* u_hA = u_info + 128 * 12;                      //       [128][33]
  float* u_hB = u_hA + 128 * 33;                        //       [128][33]
  float* u_y = u_hB + 128 * 33;                         //       [128][45]
  float* u_p0 = reinterpret_cast<float*>(U);            // cvae: [IK_CHUNK][96]
  float* u_sp = u_p0 + IK_CHUNK * 96;                   //       [IK_CHUNK][64]
  float* u_d0 = u_sp + IK_CHUNK * 64;                   //       [IK_CHUNK][32]
  auto Wp = [&](int o) -> const float* { return Wsm + (o - off.fov_b0); };

  for (int i = tid; i < ia.n_w; i += IK_THREADS) Wsm[i] = __ldg(&ia.blob[off.fov_b0 + i]);
  for (int i = tid; i < n; i += IK_THREADS) {
    const double sx = bt.starts[2 * (a0 + i)], sy = bt.starts[2 * (a0 + i) + 1];
    s_cur[2 * i] = sx; s_cur[2 * i + 1] = sy;
    s_prev[2 * i] = sx; s_prev[2 * i + 1] = sy;
    s_arrived[i] = 0;
  }
  __syncthreads();

  int t = 1, kt = 0, makespan = 0, steps = 0;
  const bool all_rows = bt.tr_samples != nullptr;  // trace: the CVAE output of every agent-step is recorded
  const int k_nb = ia.use_k_neighbor ? min(ia.num_neighbors, n - 1) : n - 1;
  const float eps = 1.1920928955078125e-07f;
  float* part = ia.part + (size_t)b * IK_SLICES * bt.n_max * 64;
  const bool run = bt.N_traj > 0 && n > 0;

  while (run) {
    // ---- which agents take the learned branch in this step (learned_sampler.py:148-157): every CTA for all agents
    for (int i = tid; i < n; i += IK_THREADS)
      s_learned[i] = (all_rows || gate_is_learned(bt, b, a0 + i, i, kt, t, makespan, s_arrived[i] != 0)) ? 1 : 0;
    __syncthreads();
    if (warp == 0) {  // own learned agents in index order; whether any agent of the instance is learned
      int cnt = 0;
      bool any = false;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < n && s_learned[i];
        any = any || f;
        const bool mine = f && i >= own_lo && i < own_hi;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
        if (mine) s_lrn[cnt + __popc(m & ((1u << lane) - 1u))] = i;
        cnt += __popc(m);
      }
      any = __any_sync(0xFFFFFFFFu, any);
      if (lane == 0) { s_misc[0] = cnt; s_misc[1] = any ? 1 : 0; }
    }
    __syncthreads();
    const int n_lrn = s_misc[0];
    const bool any_learned = s_misc[1] != 0;

    if (ia.nn && any_learned) {
      // ================= raster: crop bits of every agent, this CTA's K-slices =================
      for (int i = tid; i < n; i += IK_THREADS) {
        const int cx = grid_cell(s_cur[2 * i], M), cy = grid_cell(s_cur[2 * i + 1], M);
        const bool centre_in = cx < M && cy < M;
        s_cx[i] = cx; s_cy[i] = cy;
        s_cc[i] = centre_in ? (int)__ldg(&bt.ctg[(size_t)(a0 + i) * ctg_field_elems(M, 1) + ctg_index(cx, cy, M, 1)]) : -1;
      }
      __syncthreads();
      {
        const int nw = 3 * spc, w_lo = 3 * s_lo;
        const uint8_t* occ = bt.occ + (size_t)b * M * M;
        for (int it = warp; it < n * nw; it += IK_WARPS) {
          const int i = it / nw, w = w_lo + (it - i * nw);
          const int j = 32 * w + lane;
          const int cc = s_cc[i];
          bool bit = false;
          if (j < KD && cc >= 0) {
            const int ch = j >= FF ? 1 : 0, jj = j - ch * FF;
            const int xf = jj / F, yf = jj - xf * F;
            const int x = s_cx[i] - buf + xf, y = s_cy[i] - buf + yf;
            if (x >= 0 && x < M && y >= 0 && y < M) {
              if (!ch) {
                bit = __ldg(&occ[x * M + y]) == 0;
              } else {
                const int c = (int)__ldg(&bt.ctg[(size_t)(a0 + i) * ctg_field_elems(M, 1) + ctg_index(x, y, M, 1)]);
                bit = c != CTRM_UNREACH && (cc == CTRM_UNREACH || c < cc);
              }
            }
          }
          const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
          if (lane == 0) s_bits[i * IK_WORDS + w] = word;
        }
      }
      __syncthreads();
      // ================= fov L0: partial sums of this CTA's slices, thread = (agent, four of the 64 columns) =================
      {
        const float* W0 = ia.blob + off.fov_w0;
        const int quad = tid & 15;
        for (int i = tid >> 4; i < n; i += IK_THREADS / 16) {
          for (int s = s_lo; s < s_hi; ++s) {
            float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              const uint32_t bits = s_bits[i * IK_WORDS + 3 * s + w];
              const int jb = 96 * s + 32 * w;
              const int nb = min(32, KD - jb);
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
              for (int bp = 0; bp < nb; ++bp) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(W0 + (size_t)(jb + bp) * 64 + 4 * quad));
                const float m = (bits >> bp) & 1u ? 1.f : 0.f;
                acc.x = fmaf(wv.x, m, acc.x); acc.y = fmaf(wv.y, m, acc.y); acc.z = fmaf(wv.z, m, acc.z); acc.w = fmaf(wv.w, m, acc.w);
              }
              tot.x += acc.x; tot.y += acc.y; tot.z += acc.z; tot.w += acc.w;
            }
            __stcg(reinterpret_cast<float4*>(part + ((size_t)s * bt.n_max + i) * 64 + 4 * quad), tot);
          }
        }
      }
      cluster.sync();
      // ================= fov tails of the own agents =================
      {
        const float* b0 = Wp(off.fov_b0);
        for (int it = tid; it < n_own * 64; it += IK_THREADS) {
          const int io = it >> 6, o = it & 63;
          float acc = b0[o];
#pragma unroll
          for (int s = 0; s < IK_SLICES; ++s) acc += __ldcg(part + ((size_t)s * bt.n_max + own_lo + io) * 64 + o);
          u_h0[io * 64 + o] = fmaxf(acc, 0.f);
        }
        __syncthreads();
        // item = (own agent, encoder e, four columns): hidden 1 -> hidden 2
        for (int it = tid; it < n_own * 16; it += IK_THREADS) {
          const int io = it >> 4, e = (it >> 3) & 1, q = it & 7;
          const float* Wm = Wp(e ? off.fovv_w1 : off.fovs_w1);
          const float* x = u_h0 + io * 64 + 32 * e;
          float4 acc = *reinterpret_cast<const float4*>(Wp(e ? off.fovv_b1 : off.fovs_b1) + 4 * q);
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(Wm + k * 32 + 4 * q);
            const float v = x[k];
            acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
          }
          float* o = u_h1 + io * 64 + 32 * e + 4 * q;
          o[0] = fmaxf(acc.x, 0.f); o[1] = fmaxf(acc.y, 0.f); o[2] = fmaxf(acc.z, 0.f); o[3] = fmaxf(acc.w, 0.f);
        }
        __syncthreads();
        // hidden 2 -> e_self (shared memory) | e_vain (back into u_h0)
        for (int it = tid; it < n_own * 16; it += IK_THREADS) {
          const int io = it >> 4, e = (it >> 3) & 1, q = it & 7;
          const float* Wm = Wp(e ? off.fovv_w2 : off.fovs_w2);
          const float* x = u_h1 + io * 64 + 32 * e;
          float4 acc = *reinterpret_cast<const float4*>(Wp(e ? off.fovv_b2 : off.fovs_b2) + 4 * q);
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(Wm + k * 32 + 4 * q);
            const float v = x[k];
            acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
          }
          float* o = e ? (u_h0 + io * 64 + 4 * q) : (s_eself + io * 32 + 4 * q);
          o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
        }
        __syncthreads();
        // vain_proj = b0 + W0[11:43]^T e_vain (agent_encoder.py:57-60), for every CTA of the cluster through L2
        for (int it = tid; it < n_own * 8; it += IK_THREADS) {
          const int io = it >> 3, q = it & 7;
          const float* Wm = Wp(off.ag_w0v);
          const float* x = u_h0 + io * 64;
          float4 acc = *reinterpret_cast<const float4*>(Wp(off.ag_b0) + 4 * q);
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(Wm + k * 32 + 4 * q);
            const float v = x[k];
            acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
          }
          __stcg(reinterpret_cast<float4*>(bt.vain_proj + (size_t)(a0 + own_lo + io) * 32 + 4 * q), acc);
        }
      }
      cluster.sync();
      // ================= comm + cvae for the own learned agents, IK_CHUNK at a time =================
      if (n_lrn > 0) {
        for (int i = tid; i < n * 8; i += IK_THREADS)
          reinterpret_cast<float4*>(s_vp)[i] = __ldcg(reinterpret_cast<const float4*>(bt.vain_proj + (size_t)a0 * 32) + i);
      }
      for (int c0 = 0; c0 < n_lrn; c0 += IK_CHUNK) {
        const int nc = min(IK_CHUNK, n_lrn - c0);
        __syncthreads();  // s_vp loaded / the previous chunk has left the scratch
        // ---- k nearest neighbours (format_input.py:239-243): float32(sqrt_f64(dx^2 + dy^2)) keys, ties by index, self first
        if (warp < nc) {
          const int il = s_lrn[c0 + warp];
          const double cix = s_cur[2 * il], ciy = s_cur[2 * il + 1];
          unsigned long long* key = s_keys + warp * IK_NMAX;
          for (int j = lane; j < n; j += 32) {
            const double dx = f64sub(s_cur[2 * j], cix), dy = f64sub(s_cur[2 * j + 1], ciy);
            const float d = __double2float_rn(__dsqrt_rn(f64add(f64mul(dx, dx), f64mul(dy, dy))));
            key[j] = (j == il) ? 0ull : (((unsigned long long)(__float_as_uint(d) + 1u) << 32) | (unsigned)j);
          }
          __syncwarp();
          for (int j = lane; j < n; j += 32) {
            const unsigned long long kj = key[j];
            int rank = 0;
            for (int j2 = 0; j2 < n; ++j2) rank += key[j2] < kj ? 1 : 0;
            if (rank <= k_nb) s_order[warp * IK_PMAX + rank] = j;
          }
        }
        __syncthreads();
        // ---- pair features (compiled.py:30-97) of (agent, slot): slot 0 = the agent itself, slot s = its s-th nearest
        if (tid < nc * IK_PMAX) {
          const int c = tid / IK_PMAX, s = tid - c * IK_PMAX;
          if (s <= k_nb) {
            const int il = s_lrn[c0 + c], j = s_order[c * IK_PMAX + s];
            const double cix = s_cur[2 * il], ciy = s_cur[2 * il + 1];
            float* info = u_info + tid * 12;
            ik_nvm(f64sub(s_cur[2 * j], cix), f64sub(s_cur[2 * j + 1], ciy), info + 0);
            ik_nvm(f64sub(__ldg(&bt.goals[2 * (a0 + j)]), cix), f64sub(__ldg(&bt.goals[2 * (a0 + j) + 1]), ciy), info + 3);
            ik_nvm(f64sub(s_prev[2 * j], cix), f64sub(s_prev[2 * j + 1], ciy), info + 6);
            info[9] = (float)__ldg(&bt.rads[a0 + j]);
            info[10] = (float)__ldg(&bt.speeds[a0 + j]);
            s_pj[tid] = j;
            if (s == 0) {  // get_self_info (compiled.py:100-125): row (i, i), columns 3..10
              float* xr = s_X + c * CTRM_DIM_X;
#pragma unroll
              for (int k = 0; k < 8; ++k) xr[k] = info[3 + k];
            }
          } else {
            s_pj[tid] = -1;
          }
        }
        __syncthreads();
        // ---- AgentEncoder on the pairs: item = (pair, four columns)
        for (int it = tid; it < nc * IK_PMAX * 8; it += IK_THREADS) {
          const int pr = it >> 3, q = it & 7;
          const int j = s_pj[pr];
          if (j < 0) continue;
          const float* Wm = Wp(off.ag_w0i);
          const float* x = u_info + pr * 12;
          float4 acc = *reinterpret_cast<const float4*>(s_vp + j * 32 + 4 * q);
#pragma unroll
          for (int k = 0; k < CTRM_DIM_INFO; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(Wm + k * 32 + 4 * q);
            const float v = x[k];
            acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
          }
          float* o = u_hA + pr * 33 + 4 * q;
          o[0] = fmaxf(acc.x, 0.f); o[1] = fmaxf(acc.y, 0.f); o[2] = fmaxf(acc.z, 0.f); o[3] = fmaxf(acc.w, 0.f);
        }
        __syncthreads();
        block_layer<true>(u_hA, 33, 32, Wp(off.ag_w1), 32, Wp(off.ag_b1), nc * IK_PMAX, 8, u_hB, 33, tid);
        __syncthreads();
        block_layer<false>(u_hB, 33, 32, Wp(off.ag_w2), 44, Wp(off.ag_b2), nc * IK_PMAX, 11, u_y, 45, tid);
        __syncthreads();
        // ---- attention softmax over the neighbours and message sum (format_input.py:250-275); X = [self_info | e_self | comm]
        if (warp < nc) {
          const int il = s_lrn[c0 + warp];
          const float* y0 = u_y + (warp * IK_PMAX) * 45;
          const int s = lane;
          const bool has_pair = s <= k_nb && s < IK_PMAX;
          float sim = 0.f;
          if (has_pair) {
            const float* ys = y0 + s * 45;
#pragma unroll
            for (int c = 0; c < 10; ++c) {
              const float d = ys[32 + c] - y0[32 + c];
              sim += d * d;
            }
          }
          sim = -sim;
          const int first = ia.include_self ? 0 : 1;
          const bool in_softmax = has_pair && s >= first;
          const float mx = warp_max(in_softmax ? sim : -INFINITY);
          const float ev = in_softmax ? expf(sim - mx) : 0.f;
          const float sum = warp_sum(ev);
          const float wgt = in_softmax ? ev / sum : 0.f;
          float comm = 0.f;
          for (int rr = 1; rr <= k_nb; ++rr) comm += y0[rr * 45 + lane] * __shfl_sync(0xFFFFFFFFu, wgt, rr);
          float* xr = s_X + warp * CTRM_DIM_X;
          xr[40 + lane] = comm;
          xr[8 + lane] = s_eself[(il - own_lo) * 32 + lane];
        }
        __syncthreads();
        // ---- first layers of indicator | encoder_input | decoder (x rows) side by side: item = (agent, net, four columns)
        for (int it = tid; it < nc * 24; it += IK_THREADS) {
          const int c = it / 24, rem = it - c * 24, net = rem >> 3, q = rem & 7;
          const float* Wm = net == 0 ? Wp(off.ind_w0) : (net == 1 ? Wp(off.enc_w0) : Wp(off.dec_w0) + 64 * 32);
          const float* x = s_X + c * CTRM_DIM_X;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
          for (int k = 0; k < CTRM_DIM_X; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(Wm + k * 32 + 4 * q);
            const float v = x[k];
            acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
          }
          float* o = u_p0 + c * 96 + 32 * net + 4 * q;
          o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
        }
        __syncthreads();
        // ---- indicator -> argmax (learned_sampler.py:99-104); prior -> sqrt of the clamped probabilities (cvae.py:57-68,
        //      :132-136; cvae_tc.cu); decoder first layer without its z rows.  One warp per agent, a hidden vector = a register
        if (warp < nc) {
          const float* p0 = u_p0 + warp * 96;
          float h = fmaxf(p0[lane] + Wp(off.ind_b0)[lane], 0.f);
          h = fmaxf(warp_layer32(h, Wp(off.ind_w1), 32, lane, Wp(off.ind_b1)[lane]), 0.f);
          const float lg = warp_layer32(h, Wp(off.ind_w2), 4, lane & 3, Wp(off.ind_b2)[lane & 3]);
          const float l0 = __shfl_sync(0xFFFFFFFFu, lg, 0), l1 = __shfl_sync(0xFFFFFFFFu, lg, 1), l2 = __shfl_sync(0xFFFFFFFFu, lg, 2);
          int ind = 0;
          float best = l0;
          if (l1 > best) { best = l1; ind = 1; }
          if (l2 > best) { best = l2; ind = 2; }
          if (lane == 0) {
            s_ind[warp] = ind;
            bt.ind[a0 + s_lrn[c0 + warp]] = ind;
          }
          u_d0[warp * 32 + lane] = (p0[64 + lane] + Wp(off.dec_b0)[lane]) + Wp(off.dec_w0)[(64 + CTRM_DIM_X + ind) * 32 + lane];
          float he = fmaxf((p0[32 + lane] + Wp(off.enc_b0)[lane]) + Wp(off.enc_w0)[(CTRM_DIM_X + ind) * 32 + lane], 0.f);
          he = fmaxf(warp_layer32(he, Wp(off.enc_w1), 32, lane, Wp(off.enc_b1)[lane]), 0.f);
          const float g0 = warp_layer32(he, Wp(off.enc_w2), 64, lane, Wp(off.enc_b2)[lane]);
          const float g1 = warp_layer32(he, Wp(off.enc_w2), 64, 32 + lane, Wp(off.enc_b2)[32 + lane]);
          const float mx = warp_max(fmaxf(g0, g1));
          const float lse = logf(warp_sum(expf(g0 - mx) + expf(g1 - mx)));
          const float q0 = expf((g0 - mx) - lse), q1 = expf((g1 - mx) - lse);  // probs = exp(log_softmax)
          const float ps = warp_sum(q0 + q1);                                  // Categorical normalises probs / sum
          u_sp[warp * 64 + lane] = sqrtf(fminf(fmaxf(q0 / ps, eps), 1.f - eps));
          u_sp[warp * 64 + 32 + lane] = sqrtf(fminf(fmaxf(q1 / ps, eps), 1.f - eps));
        }
        __syncthreads();
        // ---- K relaxed one-hot samples and the decoder, one warp per (agent, copy): z_i = sqrt(p_i / E_i) / sum (temperature 2,
        //      cvae_tc.cu), lane holds latents 2 lane and 2 lane + 1
        for (int task = warp; task < nc * bt.K; task += IK_WARPS) {
          const int c = task / bt.K, k = task - c * bt.K;
          const int il = s_lrn[c0 + c];
          const uint4 w4 = philox4x32(bt.inst_base + (uint32_t)b, rng_ctr1(kt, t), (uint32_t)(k * n + il),
                                      rng_ctr3(STREAM_GUMBEL, (uint32_t)(lane >> 1)), bt.seed_lo, bt.seed_hi);
          const float ua = u32_to_f32((lane & 1) ? w4.z : w4.x), ub = u32_to_f32((lane & 1) ? w4.w : w4.y);
          const float Ea = -logf(fminf(fmaxf(ua, eps), 1.f - eps)), Eb = -logf(fminf(fmaxf(ub, eps), 1.f - eps));
          const float ra = u_sp[c * 64 + 2 * lane] * rsqrtf(Ea), rb = u_sp[c * 64 + 2 * lane + 1] * rsqrtf(Eb);
          const float inv = 1.f / warp_sum(ra + rb);
          const float za = fminf(ra * inv, 1.f), zb = fminf(rb * inv, 1.f);
          const float* Wz = Wp(off.dec_w0);
          float h = u_d0[c * 32 + lane];
#pragma unroll
          for (int m = 0; m < 32; ++m) {
            h = fmaf(__shfl_sync(0xFFFFFFFFu, za, m), Wz[(2 * m) * 32 + lane], h);
            h = fmaf(__shfl_sync(0xFFFFFFFFu, zb, m), Wz[(2 * m + 1) * 32 + lane], h);
          }
          h = fmaxf(h, 0.f);
          h = fmaxf(warp_layer32(h, Wp(off.dec_w1), 32, lane, Wp(off.dec_b1)[lane]), 0.f);
          const float o = warp_layer32(h, Wp(off.dec_w2), 4, lane & 3, Wp(off.dec_b2)[lane & 3]);
          if (lane < 3) bt.samples[((size_t)(a0 + il) * bt.K + k) * 3 + lane] = o;
        }
      }
      __syncthreads();  // the samples are in place
    }

    // ================= step: select + merge_samples, one warp per own agent (rollout.cuh) =================
    for (int i = own_lo + warp; i < own_hi; i += IK_WARPS) {
      const int a = a0 + i;
      StepCtx sc;
      sc.b = b; sc.t = t; sc.kt = kt; sc.makespan = makespan;
      sc.a0 = a0; sc.N = n; sc.o0 = o0; sc.nO = nO;
      sc.inv_nO = (nO >= 1 && nO <= 128) ? c_pair_inv.v[nO] : 0u;
      const double2 q0 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a);
      const double2 q1 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 2);
      const double2 q2 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 4);
      sc.gx = q0.x; sc.gy = q0.y; sc.speed = q1.x; sc.speed2 = q1.y; sc.merge2 = q2.x; sc.rad = q2.y;
      sc.sobs = nullptr;
      if (nO <= IK_OBS_MAX) {
        double* so = s_obs + warp * 3 * IK_OBS_MAX;
        __syncwarp();
        for (int o = lane; o < nO; o += 32) {
          const double* g = bt.obs_xyr + 3 * (size_t)(o0 + o);
          so[3 * o] = __ldg(g);
          so[3 * o + 1] = __ldg(g + 1);
          so[3 * o + 2] = f64add(sc.rad, __ldg(g + 2));  // rad1 + rad2, rounded once (collision_check.cpp:56)
        }
        __syncwarp();
        sc.sobs = so;
      }
      const double cx = s_cur[2 * i], cy = s_cur[2 * i + 1];
      MergePrefetch<W> pf;
      merge_prefetch<W>(bt, a, t, lane, pf);
      double lx, ly;
      unsigned int count = select_agent(bt, sc, a, lane, cx, cy, s_arrived[i] != 0, &lx, &ly);
      count += merge_agent<W>(bt, sc, a, lane, lx, ly, pf);
      if (lane == 0 && count) atomicAdd(&bt.cnt_collide[b], (unsigned long long)count);
    }
    cluster.sync();
    // ================= advance (learned_sampler.py:186-196, :65-73): every CTA keeps its own copy of the instance state ====
    bool arr_ok = true;
    double2 nx = make_double2(0.0, 0.0);
    if (tid < n) {
      arr_ok = __ldcg(&bt.arrived[a0 + tid]) != 0;
      nx = __ldcg(reinterpret_cast<const double2*>(bt.nxt + 2 * (size_t)(a0 + tid)));
      s_arrived[tid] = arr_ok ? 1 : 0;
    }
    const bool all = __syncthreads_and(arr_ok) != 0;
    cluster.sync();  // every CTA has read this step's flags and positions before the next step overwrites them
    ++steps;
    bool end_roll = false;
    if (all) {
      makespan = makespan ? max(makespan, t) : t;
      end_roll = true;
    } else if (t >= bt.max_T) {
      end_roll = true;
    }
    if (!end_roll) {
      if (tid < n) {
        s_prev[2 * tid] = s_cur[2 * tid]; s_prev[2 * tid + 1] = s_cur[2 * tid + 1];
        s_cur[2 * tid] = nx.x; s_cur[2 * tid + 1] = nx.y;
      }
      ++t;
    } else {
      ++kt;
      if (kt >= bt.N_traj) break;
      if (tid < n) {  // next roll-out restarts from the starts (learned_sampler.py:68-73)
        const double sx = bt.starts[2 * (a0 + tid)], sy = bt.starts[2 * (a0 + tid) + 1];
        s_cur[2 * tid] = sx; s_cur[2 * tid + 1] = sy;
        s_prev[2 * tid] = sx; s_prev[2 * tid + 1] = sy;
        s_arrived[tid] = 0;
        if (tid >= own_lo && tid < own_hi) bt.arrived[a0 + tid] = 0;
      }
      t = 1;
    }
    __syncthreads();
  }
  if (r == 0 && tid == 0) {
    bt.k_traj[b] = kt; bt.t[b] = t; bt.makespan[b] = makespan; bt.done[b] = 1;
    bt.steps_done[b] = steps;
    int* ip = bt.inst_pack + 8 * b;
    ip[0] = 1; ip[1] = t; ip[2] = kt; ip[3] = makespan;
    atomicAdd(bt.n_done, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t instance_smem_bytes(int n_w) {
  auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t s = al((size_t)n_w * 4);
  s += 2 * al(IK_NMAX * 2 * 8) + al((size_t)IK_WARPS * 3 * IK_OBS_MAX * 8) + al((size_t)IK_WARPS * IK_NMAX * 8);
  s += al(IK_NMAX * IK_WORDS * 4) + 2 * al(IK_NMAX * 32 * 4) + al(IK_CHUNK * CTRM_DIM_X * 4) + 2 * al(IK_CHUNK * IK_PMAX * 4);
  s += 4 * al(IK_NMAX * 4) + al(IK_CHUNK * 4) + 2 * al(IK_NMAX) + al(16 * 4) + al(62976);
  return s;
}

// can the per-instance kernel run this batch?  (everything else takes the batch path)
int instance_kernel_supports(const DevModel& m, const Batch& bt) {
  const ctrm_model_desc& d = m.desc;
  if (bt.n_max > IK_NMAX || bt.K > IK_KMAX || bt.randomize_indicator) return 0;
  if (d.temp != 2.0f) return 0;
  if (2 * d.fov_size * d.fov_size > 32 * IK_WORDS) return 0;
  const int k_nb = d.use_k_neighbor ? (d.num_neighbors < bt.n_max - 1 ? d.num_neighbors : bt.n_max - 1) : bt.n_max - 1;
  if (k_nb + 1 > IK_PMAX) return 0;
  if (bt.tab_uniforms != nullptr) return 0;
  return 1;
}

size_t instance_part_floats(const Batch& bt) { return (size_t)bt.B * IK_SLICES * (bt.n_max > 0 ? bt.n_max : 1) * 64; }

int launch_instance_rollout(const DevModel& m, const Batch& bt, int nn, int C, float* d_part, cudaStream_t st) {
  if (bt.B <= 0) return 0;
  if (C != 1 && C != 2 && C != 4 && C != 8) return ctrm_fail(-1, "instance kernel: cluster size must be 1, 2, 4 or 8");
  InstArgs ia;
  ia.blob = m.blob;
  ia.off = m.off;
  ia.F = m.desc.fov_size;
  ia.num_neighbors = m.desc.num_neighbors;
  ia.use_k_neighbor = m.desc.use_k_neighbor;
  ia.include_self = m.desc.include_self_attention;
  ia.nn = nn;
  ia.C = C;
  ia.n_w = m.off.total - m.off.fov_b0;
  ia.part = d_part;
  const size_t smem = instance_smem_bytes(ia.n_w);
  if (smem > 227 * 1024) return ctrm_fail(-1, "instance kernel: weights do not fit in shared memory");
  static SmemAttrCache attr[4];
  switch (bt.W) {
    case 1: CTRM_CUDA_OK(attr[0].ensure(instance_rollout_kernel<1>, smem)); break;
    case 2: CTRM_CUDA_OK(attr[1].ensure(instance_rollout_kernel<2>, smem)); break;
    case 3: CTRM_CUDA_OK(attr[2].ensure(instance_rollout_kernel<3>, smem)); break;
    case 4: CTRM_CUDA_OK(attr[3].ensure(instance_rollout_kernel<4>, smem)); break;
    default: return ctrm_fail(-1, "unsupported adjacency mask width");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(bt.B * C), 1, 1);
  cfg.blockDim = dim3(IK_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  switch (bt.W) {
    case 1: CTRM_CUDA_OK(cudaLaunchKernelEx(&cfg, instance_rollout_kernel<1>, bt, ia)); break;
    case 2: CTRM_CUDA_OK(cudaLaunchKernelEx(&cfg, instance_rollout_kernel<2>, bt, ia)); break;
    case 3: CTRM_CUDA_OK(cudaLaunchKernelEx(&cfg, instance_rollout_kernel<3>, bt, ia)); break;
    default: CTRM_CUDA_OK(cudaLaunchKernelEx(&cfg, instance_rollout_kernel<4>, bt, ia)); break;
  }
  return 0;
}
