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
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "rollout.cuh"

namespace cg = cooperative_groups;

#define CHECK_RC(x)     \
  do {                  \
    int _rc = (x);      \
    if (_rc) return _rc; \
  } while (0)

#define IK_THREADS 512
#define IK_WARPS 16
#define IK_NMAX 64     // agents per instance
#define IK_SLICES 8    // K-slices of the FOV first layer (96 inputs each)
#define IK_WORDS 24    // 32-bit words of one agent's crop bits
#define IK_CHUNK 8     // learned agents per pass of the comm / cvae phases
#define IK_PMAX 16     // pairs per agent: itself + at most 15 neighbours
#define IK_OBS_MAX 32  // obstacles staged per warp in shared memory (more: read through L1)
#define IK_KMAX 8      // CVAE candidates per agent
#define IK_NPROF 16

struct InstArgs {
  const float* blob;
  ctrm_weight_offsets_t off;
  int F, num_neighbors, use_k_neighbor, include_self;
  int nn;        // run the networks (0: teacher-forced run that does not record the CVAE output)
  int C;         // CTAs per instance
  int n_w;       // floats of the blob from fov_b0 on (resident in shared memory)
  int w0_frag;   // the CTA's K-slice of the FOV first layer is resident as bf16 x 3 mma.sync fragments (C = 8: one slice per CTA)
  int chunk;     // learned agents per pass of the comm / cvae phases (4 or 8)
  int u_bytes;   // size of the phase scratch
  int vc_smem;   // the own agents' layer counts live in shared memory during the roll-outs
  int own_max;   // ceil(n_max / C)
  float* part;   // [B][IK_SLICES][n_max][64] partial sums of the FOV first layers
  long long* prof;  // optional [B * C][IK_NPROF] clock cycles per phase (CTRM_INSTANCE_PROF=1)
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

// four consecutive rows p0 .. p0 + 3 of a [rows x K] . [K x 32 (+ NC2)] layer out of shared memory, one warp: lane = output column
// (and column 32 + lane for lane < NC2).  Per four k: four conflict-free weight loads and four broadcast 16-byte activation
// loads feed sixteen FMAs per lane — item-per-thread tilings of these layers were bound by shared-memory wavefronts.
// brow[r]: the bias row of row r (nullptr: zero); in rows are 16-byte aligned (ldin a multiple of 4), K a multiple of 4.
template <int K, int NC2, bool RELU>
__device__ __forceinline__ void warp_rows4(const float* __restrict__ in, int ldin, const float* __restrict__ Wt, int ldw,
                                           const float* const* brow, float* __restrict__ out, int ldout, int p0, int lane) {
  float acc[4], acc2[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    acc[r] = brow[r] != nullptr ? brow[r][lane] : 0.f;
    acc2[r] = (NC2 > 0 && lane < NC2 && brow[r] != nullptr) ? brow[r][32 + lane] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    float w[4], w2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[i] = Wt[(k + i) * ldw + lane];
      w2[i] = (NC2 > 0 && lane < NC2) ? Wt[(k + i) * ldw + 32 + lane] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(in + (size_t)(p0 + r) * ldin + k);
      acc[r] = fmaf(x.x, w[0], acc[r]); acc[r] = fmaf(x.y, w[1], acc[r]); acc[r] = fmaf(x.z, w[2], acc[r]); acc[r] = fmaf(x.w, w[3], acc[r]);
      if (NC2 > 0) {
        acc2[r] = fmaf(x.x, w2[0], acc2[r]); acc2[r] = fmaf(x.y, w2[1], acc2[r]); acc2[r] = fmaf(x.z, w2[2], acc2[r]); acc2[r] = fmaf(x.w, w2[3], acc2[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    out[(size_t)(p0 + r) * ldout + lane] = RELU ? fmaxf(acc[r], 0.f) : acc[r];
    if (NC2 > 0 && lane < NC2) out[(size_t)(p0 + r) * ldout + 32 + lane] = RELU ? fmaxf(acc2[r], 0.f) : acc2[r];
  }
}

// four output columns 4q .. 4q + 3 of a 32-input layer row; the inputs are cut into four strands (kq = lane & 3, adjacent lanes)
// that meet by shuffle: chains of 8 instead of 32 — these small layers run with a handful of warps and are bound by the
// length of one thread's dependent chain.  Whole warps must call this together.
__device__ __forceinline__ float4 splitk_32(const float* __restrict__ x, const float* __restrict__ Wm, int q, int kq) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int k = 8 * kq + kk;
    const float4 w = *reinterpret_cast<const float4*>(Wm + k * 32 + 4 * q);
    const float v = x[k];
    acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
  }
#pragma unroll
  for (int d = 1; d <= 2; d <<= 1) {
    acc.x += __shfl_xor_sync(0xFFFFFFFFu, acc.x, d); acc.y += __shfl_xor_sync(0xFFFFFFFFu, acc.y, d);
    acc.z += __shfl_xor_sync(0xFFFFFFFFu, acc.z, d); acc.w += __shfl_xor_sync(0xFFFFFFFFu, acc.w, d);
  }
  return acc;
}

// D[16 x 8] += A[16 x 16] . B[16 x 8], bf16 operands, fp32 accumulators (warp-level mma.sync)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two crop bits -> two bf16 values (0.0 / 1.0) in one register
__device__ __forceinline__ uint32_t bits2_bf16(uint32_t t) { return (t & 1u) * 0x3F80u + ((t >> 1) & 1u) * 0x3F800000u; }
// part `sp` (0 hi, 1 mid, 2 lo) of the three-way bf16 split of an fp32 weight: hi + mid + lo == w to 24 bits
__device__ __forceinline__ uint32_t bf16_part(float w, int sp) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const float r1 = w - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  return (uint32_t)__bfloat16_as_ushort(sp == 0 ? hi : (sp == 1 ? mid : lo));
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
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(carve(IK_CHUNK * IK_NMAX * 8));
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
  uint8_t* U = carve((size_t)ia.u_bytes);
  double* s_goal = reinterpret_cast<double*>(carve(IK_NMAX * 2 * 8));
  float* s_radf = reinterpret_cast<float*>(carve(IK_NMAX * 4));
  float* s_speedf = reinterpret_cast<float*>(carve(IK_NMAX * 4));
  uint2* s_w0f = ia.w0_frag ? reinterpret_cast<uint2*>(carve(8 * 6 * 3 * 32 * 8)) : nullptr;  // [n-tile][k-step][part][lane]
  const int CH = ia.chunk;
  double* s_pack = reinterpret_cast<double*>(carve((size_t)ia.own_max * 6 * 8));         // own agents' constants (Batch::agent_pack)
  float* s_samples = reinterpret_cast<float*>(carve((size_t)ia.own_max * bt.K * 3 * 4));  // own agents' CVAE samples of the step
  int32_t* s_nl = reinterpret_cast<int32_t*>(carve((size_t)ia.own_max * 4));
  int32_t* s_vcnt = ia.vc_smem ? reinterpret_cast<int32_t*>(carve((size_t)ia.own_max * bt.L * 4)) : nullptr;
  // union scratch U
  float* u_h0 = reinterpret_cast<float*>(U);            // fov: [own agents][64]
  float* u_h1 = u_h0 + q_own * 64;                      //      [own agents][64]
  float* u_info = reinterpret_cast<float*>(U);          // comm: [CH * 16][12]
  float* u_hA = u_info + CH * IK_PMAX * 12;             //       [CH * 16][36]
  float* u_hB = u_hA + CH * IK_PMAX * 36;               //       [CH * 16][36]
  float* u_y = u_hB + CH * IK_PMAX * 36;                //       [CH * 16][48]
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
    s_goal[2 * i] = bt.goals[2 * (a0 + i)]; s_goal[2 * i + 1] = bt.goals[2 * (a0 + i) + 1];
    s_radf[i] = (float)bt.rads[a0 + i];
    s_speedf[i] = (float)bt.speeds[a0 + i];
  }
  for (int i = tid; i < n_own * 6; i += IK_THREADS) s_pack[i] = bt.agent_pack[6 * (size_t)(a0 + own_lo) + i];
  // the own agents' layer counts and sizes in shared memory (TimedRoadmap(start): one layer with the start vertex), written
  // back at the end: every step reads them first, and through global memory that is a dependent L2 round trip
  if (s_vcnt != nullptr) {
    for (int i = tid; i < n_own * bt.L; i += IK_THREADS) s_vcnt[i] = (i % bt.L) == 0 ? 1 : 0;
    for (int i = tid; i < n_own; i += IK_THREADS) s_nl[i] = 1;
  }
  if (s_w0f != nullptr) {  // this CTA's slice of W0 as mma.sync B fragments: [n-tile of 8 columns][k-step of 16][bf16 part][lane]
    const float* W0 = ia.blob + off.fov_w0;
    for (int idx = tid; idx < 8 * 6 * 3 * 32; idx += IK_THREADS) {
      const int ln = idx & 31, sp = (idx >> 5) % 3, ks = (idx / 96) % 6, nt = idx / 576;
      const int col = 8 * nt + (ln >> 2), k0 = 96 * s_lo + 16 * ks + 2 * (ln & 3);
      auto wv = [&](int j) { return j < KD ? __ldg(&W0[(size_t)j * 64 + col]) : 0.f; };
      s_w0f[idx] = make_uint2(bf16_part(wv(k0), sp) | (bf16_part(wv(k0 + 1), sp) << 16),
                              bf16_part(wv(k0 + 8), sp) | (bf16_part(wv(k0 + 9), sp) << 16));
    }
  }
  __syncthreads();

  int t = 1, kt = 0, makespan = 0, steps = 0;
  const bool all_rows = bt.tr_samples != nullptr;  // trace: the CVAE output of every agent-step is recorded
  const int k_nb = ia.use_k_neighbor ? min(ia.num_neighbors, n - 1) : n - 1;
  const float eps = 1.1920928955078125e-07f;
  float* part = ia.part + (size_t)b * IK_SLICES * bt.n_max * 64;
  const bool run = bt.N_traj > 0 && n > 0;
  long long pacc[IK_NPROF];
  long long pt0 = 0;
#pragma unroll
  for (int i = 0; i < IK_NPROF; ++i) pacc[i] = 0;
  const bool prof = ia.prof != nullptr && tid == 0;
  if (prof) pt0 = clock64();
#define IK_MARK(k)                      \
  if (prof) {                           \
    const long long now = clock64();    \
    pacc[k] += now - pt0;               \
    pt0 = now;                          \
  }

  // the batch kernels' device functions index per-agent state by the global agent number: hand them a view of the batch whose
  // hot per-agent arrays are the shared-memory copies of the own agents
  Batch bv = bt;
  bv.samples = s_samples - (size_t)(a0 + own_lo) * bt.K * 3;
  if (s_vcnt != nullptr) {
    bv.vcnt = s_vcnt - (size_t)(a0 + own_lo) * bt.L;
    bv.nlayers = s_nl - (size_t)(a0 + own_lo);
  }
  const bool obs_static = n_own <= IK_WARPS && nO <= IK_OBS_MAX;  // warp w serves own agent w in every step: stage its obstacles once
  if (obs_static && warp < n_own) {
    double* so = s_obs + warp * 3 * IK_OBS_MAX;
    const double rad = s_pack[6 * warp + 5];
    for (int o = lane; o < nO; o += 32) {
      const double* g = bt.obs_xyr + 3 * (size_t)(o0 + o);
      so[3 * o] = __ldg(g);
      so[3 * o + 1] = __ldg(g + 1);
      so[3 * o + 2] = f64add(rad, __ldg(g + 2));  // rad1 + rad2, rounded once (collision_check.cpp:56)
    }
  }
  cluster.sync();  // every CTA of the cluster is running before anyone stores into a neighbour's shared memory

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
    IK_MARK(0)

    if (ia.nn && any_learned) {
      // ================= raster: crop bits of the OWN agents (their cost fields and crops stay in this SM's L1 from step to
      // step), every word stored straight into the shared memory of the CTA that owns its K-slice (distributed shared memory)
      for (int io = tid; io < n_own; io += IK_THREADS) {
        const int i = own_lo + io;
        const int cx = grid_cell(s_cur[2 * i], M), cy = grid_cell(s_cur[2 * i + 1], M);
        const bool centre_in = cx < M && cy < M;
        s_cx[io] = cx; s_cy[io] = cy;
        s_cc[io] = centre_in ? (int)__ldg(&bt.ctg[(size_t)(a0 + i) * ctg_field_elems(M, 1) + ctg_index(cx, cy, M, 1)]) : -1;
      }
      __syncthreads();
      {
        // item = (own agent, word, quarter of the word): eight independent loads per thread, all in flight at once; the four
        // quarters meet by shuffle
        const uint8_t* occ = bt.occ + (size_t)b * M * M;
        const unsigned invF = 65536u / (unsigned)F + 1u;  // jj / F == (jj * invF) >> 16 for jj < F * F, F <= 41
        const int n_items = n_own * IK_WORDS * 4;
        for (int it0 = 0; it0 < n_items; it0 += IK_THREADS) {
          const int it = it0 + tid;
          const bool act = it < n_items;
          const int qd = it & 3, iw = it >> 2;
          const int io = act ? iw / IK_WORDS : 0, w = iw - io * IK_WORDS;
          const int cc = act ? s_cc[io] : -1;
          const int cx = s_cx[io] - buf, cy = s_cy[io] - buf;
          const uint16_t* cf = bt.ctg + (size_t)(a0 + own_lo + io) * ctg_field_elems(M, 1);
          // loads first, branch-free (a cell outside the crop / the map reads as "unreachable" / "occupied": bit 0), tests after
          int vc[8], vo[8];
          bool isc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = 32 * w + 8 * qd + e;
            const int ch = j >= FF ? 1 : 0, jj = j - ch * FF;
            const int xf = (int)(((unsigned)jj * invF) >> 16), yf = jj - xf * F;
            const int x = cx + xf, y = cy + yf;
            const bool ok = j < KD && cc >= 0 && x >= 0 && x < M && y >= 0 && y < M;
            isc[e] = ch != 0;
            vc[e] = (ok && ch) ? (int)__ldg(&cf[ctg_index(x, y, M, 1)]) : CTRM_UNREACH;
            vo[e] = (ok && !ch) ? (int)__ldg(&occ[x * M + y]) : 1;
          }
          uint32_t byte = 0u;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const bool bit = isc[e] ? (vc[e] != CTRM_UNREACH && (cc == CTRM_UNREACH || vc[e] < cc)) : (vo[e] == 0);
            byte |= (bit ? 1u : 0u) << e;
          }
          uint32_t word = byte << (8 * qd);
          word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
          word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
          if (act && qd == 0) {
            uint32_t* dst = cluster.map_shared_rank(s_bits, (unsigned)((w / 3) / spc));
            dst[(own_lo + io) * IK_WORDS + w] = word;
          }
        }
      }
      cluster.sync();  // every agent's bits of this CTA's slices have arrived
      IK_MARK(1)
      // ================= fov L0: partial sums of this CTA's slices =================
      if (s_w0f != nullptr) {
        // one slice per CTA, on the tensor cores: A = crop bits as bf16 0/1 (exact), B = the resident three-way bf16 split of W0
        // (exact to 24 bits), fp32 accumulators, the large (hi) and the small (mid, lo) products in separate accumulators.
        // warp = (n-tile of 8 columns, m-tile of 16 agents)
        const int nt = warp & 7, gid = lane >> 2, tig = lane & 3;
        for (int mt = warp >> 3; 16 * mt < n; mt += 2) {
          float accH[4] = {0.f, 0.f, 0.f, 0.f}, accL[4] = {0.f, 0.f, 0.f, 0.f};
          const int r0 = 16 * mt + gid, r1 = r0 + 8;
#pragma unroll
          for (int ks = 0; ks < 6; ++ks) {
            const int widx = 3 * s_lo + (ks >> 1), sh = 16 * (ks & 1) + 2 * tig;
            const uint32_t x0 = s_bits[r0 * IK_WORDS + widx] >> sh, x1 = s_bits[r1 * IK_WORDS + widx] >> sh;
            const uint32_t a[4] = {bits2_bf16(x0), bits2_bf16(x1), bits2_bf16(x0 >> 8), bits2_bf16(x1 >> 8)};
            const uint2* f = s_w0f + ((nt * 6 + ks) * 3) * 32 + lane;
            const uint2 fh = f[0], fm = f[32], fl = f[64];
            mma_bf16_16816(accH, a, fh.x, fh.y);
            mma_bf16_16816(accL, a, fm.x, fm.y);
            mma_bf16_16816(accL, a, fl.x, fl.y);
          }
          const int col = 8 * nt + 2 * tig;
          if (r0 < n) __stcg(reinterpret_cast<float2*>(part + ((size_t)s_lo * bt.n_max + r0) * 64 + col),
                             make_float2(accH[0] + accL[0], accH[1] + accL[1]));
          if (r1 < n) __stcg(reinterpret_cast<float2*>(part + ((size_t)s_lo * bt.n_max + r1) * 64 + col),
                             make_float2(accH[2] + accL[2], accH[3] + accL[3]));
        }
      } else {
        // CUDA cores, W0 streamed through L1 / L2: thread = (agent, four of the 64 columns)
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
      IK_MARK(2)
      cluster.sync();
      IK_MARK(3)
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
        // item = (own agent, encoder e, four columns, k strand): hidden 1 -> hidden 2
        for (int it = tid; it < n_own * 64; it += IK_THREADS) {
          const int kq = it & 3, q = (it >> 2) & 7, e = (it >> 5) & 1, io = it >> 6;
          float4 acc = splitk_32(u_h0 + io * 64 + 32 * e, Wp(e ? off.fovv_w1 : off.fovs_w1), q, kq);
          if (kq == 0) {
            const float4 bb = *reinterpret_cast<const float4*>(Wp(e ? off.fovv_b1 : off.fovs_b1) + 4 * q);
            float* o = u_h1 + io * 64 + 32 * e + 4 * q;
            o[0] = fmaxf(acc.x + bb.x, 0.f); o[1] = fmaxf(acc.y + bb.y, 0.f); o[2] = fmaxf(acc.z + bb.z, 0.f); o[3] = fmaxf(acc.w + bb.w, 0.f);
          }
        }
        __syncthreads();
        // hidden 2 -> e_self (shared memory) | e_vain (back into u_h0)
        for (int it = tid; it < n_own * 64; it += IK_THREADS) {
          const int kq = it & 3, q = (it >> 2) & 7, e = (it >> 5) & 1, io = it >> 6;
          float4 acc = splitk_32(u_h1 + io * 64 + 32 * e, Wp(e ? off.fovv_w2 : off.fovs_w2), q, kq);
          if (kq == 0) {
            const float4 bb = *reinterpret_cast<const float4*>(Wp(e ? off.fovv_b2 : off.fovs_b2) + 4 * q);
            float* o = e ? (u_h0 + io * 64 + 4 * q) : (s_eself + io * 32 + 4 * q);
            o[0] = acc.x + bb.x; o[1] = acc.y + bb.y; o[2] = acc.z + bb.z; o[3] = acc.w + bb.w;
          }
        }
        __syncthreads();
        // vain_proj = b0 + W0[11:43]^T e_vain (agent_encoder.py:57-60), for every CTA of the cluster through L2
        for (int it = tid; it < n_own * 32; it += IK_THREADS) {
          const int kq = it & 3, q = (it >> 2) & 7, io = it >> 5;
          float4 acc = splitk_32(u_h0 + io * 64, Wp(off.ag_w0v), q, kq);
          if (kq == 0) {
            const float4 bb = *reinterpret_cast<const float4*>(Wp(off.ag_b0) + 4 * q);
            __stcg(reinterpret_cast<float4*>(bt.vain_proj + (size_t)(a0 + own_lo + io) * 32 + 4 * q),
                   make_float4(acc.x + bb.x, acc.y + bb.y, acc.z + bb.z, acc.w + bb.w));
          }
        }
      }
      IK_MARK(4)
      cluster.sync();
      IK_MARK(5)
      // ================= comm + cvae for the own learned agents, IK_CHUNK at a time =================
      if (n_lrn > 0) {
        for (int i = tid; i < n * 8; i += IK_THREADS)
          reinterpret_cast<float4*>(s_vp)[i] = __ldcg(reinterpret_cast<const float4*>(bt.vain_proj + (size_t)a0 * 32) + i);
      }
      for (int c0 = 0; c0 < n_lrn; c0 += CH) {
        const int nc = min(CH, n_lrn - c0);
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
        IK_MARK(6)
        // ---- pair features (compiled.py:30-97) of (agent, slot): slot 0 = the agent itself, slot s = its s-th nearest
        if (tid < nc * IK_PMAX) {
          const int c = tid / IK_PMAX, s = tid - c * IK_PMAX;
          if (s <= k_nb) {
            const int il = s_lrn[c0 + c], j = s_order[c * IK_PMAX + s];
            const double cix = s_cur[2 * il], ciy = s_cur[2 * il + 1];
            float* info = u_info + tid * 12;
            ik_nvm(f64sub(s_cur[2 * j], cix), f64sub(s_cur[2 * j + 1], ciy), info + 0);
            ik_nvm(f64sub(s_goal[2 * j], cix), f64sub(s_goal[2 * j + 1], ciy), info + 3);
            ik_nvm(f64sub(s_prev[2 * j], cix), f64sub(s_prev[2 * j + 1], ciy), info + 6);
            info[9] = s_radf[j];
            info[10] = s_speedf[j];
            info[11] = 0.f;  // K padded to 12 (the weight row it meets is finite)
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
        IK_MARK(7)
        // ---- AgentEncoder on the pairs (agent_encoder.py:34-65): a warp takes four pair rows, lane = output column
        for (int p0 = 4 * warp; p0 < nc * IK_PMAX; p0 += 4 * IK_WARPS) {
          const float* brow[4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) brow[rr] = s_pj[p0 + rr] >= 0 ? s_vp + s_pj[p0 + rr] * 32 : nullptr;
          warp_rows4<12, 0, true>(u_info, 12, Wp(off.ag_w0i), 32, brow, u_hA, 36, p0, lane);
        }
        __syncthreads();
        for (int p0 = 4 * warp; p0 < nc * IK_PMAX; p0 += 4 * IK_WARPS) {
          const float* b1 = Wp(off.ag_b1);
          const float* brow[4] = {b1, b1, b1, b1};
          warp_rows4<32, 0, true>(u_hA, 36, Wp(off.ag_w1), 32, brow, u_hB, 36, p0, lane);
        }
        __syncthreads();
        for (int p0 = 4 * warp; p0 < nc * IK_PMAX; p0 += 4 * IK_WARPS) {
          const float* b2 = Wp(off.ag_b2);
          const float* brow[4] = {b2, b2, b2, b2};
          warp_rows4<32, 12, false>(u_hB, 36, Wp(off.ag_w2), 44, brow, u_y, 48, p0, lane);
        }
        __syncthreads();
        IK_MARK(8)
        // ---- attention softmax over the neighbours and message sum (format_input.py:250-275); X = [self_info | e_self | comm]
        if (warp < nc) {
          const int il = s_lrn[c0 + warp];
          const float* y0 = u_y + (warp * IK_PMAX) * 48;
          const int s = lane;
          const bool has_pair = s <= k_nb && s < IK_PMAX;
          float sim = 0.f;
          if (has_pair) {
            const float* ys = y0 + s * 48;
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
          for (int rr = 1; rr <= k_nb; ++rr) comm += y0[rr * 48 + lane] * __shfl_sync(0xFFFFFFFFu, wgt, rr);
          float* xr = s_X + warp * CTRM_DIM_X;
          xr[40 + lane] = comm;
          xr[8 + lane] = s_eself[(il - own_lo) * 32 + lane];
        }
        __syncthreads();
        IK_MARK(9)
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
        IK_MARK(10)
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
        IK_MARK(11)
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
          if (lane < 3) s_samples[((size_t)(il - own_lo) * bt.K + k) * 3 + lane] = o;
        }
      }
      __syncthreads();  // the samples are in place
      IK_MARK(12)
    }

    // ================= step: select + merge_samples, one warp per own agent (rollout.cuh) =================
    for (int i = own_lo + warp; i < own_hi; i += IK_WARPS) {
      const int a = a0 + i;
      StepCtx sc;
      sc.b = b; sc.t = t; sc.kt = kt; sc.makespan = makespan;
      sc.a0 = a0; sc.N = n; sc.o0 = o0; sc.nO = nO;
      sc.inv_nO = (nO >= 1 && nO <= 128) ? c_pair_inv.v[nO] : 0u;
      const double* ap = s_pack + 6 * (i - own_lo);
      sc.gx = ap[0]; sc.gy = ap[1]; sc.speed = ap[2]; sc.speed2 = ap[3]; sc.merge2 = ap[4]; sc.rad = ap[5];
      sc.sobs = nullptr;
      if (obs_static) {
        sc.sobs = s_obs + warp * 3 * IK_OBS_MAX;
      } else if (nO <= IK_OBS_MAX) {
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
      merge_prefetch<W>(bv, a, t, lane, pf);
      double lx, ly;
      unsigned int count = select_agent(bv, sc, a, lane, cx, cy, s_arrived[i] != 0, &lx, &ly);
      count += merge_agent<W>(bv, sc, a, lane, lx, ly, pf);
      if (lane == 0 && count) atomicAdd(&bt.cnt_collide[b], (unsigned long long)count);
    }
    IK_MARK(13)
    cluster.sync();
    IK_MARK(14)
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
    IK_MARK(15)
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
  if (s_vcnt != nullptr) {
    __syncthreads();
    for (int i = tid; i < n_own * bt.L; i += IK_THREADS) bt.vcnt[(size_t)(a0 + own_lo) * bt.L + i] = s_vcnt[i];
    for (int i = tid; i < n_own; i += IK_THREADS) bt.nlayers[a0 + own_lo + i] = s_nl[i];
  }
  if (prof)
    for (int i = 0; i < IK_NPROF; ++i) ia.prof[(size_t)blockIdx.x * IK_NPROF + i] = pacc[i];
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
// shared memory of one CTA at cluster size C; fills the layout-dependent fields of `ia`
static size_t instance_smem_bytes(InstArgs& ia, int n_max, int K, int L, int C) {
  auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
  ia.w0_frag = C == 8 ? 1 : 0;        // one K-slice per CTA: its mma.sync fragments stay resident
  ia.chunk = C == 8 ? 4 : IK_CHUNK;   // own agents <= 8 at C = 8: smaller passes leave room for the fragments
  const int own = (n_max + C - 1) / C;
  ia.own_max = own;
  ia.vc_smem = (size_t)own * L * 4 <= 12288 ? 1 : 0;
  const size_t u_comm = (size_t)ia.chunk * IK_PMAX * (12 + 36 + 36 + 48) * 4, u_fov = (size_t)own * 64 * 2 * 4;
  ia.u_bytes = (int)al(u_comm > u_fov ? u_comm : u_fov);
  size_t s = al((size_t)ia.n_w * 4);
  s += 2 * al(IK_NMAX * 2 * 8) + al((size_t)IK_WARPS * 3 * IK_OBS_MAX * 8) + al((size_t)IK_CHUNK * IK_NMAX * 8);
  s += al(IK_NMAX * IK_WORDS * 4) + 2 * al(IK_NMAX * 32 * 4) + al(IK_CHUNK * CTRM_DIM_X * 4) + 2 * al(IK_CHUNK * IK_PMAX * 4);
  s += 4 * al(IK_NMAX * 4) + al(IK_CHUNK * 4) + 2 * al(IK_NMAX) + al(16 * 4) + (size_t)ia.u_bytes;
  s += al(IK_NMAX * 2 * 8) + 2 * al(IK_NMAX * 4);
  if (ia.w0_frag) s += 8 * 6 * 3 * 32 * 8;
  s += al((size_t)own * 6 * 8) + al((size_t)own * K * 3 * 4) + al((size_t)own * 4);
  if (ia.vc_smem) s += al((size_t)own * L * 4);
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

template <typename Kernel>
static int launch_instance_with(Kernel kernel, SmemAttrCache& attr, const Batch& bt, InstArgs& ia, int C_req, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(IK_THREADS, 1, 1);
  cfg.stream = st;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // the largest cluster size (C_req, or 8, 4, 2, 1) at which every instance's cluster is resident at once: a second wave of
  // clusters would double the latency, which more CTAs per instance cannot win back
  {
    size_t most = 0;
    for (int c = 1; c <= 8; c <<= 1) most = std::max(most, instance_smem_bytes(ia, bt.n_max, bt.K, bt.L, c));
    if (most > 227 * 1024) return ctrm_fail(-1, "instance kernel: weights do not fit in shared memory");
    CTRM_CUDA_OK(attr.ensure(kernel, most));
  }
  int C = 1;
  for (int c = C_req > 0 ? C_req : 8; c >= 1; c >>= 1) {
    cfg.gridDim = dim3((unsigned)(bt.B * c), 1, 1);
    cfg.dynamicSmemBytes = instance_smem_bytes(ia, bt.n_max, bt.K, bt.L, c);
    at[0].val.clusterDim.x = (unsigned)c;
    int n_clusters = 0;
    CTRM_CUDA_OK(cudaOccupancyMaxActiveClusters(&n_clusters, kernel, &cfg));
    C = c;
    if (n_clusters >= bt.B || C_req > 0) break;
  }
  ia.C = C;
  cfg.gridDim = dim3((unsigned)(bt.B * C), 1, 1);
  cfg.dynamicSmemBytes = instance_smem_bytes(ia, bt.n_max, bt.K, bt.L, C);
  at[0].val.clusterDim.x = (unsigned)C;
  CTRM_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, bt, ia));
  return 0;
}

// C = 0: as many CTAs per instance (8, 4, 2, 1) as keep every cluster of the batch resident at once
int launch_instance_rollout(const DevModel& m, const Batch& bt, int nn, int C, float* d_part, cudaStream_t st) {
  if (bt.B <= 0) return 0;
  if (C != 0 && C != 1 && C != 2 && C != 4 && C != 8) return ctrm_fail(-1, "instance kernel: cluster size must be 0 (auto), 1, 2, 4 or 8");
  InstArgs ia;
  ia.blob = m.blob;
  ia.off = m.off;
  ia.F = m.desc.fov_size;
  ia.num_neighbors = m.desc.num_neighbors;
  ia.use_k_neighbor = m.desc.use_k_neighbor;
  ia.include_self = m.desc.include_self_attention;
  ia.nn = nn;
  ia.C = 1;
  ia.n_w = m.off.total - m.off.fov_b0;
  ia.w0_frag = 0;
  ia.chunk = IK_CHUNK;
  ia.u_bytes = 0;
  ia.vc_smem = 0;
  ia.own_max = bt.n_max;
  ia.part = d_part;
  ia.prof = nullptr;
  const bool want_prof = getenv("CTRM_INSTANCE_PROF") != nullptr;
  if (want_prof) CTRM_CUDA_OK(cudaMalloc(&ia.prof, (size_t)bt.B * 8 * IK_NPROF * sizeof(long long)));
  static SmemAttrCache attr[4];
  switch (bt.W) {
    case 1: CHECK_RC(launch_instance_with(instance_rollout_kernel<1>, attr[0], bt, ia, C, st)); break;
    case 2: CHECK_RC(launch_instance_with(instance_rollout_kernel<2>, attr[1], bt, ia, C, st)); break;
    case 3: CHECK_RC(launch_instance_with(instance_rollout_kernel<3>, attr[2], bt, ia, C, st)); break;
    case 4: CHECK_RC(launch_instance_with(instance_rollout_kernel<4>, attr[3], bt, ia, C, st)); break;
    default: return ctrm_fail(-1, "unsupported adjacency mask width");
  }
  if (want_prof) {  // debugging aid: cycles per phase of the first and the last CTA, thread 0's view
    const int Cu = ia.C;
    std::vector<long long> h((size_t)bt.B * Cu * IK_NPROF);
    CTRM_CUDA_OK(cudaStreamSynchronize(st));
    CTRM_CUDA_OK(cudaMemcpy(h.data(), ia.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(ia.prof);
    static const char* names[IK_NPROF] = {"gate", "raster", "fov_l0", "bar1", "fov_tail", "bar2", "vp+rank", "info", "agent_mlp",
                                          "softmax", "cvae_l0", "prior", "decode", "step", "bar3", "advance+bar4"};
    for (size_t cta : {(size_t)0, (size_t)bt.B * Cu - 1}) {
      fprintf(stderr, "[instance kernel] C=%d CTA %zu Mcycles:", Cu, cta);
      for (int i = 0; i < IK_NPROF; ++i) fprintf(stderr, " %s %.2f", names[i], 1e-6 * (double)h[cta * IK_NPROF + i]);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}
