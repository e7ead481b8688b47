// ONE THREAD-BLOCK CLUSTER PER INSTANCE: the whole roll-out loop of
// get_timed_roadmaps_multiple_paths_with_learned_indicator (roadmap_learned/learned_sampler.py:65-196) inside one kernel.
//
// The batch path (api.cu enqueue_step) runs every roll-out step as seven kernels over all instances; with a handful of
// instances each of those kernels is one 128-row tensor-core tile that is 80 % padding and costs a launch plus three to seven
// dependent MMA round trips (12-18 us per kernel, 63 us per step, 96 ms for one 28-agent instance).  Here the N_traj x max_T
// steps of an instance never leave the chip: a cluster of C = 1, 2, 4 or 8 CTAs owns the instance, every CTA keeps all network
// weights but the FOV encoders' first layer resident in shared memory (92 KB fp32), the agents are split evenly over the CTAs,
// and what the CTAs exchange travels through DISTRIBUTED SHARED MEMORY (stores into the consumer's shared memory, then one
// barrier.cluster) — a step has two cluster barriers and no global-memory round trip between its phases:
//
//   gate     which agents take the learned branch (learned_sampler.py:148-157), every CTA for all agents
//   fov L0   a warp = (own agent, quarter of the 2 F^2 inputs): it builds six 32-bit words of crop bits (cost_to_go.cpp:109-186;
//            six coalesced loads per lane and a ballot each) and, the input being BINARY, evaluates its part of the 2F^2 -> 64
//            first layers of both FOV encoders (fov_encoder.py:34-59) as TABLE LOOK-UPS: for every byte of crop bits one row of
//            a prebuilt table lut[byte position][256 patterns][64] (the sum of the weight rows of the set bits, summed in
//            float64 at ctrm_load_weights, 6 MB, L2-resident) — 24 independent 256-byte loads and adds per warp instead of a
//            K = 722 contraction; the 185 KB of first-layer weights never enter the SM
//   fov tail own agents: the four quarter sums added as a fixed tree (the result does not depend on C), 32->32->32 tails,
//            vain_proj stored into EVERY CTA's shared memory
//   -- cluster barrier 1 --
//   agent    the 16 warps form four GROUPS of four warps; a group takes one own agent at a time through
//            comm (k nearest neighbours, pair features compiled.py:30-97, AgentEncoder 43->32->32->42 on the k+1 pairs, attention
//            softmax and message sum, format_input.py:212-277), cvae (indicator + argmax, the prior for all three indicator
//            values side by side, K relaxed one-hot samples, decoder; cvae.py:57-79, :128-138; learned_sampler.py:99-121) and
//            the step (select, then merge_samples with the parents / children / merge candidates / goal test on one warp each),
//            synchronised by named barriers of 128 threads; the agent's next position and arrival flag are stored into every CTA
//   -- cluster barrier 2 --
//   advance  every CTA advances its own copy of (k_traj, t, makespan) and of all positions
//
// The MLPs run on the CUDA cores in plain fp32 (FMA, fixed summation orders): at 28 rows per step a 128-row MMA tile with a
// TMEM round trip per layer is slower than 512 threads walking 32-wide layers out of shared memory.  They agree with the
// tensor-core kernels of the batch path (fp32-exact BF16x3 / integer digits) and with the reference to ~1e-6 relative; all
// geometry, draws and roadmap updates use the device functions of the batch path (rollout.cuh) — the same bits for every fp64
// operation, verdict and counter.  Under teacher forcing the two paths are bit-identical.
#include <cooperative_groups.h>

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
#define IK_GROUPS 4    // groups of four warps: one own agent at a time through comm, cvae and the step
#define IK_NMAX 64     // agents per instance
#define IK_WORDS 24    // 32-bit words of one agent's crop bits
#define IK_PMAX 16     // pairs per agent: itself + at most 15 neighbours
#define IK_OBS_MAX 64  // obstacles staged per group in shared memory (more: read through L1)
#define IK_KMAX 8      // CVAE candidates per agent
#define IK_NPROF 16
#define IK_BYTES (4 * IK_WORDS)  // byte positions of the crop bits = rows of 256 patterns in the first-layer table

// byte offsets of the shared-memory regions (computed by the host, ik_layout)
struct IkLayout {
  int cur, prev, goal, nxt, vp, eself, l0p, radf, speedf, flags, pack, samples, nl, vcnt, obs, prof, pre, wvf, U, total;
};

struct InstArgs {
  const float* blob;
  const float* lut;    // [IK_BYTES][256][64] FOV first layer as a table over the bytes of the crop bits (instance_w0_image_kernel)
  ctrm_weight_offsets_t off;
  int F, num_neighbors, use_k_neighbor, include_self;
  int nn;        // run the networks (0: teacher-forced run that does not record the CVAE output)
  int C;         // CTAs per instance
  int n_w;       // floats of the blob from fov_b0 on (resident in shared memory)
  int l0_split;  // a warp takes a quarter of an agent's inputs (few agents per CTA) instead of a whole agent
  int pre_rank;  // the neighbours of the own learned agents are ranked beside the FOV tails (PreAgent records)
  int vc_smem;   // the own agents' layer counts live in shared memory during the roll-outs
  int own_max;   // ceil(n_max / C)
  IkLayout lay;
  long long* prof;  // optional [B * C][IK_NPROF] clock cycles per phase (CTRM_INSTANCE_PROF=1)
};

// neighbour order of one own agent, ranked beside the FOV tails by a warp that has none (it depends on positions only)
struct __align__(16) PreAgent {
  unsigned long long key[IK_NMAX];
  int order[IK_PMAX];
};

// scratch of one group (one agent at a time)
struct __align__(16) GroupScratch {
  float info[IK_PMAX * 12];  // pair features, K padded to 12
  float hA[IK_PMAX * 36];
  float hB[IK_PMAX * 36];
  float y[IK_PMAX * 48];     // [message 32 | attention 10 | pad]
  float X[CTRM_DIM_X];       // [self_info 8 | e_self 32 | comm 32]
  float p0[96];              // first layers of indicator | encoder_input | decoder (x rows), without bias
  float sp[3 * 64];          // sqrt of the clamped prior probabilities, one row per indicator value
  float d0[3 * 32];          // x part of the decoder's first layer, one row per indicator value
  unsigned long long keys[IK_NMAX];
  double lx, ly;             // loc_pred of the step
  uint32_t P[CTRM_MAX_W], Cm[CTRM_MAX_W];
  int order[IK_PMAX];
  int pj[IK_PMAX];
  int ind;
  int spec_tested, spec_arr;  // goal test of (lx, ly), evaluated beside the parent / child scans
  int pad;
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
#pragma unroll 8
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
#pragma unroll 2
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

// four output columns 4q .. 4q + 3 of a 32-input layer row; the inputs are cut into four strands kq that meet by shuffle: chains
// of 8 instead of 32.  Lane layout inside a warp: q = lane & 7, kq = (lane >> 3) & 3 — the eight lanes of a quarter-warp read
// eight different 16-byte columns of ONE weight row (a conflict-free 128-byte wavefront) and the same input (a broadcast);
// strands on adjacent lanes put four rows on the same banks (4-way conflicts, measured as the bound of these phases).
// Whole warps must call this together.
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
  for (int d = 8; d <= 16; d <<= 1) {
    acc.x += __shfl_xor_sync(0xFFFFFFFFu, acc.x, d); acc.y += __shfl_xor_sync(0xFFFFFFFFu, acc.y, d);
    acc.z += __shfl_xor_sync(0xFFFFFFFFu, acc.z, d); acc.w += __shfl_xor_sync(0xFFFFFFFFu, acc.w, d);
  }
  return acc;
}

// a 32 -> 4 head (W: [32][4], the last column is padding): lane = (output lane & 3, strand of four inputs lane >> 2); every
// lane returns bias + sum for its output.  Whole warps call this together.
__device__ __forceinline__ float head4(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, int lane) {
  const int col = lane & 3, st = lane >> 2;
  const float4 xv = *reinterpret_cast<const float4*>(x + 4 * st);
  float acc = xv.x * W[(4 * st) * 4 + col];
  acc = fmaf(xv.y, W[(4 * st + 1) * 4 + col], acc);
  acc = fmaf(xv.z, W[(4 * st + 2) * 4 + col], acc);
  acc = fmaf(xv.w, W[(4 * st + 3) * 4 + col], acc);
#pragma unroll
  for (int d = 4; d <= 16; d <<= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  return acc + bias[col];
}

// The FOV encoders' first layer as a table over the BYTES of the binary input: lut[p][v][c] = sum over the set bits i of v of
// W0[8 p + i][c] (rows past 2 F^2 count as zero), summed in float64 and rounded once.
__global__ void instance_w0_image_kernel(const float* __restrict__ W0, int KD, float* __restrict__ lut) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= IK_BYTES * 256 * 64) return;
  const int c = gidx & 63, v = (gidx >> 6) & 255, p = gidx >> 14;
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = 8 * p + i;
    if (((v >> i) & 1) && j < KD) acc += (double)W0[(size_t)j * 64 + c];
  }
  lut[gidx] = (float)acc;
}

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// valid_edge(loc, goal) of learned_sampler.py:182-183 for one warp: bit 0 = the swept test ran (counted by the reference's
// collision counter), bit 1 = the agent arrives
__device__ __noinline__ int goal_test(const Batch& bt, const StepCtx& sc, double x, double y, int lane) {
  if (!(bounds_ok(sc.gx, sc.gy, sc.rad) && speed_ok(x, y, sc.gx, sc.gy, sc.speed, sc.speed2))) return 0;
  bool hit = false;
  for (int o = lane; o < sc.nO; o += 32) {
    double ox, oy, rs;
    obstacle_at(bt, sc, o, &ox, &oy, &rs);
    hit = hit || (!sweep_far(x, y, sc.gx, sc.gy, ox, oy, rs) && sweep_static(x, y, sc.gx, sc.gy, ox, oy, rs));
  }
  return __any_sync(0xFFFFFFFFu, hit) ? 1 : 3;
}

// WIDE: select + merge_samples as one warp per agent (many agents per CTA: one or two CTAs per instance) instead of one group
// of four warps per agent.  Both are the same fp64 operations: the roadmaps do not depend on it.
template <int W, bool WIDE>
__global__ void __launch_bounds__(IK_THREADS, 1) instance_rollout_kernel(Batch bt, InstArgs ia) {
  extern __shared__ __align__(16) uint8_t ik_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = ia.C;
  const int b = blockIdx.x / C, r = blockIdx.x - b * C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = warp >> 2, gt = tid & 127;
  // role of the warp inside its group, rotated by the group number: warp w runs on scheduler w & 3, and the same role of all
  // four groups (the select warp, the resolving warp ...) would otherwise share one scheduler while three idle
  const int gw = ((warp & 3) + g) & 3;
  const int a0 = bt.agent_off[b], n = bt.agent_off[b + 1] - a0;
  const int o0 = bt.obs_off[b], nO = bt.obs_off[b + 1] - o0;
  const int M = bt.M, F = ia.F, FF = F * F, KD = 2 * FF, buf = (F - 1) / 2;
  const int own_lo = (r * n) / C, own_hi = ((r + 1) * n) / C, n_own = own_hi - own_lo;  // even split, sizes differ by at most one
  const int own_max = ia.own_max;
  const ctrm_weight_offsets_t& off = ia.off;
  const IkLayout& ly_ = ia.lay;

  // ---- shared memory
  float* Wsm = reinterpret_cast<float*>(ik_smem);  // blob[fov_b0 .. total)
  double* s_cur = reinterpret_cast<double*>(ik_smem + ly_.cur);
  double* s_prev = reinterpret_cast<double*>(ik_smem + ly_.prev);
  double* s_goal = reinterpret_cast<double*>(ik_smem + ly_.goal);
  double* s_nxt = reinterpret_cast<double*>(ik_smem + ly_.nxt);      // [2][IK_NMAX][2]: stored by the owners, by step parity
  float* s_vp = reinterpret_cast<float*>(ik_smem + ly_.vp);           // [IK_NMAX][32] vain_proj of all agents
  float* s_eself = reinterpret_cast<float*>(ik_smem + ly_.eself);     // [own][32]
  float* s_l0p = reinterpret_cast<float*>(ik_smem + ly_.l0p);         // first-layer sums: [own][4 quarters][64] (l0_split) or [own][64]
  float* s_radf = reinterpret_cast<float*>(ik_smem + ly_.radf);
  float* s_speedf = reinterpret_cast<float*>(ik_smem + ly_.speedf);
  uint8_t* s_arrived = ik_smem + ly_.flags;             // [IK_NMAX]
  uint8_t* s_learned = s_arrived + IK_NMAX;             // [IK_NMAX]
  uint8_t* s_arrn = s_learned + IK_NMAX;                // [2][IK_NMAX] arrival flags after the step, stored by the owners
  double* s_pack = reinterpret_cast<double*>(ik_smem + ly_.pack);      // own agents' constants (Batch::agent_pack)
  float* s_samples = reinterpret_cast<float*>(ik_smem + ly_.samples);  // own agents' CVAE samples of the step
  int32_t* s_nl = reinterpret_cast<int32_t*>(ik_smem + ly_.nl);
  int32_t* s_vcnt = ia.vc_smem ? reinterpret_cast<int32_t*>(ik_smem + ly_.vcnt) : nullptr;
  double* s_obs = reinterpret_cast<double*>(ik_smem + ly_.obs);        // [IK_GROUPS][3 * IK_OBS_MAX]
  PreAgent* s_pre = reinterpret_cast<PreAgent*>(ik_smem + ly_.pre);  // [own_max] when ia.pre_rank
  float* s_wvf = reinterpret_cast<float*>(ik_smem + ly_.wvf);  // [32][32] + [32]: vain tail's last layer fused with agent.mlp.0's vain rows
  uint8_t* U = ik_smem + ly_.U;
  GroupScratch& gs = reinterpret_cast<GroupScratch*>(U)[g];  // agent phases
  float* u_h0 = reinterpret_cast<float*>(U);                 // fov tail: [own agents][64]
  float* u_h1 = u_h0 + own_max * 64;                         //           [own agents][64]
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
  __syncthreads();
  // vain_proj = b0 + W0v^T (W2v^T h + b2v) (fov_encoder.py:34-59, agent_encoder.py:57-60) has no non-linearity between its two
  // layers: one 32 x 32 layer with the product matrix (summed in float64) instead of two in a row on every step's critical path
  for (int idx = tid; idx < 33 * 32; idx += IK_THREADS) {
    const int k = idx >> 5, c2 = idx & 31;  // k = 32: the bias row
    double acc = k == 32 ? (double)Wp(off.ag_b0)[c2] : 0.0;
    for (int c = 0; c < 32; ++c)
      acc += (double)(k == 32 ? Wp(off.fovv_b2)[c] : Wp(off.fovv_w2)[k * 32 + c]) * (double)Wp(off.ag_w0v)[c * 32 + c2];
    s_wvf[idx] = (float)acc;
  }
  __syncthreads();

  int t = 1, kt = 0, makespan = 0, steps = 0;
  const bool all_rows = bt.tr_samples != nullptr;  // trace: the CVAE output of every agent-step is recorded
  const int k_nb = ia.use_k_neighbor ? min(ia.num_neighbors, n - 1) : n - 1;
  const float eps = 1.1920928955078125e-07f;
  const bool run = bt.N_traj > 0 && n > 0;
  unsigned int cnt_acc = 0;  // collide_continuous_sphere calls of this warp (warp-uniform)
  long long* s_prof = reinterpret_cast<long long*>(ik_smem + ly_.prof);  // cycles per phase, thread 0's view
  long long pt0 = 0;
  const bool prof = ia.prof != nullptr && tid == 0;
  if (prof) {
    if (tid == 0)
      for (int i = 0; i < IK_NPROF; ++i) s_prof[i] = 0;
    pt0 = clock64();
  }
#define IK_MARK(k)                      \
  if (prof) {                           \
    const long long now = clock64();    \
    s_prof[k] += now - pt0;             \
    pt0 = now;                          \
  }
#define IK_CLUSTER_SYNC()            \
  do {                               \
    if (C > 1) cluster.sync();       \
    else __syncthreads();            \
  } while (0)

  // the batch kernels' device functions index per-agent state by the global agent number: hand them a view of the batch whose
  // hot per-agent arrays are the shared-memory copies of the own agents
  Batch bv = bt;
  bv.samples = s_samples - (size_t)(a0 + own_lo) * bt.K * 3;
  if (s_vcnt != nullptr) {
    bv.vcnt = s_vcnt - (size_t)(a0 + own_lo) * bt.L;
    bv.nlayers = s_nl - (size_t)(a0 + own_lo);
  }
  // obstacles of the group's agent as (x, y, rad_agent + rad_obstacle); a group that serves one agent stages them once
  const bool obs_smem = nO <= IK_OBS_MAX;
  constexpr bool wide_step = WIDE;
  const bool obs_static = !wide_step && n_own <= IK_GROUPS && obs_smem;
  double* g_obs = s_obs + g * 3 * IK_OBS_MAX;
  auto stage_obstacles = [&](double rad) {
    for (int o = gt; o < nO; o += 128) {
      const double* gp = bt.obs_xyr + 3 * (size_t)(o0 + o);
      g_obs[3 * o] = __ldg(gp);
      g_obs[3 * o + 1] = __ldg(gp + 1);
      g_obs[3 * o + 2] = f64add(rad, __ldg(gp + 2));  // rad1 + rad2, rounded once (collision_check.cpp:56)
    }
  };
  if (obs_static && g < n_own) stage_obstacles(s_pack[6 * g + 5]);
  cluster.sync();  // every CTA of the cluster is running before anyone stores into a neighbour's shared memory

  while (run) {
    // ---- which agents take the learned branch in this step (learned_sampler.py:148-157): every CTA for all agents
    bool lrn = false;
    if (tid < n) {
      const bool gate = gate_is_learned(bt, b, a0 + tid, tid, kt, t, makespan, s_arrived[tid] != 0);
      lrn = all_rows || gate;
      s_learned[tid] = (lrn ? 1 : 0) | (gate ? 2 : 0);  // bit 0: run the networks for this agent, bit 1: the gate itself
    }
    const bool any_learned = __syncthreads_or(lrn ? 1 : 0) != 0;  // also publishes the positions of the previous advance
    IK_MARK(0)

    if (ia.nn && any_learned) {
      // ================= FOV crops and first layers of the OWN agents.  A warp builds six consecutive 32-bit words of an agent's
      // crop bits: lane l reads input 32 w + l of word w, so a load instruction covers 32 consecutive crop cells (two or three map
      // rows), and a ballot is the word (the centre's cost is loaded beside the crop: whether it is inside the map is known
      // without it).  The same warp then adds up the table rows of the words' 24 bytes, lane = two of the 64 columns: 24
      // independent 8-byte loads per lane, summed in byte order.  Quarter sums q0..q3 are combined as (q0 + q1) + (q2 + q3), by the
      // tail (l0_split: a warp per quarter) or by the warp itself (a warp per agent), so the result does not depend on the split.
      {
        const uint8_t* occ = bt.occ + (size_t)b * M * M;
        const unsigned invF = 65536u / (unsigned)F + 1u;  // jj / F == (jj * invF) >> 16 for jj < F * F, F <= 41
        const int n_bytes = (KD + 7) >> 3;                // byte positions that hold inputs
        const int split = ia.l0_split;
        for (int task = warp; task < (split ? 4 * n_own : n_own); task += IK_WARPS) {
          const int io = split ? task >> 2 : task, i = own_lo + io;
          const int cx0 = grid_cell(s_cur[2 * i], M), cy0 = grid_cell(s_cur[2 * i + 1], M);
          const bool centre_in = cx0 < M && cy0 < M;
          const uint16_t* cf = bt.ctg + (size_t)(a0 + i) * ctg_field_elems(M, 1);
          const int cc = centre_in ? (int)__ldg(&cf[ctg_index(cx0, cy0, M, 1)]) : -1;
          const int cx = cx0 - buf, cy = cy0 - buf;
          float2 qs[4];
#pragma unroll 1
          for (int oct = split ? (task & 3) : 0; oct < (split ? (task & 3) + 1 : 4); ++oct) {
            int vc[6], vo[6];
            bool isc[6];
#pragma unroll
            for (int e = 0; e < 6; ++e) {
              const int j = 32 * (6 * oct + e) + lane;
              const int ch = j >= FF ? 1 : 0, jj = j - ch * FF;
              const int xf = (int)(((unsigned)jj * invF) >> 16), yf = jj - xf * F;
              const int x = cx + xf, y = cy + yf;
              const bool ok = j < KD && centre_in && x >= 0 && x < M && y >= 0 && y < M;
              isc[e] = ch != 0;
              vc[e] = CTRM_UNREACH;
              vo[e] = 1;
              // a word lies in one channel except the one that straddles F^2: warp-uniform tests skip the other channel's index math
              if (32 * (6 * oct + e) + 31 >= FF) vc[e] = (ok && ch) ? (int)__ldg(&cf[ctg_index(x, y, M, 1)]) : CTRM_UNREACH;
              if (32 * (6 * oct + e) < FF) vo[e] = (ok && !ch) ? (int)__ldg(&occ[x * M + y]) : 1;
            }
            uint32_t mine = 0u;
#pragma unroll
            for (int e = 0; e < 6; ++e) {
              const bool bit = isc[e] ? (vc[e] != CTRM_UNREACH && (cc == CTRM_UNREACH || vc[e] < cc)) : (vo[e] == 0);
              const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
              if (lane == e) mine = word;
            }
            // table rows of this quarter's bytes (byte positions 24 oct .. 24 oct + 23)
            float2 v[24];
            const float2* lut = reinterpret_cast<const float2*>(ia.lut) + lane;
#pragma unroll
            for (int gb = 0; gb < 24; ++gb) {
              const uint32_t word = __shfl_sync(0xFFFFFFFFu, mine, gb >> 2);
              const int pos = 24 * oct + gb;
              v[gb] = pos < n_bytes ? __ldg(lut + ((size_t)pos * 256 + ((word >> (8 * (gb & 3))) & 0xFFu)) * 32) : make_float2(0.f, 0.f);
            }
            float2 acc = v[0];
#pragma unroll
            for (int gb = 1; gb < 24; ++gb) { acc.x += v[gb].x; acc.y += v[gb].y; }
            if (split) {
              *reinterpret_cast<float2*>(s_l0p + ((size_t)(io * 4 + oct)) * 64 + 2 * lane) = acc;
            } else {
              if (oct == 0) qs[0] = acc; else if (oct == 1) qs[1] = acc; else if (oct == 2) qs[2] = acc; else qs[3] = acc;
            }
          }
          if (!split)
            *reinterpret_cast<float2*>(s_l0p + (size_t)io * 64 + 2 * lane) =
                make_float2((qs[0].x + qs[1].x) + (qs[2].x + qs[3].x), (qs[0].y + qs[1].y) + (qs[2].y + qs[3].y));
        }
      }
      IK_MARK(1)
      __syncthreads();  // the first-layer sums of the own agents are in place
      IK_MARK(3)
      // ================= fov tails of the own agents: warp = (own agent, encoder), the whole chain quarter sums -> hidden 1 ->
      // hidden 2 -> embedding (-> vain_proj) inside the warp with __syncwarp between the layers — no block-wide barrier
      {
        const float* b0 = Wp(off.fov_b0);
        const int q = lane & 7, kq = lane >> 3;
        for (int task = warp; task < (ia.pre_rank ? 3 : 2) * n_own; task += IK_WARPS) {
          if (task >= 2 * n_own) {
            // ---- a warp without a tail: k nearest neighbours of an own learned agent (format_input.py:239-243:
            //      float32(sqrt_f64(dx^2 + dy^2)) keys, ties by index, self first)
            const int pio = task - 2 * n_own, i = own_lo + pio;
            if (!(s_learned[i] & 1)) continue;
            PreAgent& pa = s_pre[pio];
            const double cix = s_cur[2 * i], ciy = s_cur[2 * i + 1];
            for (int j = lane; j < n; j += 32) {
              const double dx = f64sub(s_cur[2 * j], cix), dy = f64sub(s_cur[2 * j + 1], ciy);
              const float d = __double2float_rn(__dsqrt_rn(f64add(f64mul(dx, dx), f64mul(dy, dy))));
              pa.key[j] = (j == i) ? 0ull : (((unsigned long long)(__float_as_uint(d) + 1u) << 32) | (unsigned)j);
            }
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
              const unsigned long long kj = pa.key[j];
              int rank = 0;
              for (int j2 = 0; j2 < n; ++j2) rank += pa.key[j2] < kj ? 1 : 0;
              if (rank <= k_nb) pa.order[rank] = j;
            }
            continue;
          }
          const int io = task >> 1, e = task & 1;
          float* h0 = u_h0 + io * 64 + 32 * e;
          float* h1 = u_h1 + io * 64 + 32 * e;
          {
            const int o = 32 * e + lane;
            float sum;
            if (ia.l0_split) {
              const float* pq = s_l0p + (size_t)io * 256 + o;
              sum = (pq[0] + pq[64]) + (pq[128] + pq[192]);
            } else {
              sum = s_l0p[(size_t)io * 64 + o];
            }
            h0[lane] = fmaxf(b0[o] + sum, 0.f);
          }
          __syncwarp();
          {  // hidden 1 -> hidden 2
            const float4 acc = splitk_32(h0, Wp(e ? off.fovv_w1 : off.fovs_w1), q, kq);
            if (kq == 0) {
              const float4 bb = *reinterpret_cast<const float4*>(Wp(e ? off.fovv_b1 : off.fovs_b1) + 4 * q);
              *reinterpret_cast<float4*>(h1 + 4 * q) = make_float4(fmaxf(acc.x + bb.x, 0.f), fmaxf(acc.y + bb.y, 0.f),
                                                                  fmaxf(acc.z + bb.z, 0.f), fmaxf(acc.w + bb.w, 0.f));
            }
          }
          __syncwarp();
          {  // hidden 2 -> e_self (shared memory) | vain_proj = b0 + W0[11:43]^T e_vain (agent_encoder.py:57-60) through the fused
             // matrix, stored into every CTA of the cluster (strand kq -> CTAs kq and kq + 4: the four strands hold the same sum
             // after the shuffles)
            const float4 acc = splitk_32(h1, e ? s_wvf : Wp(off.fovs_w2), q, kq);
            const float4 bb = *reinterpret_cast<const float4*>((e ? s_wvf + 32 * 32 : Wp(off.fovs_b2)) + 4 * q);
            const float4 v = make_float4(acc.x + bb.x, acc.y + bb.y, acc.z + bb.z, acc.w + bb.w);
            if (e) {
              for (int c = kq; c < C; c += 4)
                *reinterpret_cast<float4*>(cluster.map_shared_rank(s_vp, (unsigned)c) + (size_t)(own_lo + io) * 32 + 4 * q) = v;
            } else if (kq == 0) {
              *reinterpret_cast<float4*>(s_eself + io * 32 + 4 * q) = v;
            }
          }
        }
      }
      IK_MARK(4)
      IK_CLUSTER_SYNC();
      IK_MARK(5)
    }

    // ================= the own agents, one per group at a time: comm -> cvae -> select -> merge =================
    const int pb = steps & 1;
    for (int io = g; io < n_own; io += IK_GROUPS) {
      const int i = own_lo + io, a = a0 + i;
      const bool learned = (s_learned[i] & 1) != 0;
      const double* ap = s_pack + 6 * io;
      if (!wide_step && !obs_static && obs_smem) {
        stage_obstacles(ap[5]);
        group_bar(g);
      }
      if (ia.nn && learned) {
        // ---- k nearest neighbours (format_input.py:239-243): float32(sqrt_f64(dx^2 + dy^2)) keys, ties by index, self first
        const double cix = s_cur[2 * i], ciy = s_cur[2 * i + 1];
        const int* p_order = ia.pre_rank ? s_pre[io].order : gs.order;
        if (!ia.pre_rank) {
        if (gt < n) {
          const int j = gt;
          const double dx = f64sub(s_cur[2 * j], cix), dy = f64sub(s_cur[2 * j + 1], ciy);
          const float d = __double2float_rn(__dsqrt_rn(f64add(f64mul(dx, dx), f64mul(dy, dy))));
          gs.keys[j] = (j == i) ? 0ull : (((unsigned long long)(__float_as_uint(d) + 1u) << 32) | (unsigned)j);
        }
        group_bar(g);
        {  // thread = (agent j, a part of the comparison range); the parts meet by shuffle
          const int sh = n <= 32 ? 2 : 1, parts = 1 << sh;
          const int j = gt >> sh, part = gt & (parts - 1);
          const int len = (n + parts - 1) >> sh;
          const int lo = part * len, hi = min(n, lo + len);
          const bool act = j < n;
          const unsigned long long kj = act ? gs.keys[j] : 0ull;
          int rank = 0;
          for (int j2 = lo; j2 < hi; ++j2) rank += gs.keys[j2] < kj ? 1 : 0;
          rank += __shfl_xor_sync(0xFFFFFFFFu, rank, 1);
          if (sh == 2) rank += __shfl_xor_sync(0xFFFFFFFFu, rank, 2);
          if (act && part == 0 && rank <= k_nb) gs.order[rank] = j;
        }
        group_bar(g);
        }
        IK_MARK(6)
        if (prof) s_prof[2] += 1000000;  // occurrences of the learned chain on thread 0's agent (printed in millions)
        // ---- a warp takes four (agent, slot) pair rows through the pair features (compiled.py:30-97; slot 0 = the agent
        //      itself, slot s = its s-th nearest) and the three layers of AgentEncoder (agent_encoder.py:34-65), lane = column
        {
          const int p0 = 4 * gw;
          if (lane < 4) {
            const int s = p0 + lane;
            float* info = gs.info + s * 12;
            if (s <= k_nb) {
              const int j = p_order[s];
              ik_nvm(f64sub(s_cur[2 * j], cix), f64sub(s_cur[2 * j + 1], ciy), info + 0);
              ik_nvm(f64sub(s_goal[2 * j], cix), f64sub(s_goal[2 * j + 1], ciy), info + 3);
              ik_nvm(f64sub(s_prev[2 * j], cix), f64sub(s_prev[2 * j + 1], ciy), info + 6);
              info[9] = s_radf[j];
              info[10] = s_speedf[j];
              info[11] = 0.f;  // K padded to 12 (the weight row it meets is finite)
              gs.pj[s] = j;
              if (s == 0) {  // get_self_info (compiled.py:100-125): row (i, i), columns 3..10
#pragma unroll
                for (int k = 0; k < 8; ++k) gs.X[k] = info[3 + k];
              }
            } else {
#pragma unroll
              for (int k = 0; k < 12; ++k) info[k] = 0.f;
              gs.pj[s] = -1;
            }
          }
          __syncwarp();
          const float* brow[4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) brow[rr] = gs.pj[p0 + rr] >= 0 ? s_vp + gs.pj[p0 + rr] * 32 : nullptr;
          warp_rows4<12, 0, true>(gs.info, 12, Wp(off.ag_w0i), 32, brow, gs.hA, 36, p0, lane);
          __syncwarp();
          const float* b1 = Wp(off.ag_b1);
          const float* brow1[4] = {b1, b1, b1, b1};
          warp_rows4<32, 0, true>(gs.hA, 36, Wp(off.ag_w1), 32, brow1, gs.hB, 36, p0, lane);
          __syncwarp();
          const float* b2 = Wp(off.ag_b2);
          const float* brow2[4] = {b2, b2, b2, b2};
          warp_rows4<32, 12, false>(gs.hB, 36, Wp(off.ag_w2), 44, brow2, gs.y, 48, p0, lane);
        }
        group_bar(g);
        IK_MARK(7)
        // ---- attention softmax over the neighbours and message sum (format_input.py:250-275); X = [self_info | e_self | comm]
        //      on role 0; the other roles meanwhile draw the Gumbel noise of their decoder copy (it depends on the step's
        //      counters only): rsqrt(E) with E = -log u for latents 2 lane and 2 lane + 1
        const int dk0 = (gw + 3) & 3;  // the decoder copies of this role: dk0, dk0 + 4, ...
        float rsa = 0.f, rsb = 0.f;
        auto gumbel_rsqrt = [&](int k, float* pa, float* pb) {
          const uint4 w4 = philox4x32(bt.inst_base + (uint32_t)b, rng_ctr1(kt, t), (uint32_t)(k * n + i),
                                      rng_ctr3(STREAM_GUMBEL, (uint32_t)(lane >> 1)), bt.seed_lo, bt.seed_hi);
          const float ua = u32_to_f32((lane & 1) ? w4.z : w4.x), ub = u32_to_f32((lane & 1) ? w4.w : w4.y);
          *pa = rsqrtf(-logf(fminf(fmaxf(ua, eps), 1.f - eps)));
          *pb = rsqrtf(-logf(fminf(fmaxf(ub, eps), 1.f - eps)));
        };
        if (gw != 0 && dk0 < bt.K) {
          gumbel_rsqrt(dk0, &rsa, &rsb);
          asm volatile("" : "+f"(rsa), "+f"(rsb));  // materialise here, beside the softmax: the compiler would sink it to its use
        }
        // ... and start the first layers of indicator | encoder_input | decoder (x rows): role = net + 1, lane = (k strand, four
        // columns) in the conflict-free layout of splitk_32.  Inputs 0..39 of X (self_info, e_self) do not depend on the attention:
        // their part is summed here, the 32 comm inputs follow once the softmax warp has written them
        const int l0q = lane & 7, l0kq = lane >> 3;
        const float* l0W = gw == 1 ? Wp(off.ind_w0) : (gw == 2 ? Wp(off.enc_w0) : Wp(off.dec_w0) + 64 * 32);
        float4 l0acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gw != 0) {
#pragma unroll 5
          for (int kk = 0; kk < 10; ++kk) {
            const int k = 10 * l0kq + kk;
            const float4 w = *reinterpret_cast<const float4*>(l0W + k * 32 + 4 * l0q);
            const float v = k < 8 ? gs.X[k] : s_eself[io * 32 + k - 8];
            l0acc.x = fmaf(v, w.x, l0acc.x); l0acc.y = fmaf(v, w.y, l0acc.y); l0acc.z = fmaf(v, w.z, l0acc.z); l0acc.w = fmaf(v, w.w, l0acc.w);
          }
          asm volatile("" : "+f"(l0acc.x), "+f"(l0acc.y), "+f"(l0acc.z), "+f"(l0acc.w));
        }
        if (gw == 0) {
          const float* y0 = gs.y;
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
          gs.X[40 + lane] = comm;
          gs.X[8 + lane] = s_eself[io * 32 + lane];
        }
        group_bar(g);
        IK_MARK(8)
        // ---- the comm part of the first layers (inputs 40..71), the strands meet by shuffle
        if (gw != 0) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int k = 40 + 8 * l0kq + kk;
            const float4 w = *reinterpret_cast<const float4*>(l0W + k * 32 + 4 * l0q);
            const float v = gs.X[k];
            l0acc.x = fmaf(v, w.x, l0acc.x); l0acc.y = fmaf(v, w.y, l0acc.y); l0acc.z = fmaf(v, w.z, l0acc.z); l0acc.w = fmaf(v, w.w, l0acc.w);
          }
#pragma unroll
          for (int d = 8; d <= 16; d <<= 1) {
            l0acc.x += __shfl_xor_sync(0xFFFFFFFFu, l0acc.x, d); l0acc.y += __shfl_xor_sync(0xFFFFFFFFu, l0acc.y, d);
            l0acc.z += __shfl_xor_sync(0xFFFFFFFFu, l0acc.z, d); l0acc.w += __shfl_xor_sync(0xFFFFFFFFu, l0acc.w, d);
          }
          if (l0kq == 0) *reinterpret_cast<float4*>(gs.p0 + 32 * (gw - 1) + 4 * l0q) = l0acc;
        }
        group_bar(g);
        IK_MARK(9)
        // ---- role 0: indicator -> argmax (learned_sampler.py:99-104).  roles 1..3: the prior for indicator value v = role - 1
        //      (cvae.py:57-68, :132-136; sqrt of the clamped probabilities, cvae_tc.cu) and the decoder's first layer without its
        //      z rows — the three values side by side instead of waiting for the argmax.  Every layer in split-K form (strands
        //      of 8 or 16 inputs that meet by shuffle, vectors handed on through shared memory): chains of 8-16 FMAs instead of
        //      32-64, and a third fewer instructions than a lane-per-column layer fed by shuffles.
        float* vA = gs.hA + gw * 128;  // this warp's vectors (the pair MLP's scratch is free by now): [32] in, [32] hidden, [64] z
        float* vB = vA + 32;
        float* vZ = vA + 64;
        const int sq = lane & 7, skq = lane >> 3;  // splitk_32 layout
        if (gw == 0) {
          vA[lane] = fmaxf(gs.p0[lane] + Wp(off.ind_b0)[lane], 0.f);
          __syncwarp();
          const float4 h = splitk_32(vA, Wp(off.ind_w1), sq, skq);
          if (skq == 0) {
            const float4 bb = *reinterpret_cast<const float4*>(Wp(off.ind_b1) + 4 * sq);
            *reinterpret_cast<float4*>(vB + 4 * sq) = make_float4(fmaxf(h.x + bb.x, 0.f), fmaxf(h.y + bb.y, 0.f), fmaxf(h.z + bb.z, 0.f), fmaxf(h.w + bb.w, 0.f));
          }
          __syncwarp();
          const float lg = head4(vB, Wp(off.ind_w2), Wp(off.ind_b2), lane);
          const float l0 = __shfl_sync(0xFFFFFFFFu, lg, 0), l1 = __shfl_sync(0xFFFFFFFFu, lg, 1), l2 = __shfl_sync(0xFFFFFFFFu, lg, 2);
          int ind = 0;
          float best = l0;
          if (l1 > best) { best = l1; ind = 1; }
          if (l2 > best) { best = l2; ind = 2; }
          if (lane == 0) gs.ind = ind;
        } else {
          const int v = gw - 1;
          gs.d0[v * 32 + lane] = (gs.p0[64 + lane] + Wp(off.dec_b0)[lane]) + Wp(off.dec_w0)[(64 + CTRM_DIM_X + v) * 32 + lane];
          vA[lane] = fmaxf((gs.p0[32 + lane] + Wp(off.enc_b0)[lane]) + Wp(off.enc_w0)[(CTRM_DIM_X + v) * 32 + lane], 0.f);
          __syncwarp();
          const float4 h = splitk_32(vA, Wp(off.enc_w1), sq, skq);
          if (skq == 0) {
            const float4 bb = *reinterpret_cast<const float4*>(Wp(off.enc_b1) + 4 * sq);
            *reinterpret_cast<float4*>(vB + 4 * sq) = make_float4(fmaxf(h.x + bb.x, 0.f), fmaxf(h.y + bb.y, 0.f), fmaxf(h.z + bb.z, 0.f), fmaxf(h.w + bb.w, 0.f));
          }
          __syncwarp();
          // 32 -> 64 logits: lane = (four columns q16, strand of 16 inputs); both half-warps end up with all 64 logits
          const int q16 = lane & 15, s2 = lane >> 4;
          float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
          {
            const float* W2 = Wp(off.enc_w2);
#pragma unroll 8
            for (int kk = 0; kk < 16; ++kk) {
              const int k = 16 * s2 + kk;
              const float4 w = *reinterpret_cast<const float4*>(W2 + k * 64 + 4 * q16);
              const float x = vB[k];
              gq.x = fmaf(x, w.x, gq.x); gq.y = fmaf(x, w.y, gq.y); gq.z = fmaf(x, w.z, gq.z); gq.w = fmaf(x, w.w, gq.w);
            }
            gq.x += __shfl_xor_sync(0xFFFFFFFFu, gq.x, 16); gq.y += __shfl_xor_sync(0xFFFFFFFFu, gq.y, 16);
            gq.z += __shfl_xor_sync(0xFFFFFFFFu, gq.z, 16); gq.w += __shfl_xor_sync(0xFFFFFFFFu, gq.w, 16);
            const float4 bb = *reinterpret_cast<const float4*>(Wp(off.enc_b2) + 4 * q16);
            gq.x += bb.x; gq.y += bb.y; gq.z += bb.z; gq.w += bb.w;
          }
          // log_softmax -> probs -> Categorical's normalisation -> sqrt of the clamped probabilities; reductions over 16 lanes
          auto red16_max = [&](float x) {
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) x = fmaxf(x, __shfl_xor_sync(0xFFFFFFFFu, x, d));
            return x;
          };
          auto red16_sum = [&](float x) {
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
            return x;
          };
          const float mx = red16_max(fmaxf(fmaxf(gq.x, gq.y), fmaxf(gq.z, gq.w)));
          const float lse = logf(red16_sum((expf(gq.x - mx) + expf(gq.y - mx)) + (expf(gq.z - mx) + expf(gq.w - mx))));
          float4 pq = make_float4(expf((gq.x - mx) - lse), expf((gq.y - mx) - lse), expf((gq.z - mx) - lse), expf((gq.w - mx) - lse));
          const float ps = red16_sum((pq.x + pq.y) + (pq.z + pq.w));
          if (s2 == 0)
            *reinterpret_cast<float4*>(gs.sp + v * 64 + 4 * q16) =
                make_float4(sqrtf(fminf(fmaxf(pq.x / ps, eps), 1.f - eps)), sqrtf(fminf(fmaxf(pq.y / ps, eps), 1.f - eps)),
                            sqrtf(fminf(fmaxf(pq.z / ps, eps), 1.f - eps)), sqrtf(fminf(fmaxf(pq.w / ps, eps), 1.f - eps)));
        }
        group_bar(g);
        IK_MARK(10)
        // ---- K relaxed one-hot samples and the decoder, one warp per copy: z_i = sqrt(p_i / E_i) / sum (temperature 2,
        //      cvae_tc.cu), lane holds latents 2 lane and 2 lane + 1; decoder layers in split-K form
        for (int k = dk0; k < bt.K; k += 4) {
          const int ind = gs.ind;
          const float* sp = gs.sp + ind * 64;
          if (gw == 0 || k != dk0) gumbel_rsqrt(k, &rsa, &rsb);
          const float2 sp2 = *reinterpret_cast<const float2*>(sp + 2 * lane);
          const float ra = sp2.x * rsa, rb = sp2.y * rsb;
          const float inv = 1.f / warp_sum(ra + rb);
          __syncwarp();
          *reinterpret_cast<float2*>(vZ + 2 * lane) = make_float2(fminf(ra * inv, 1.f), fminf(rb * inv, 1.f));
          __syncwarp();
          {  // 64 -> 32: lane = (four columns, strand of 16 latents)
            const float* Wz = Wp(off.dec_w0);
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int kk = 0; kk < 16; ++kk) {
              const int m = 16 * skq + kk;
              const float4 w = *reinterpret_cast<const float4*>(Wz + m * 32 + 4 * sq);
              const float z = vZ[m];
              h.x = fmaf(z, w.x, h.x); h.y = fmaf(z, w.y, h.y); h.z = fmaf(z, w.z, h.z); h.w = fmaf(z, w.w, h.w);
            }
#pragma unroll
            for (int d = 8; d <= 16; d <<= 1) {
              h.x += __shfl_xor_sync(0xFFFFFFFFu, h.x, d); h.y += __shfl_xor_sync(0xFFFFFFFFu, h.y, d);
              h.z += __shfl_xor_sync(0xFFFFFFFFu, h.z, d); h.w += __shfl_xor_sync(0xFFFFFFFFu, h.w, d);
            }
            if (skq == 0) {
              const float4 d0 = *reinterpret_cast<const float4*>(gs.d0 + ind * 32 + 4 * sq);
              *reinterpret_cast<float4*>(vA + 4 * sq) = make_float4(fmaxf(h.x + d0.x, 0.f), fmaxf(h.y + d0.y, 0.f), fmaxf(h.z + d0.z, 0.f), fmaxf(h.w + d0.w, 0.f));
            }
          }
          __syncwarp();
          {
            const float4 h = splitk_32(vA, Wp(off.dec_w1), sq, skq);
            if (skq == 0) {
              const float4 bb = *reinterpret_cast<const float4*>(Wp(off.dec_b1) + 4 * sq);
              *reinterpret_cast<float4*>(vB + 4 * sq) = make_float4(fmaxf(h.x + bb.x, 0.f), fmaxf(h.y + bb.y, 0.f), fmaxf(h.z + bb.z, 0.f), fmaxf(h.w + bb.w, 0.f));
            }
          }
          __syncwarp();
          const float o = head4(vB, Wp(off.dec_w2), Wp(off.dec_b2), lane);
          if (lane < 3) s_samples[((size_t)io * bt.K + k) * 3 + lane] = o;
        }
        group_bar(g);
        IK_MARK(11)
      }

      // ---- step (rollout.cuh): warp 0 selects loc_pred while warps 0..2 already hold V[t-1] / V[t+1] / V[t]; then the
      //      parents, the children, the merge candidates and the goal test of merge_samples run side by side
      if (!wide_step) {
        StepCtx sc;
        sc.b = b; sc.t = t; sc.kt = kt; sc.makespan = makespan;
        sc.a0 = a0; sc.N = n; sc.o0 = o0; sc.nO = nO;
        sc.inv_nO = (nO >= 1 && nO <= 128) ? c_pair_inv.v[nO] : 0u;
        sc.gx = ap[0]; sc.gy = ap[1]; sc.speed = ap[2]; sc.speed2 = ap[3]; sc.merge2 = ap[4]; sc.rad = ap[5];
        sc.sobs = obs_smem ? g_obs : nullptr;
        const double cx = s_cur[2 * i], cy = s_cur[2 * i + 1];
        const bool arrived = s_arrived[i] != 0;
        const int nl = bv.nlayers[a];
        const int32_t* vcn = bv.vcnt + (size_t)a * bt.L;
        const int n_prev = vcn[t - 1], n_t = (nl > t) ? vcn[t] : 0, n_next = (nl > t + 1) ? vcn[t + 1] : 0;
        const bool has_next = nl > t + 1;
        // this warp's layer: 0 parents V[t-1], 1 children V[t+1], 2 merge candidates V[t]
        const int lay_t = gw == 0 ? t - 1 : (gw == 1 ? t + 1 : t);
        const int lay_n = gw == 0 ? n_prev : (gw == 1 ? n_next : (gw == 2 ? n_t : 0));
        double px[W], py[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          px[w] = 0.0; py[w] = 0.0;
          const int sl = w * 32 + lane;
          if (sl < lay_n) {
            const double2 p = *reinterpret_cast<const double2*>(&bt.vpos[2 * vslot(bt, a, lay_t, sl)]);
            px[w] = p.x; py[w] = p.y;
          }
        }
        unsigned int count = 0;
        if (gw == 0) {
          double sx, sy;
          count = select_agent(bv, sc, a, lane, cx, cy, arrived, &sx, &sy, (s_learned[i] >> 1) & 1);
          if (lane == 0) { gs.lx = sx; gs.ly = sy; }
        }
        group_bar(g);
        IK_MARK(12)
        const double lx = gs.lx, ly = gs.ly;
        uint32_t Mg[W], mpu[W][W], mcu[W][W];
        if (gw < 2) {
          // warp 0, parents: V[t-1] within speed and collision free, tested as (parent -> loc)   utils.py:58-69
          // warp 1, children: V[t+1], tested as (child -> loc)   utils.py:72-84
          const bool on = gw == 0 || has_next;
#pragma unroll
          for (int w = 0; w < W; ++w) {
            uint32_t Ew = 0u;
            if (on && w * 32 < lay_n) {
              const bool cand = (w * 32 + lane < lay_n) && dist2(px[w], py[w], lx, ly) <= sc.speed2;
              const uint32_t cm = __ballot_sync(0xFFFFFFFFu, cand);
              count += __popc(cm);
              Ew = cm & ~hits_any_flat(bt, sc, cm, px[w], py[w], lx, ly, lane);
            }
            if (lane == 0) (gw == 0 ? gs.P : gs.Cm)[w] = Ew;
          }
        } else if (gw == 2) {  // merge candidates: V[t] within merge_distance (utils.py:86-89), their edge masks loaded ahead
#pragma unroll
          for (int w = 0; w < W; ++w) {
            Mg[w] = 0u;
            bool cand = false;
            if (w * 32 < n_t) {
              cand = (w * 32 + lane < n_t) && dist2(px[w], py[w], lx, ly) <= sc.merge2;
              Mg[w] = __ballot_sync(0xFFFFFFFFu, cand);
            }
            const size_t v = vslot(bt, a, t, w * 32 + lane);
#pragma unroll
            for (int q = 0; q < W; ++q) {
              mpu[w][q] = cand ? bt.pmask[v * W + q] : 0u;
              mcu[w][q] = cand ? bt.cmask[v * W + q] : 0u;
            }
          }
        } else {  // the goal test valid_edge(loc, goal) (learned_sampler.py:182-183) for loc_next = loc_pred, the common outcome
          const int res = goal_test(bt, sc, lx, ly, lane);
          if (lane == 0) { gs.spec_tested = res & 1; gs.spec_arr = res >> 1; }
        }
        group_bar(g);
        if (gw == 2) {
          // rules (utils.py:94-133) over the candidates in index order, every candidate's test on its own lane
          uint32_t P[W], Cm[W];
#pragma unroll
          for (int w = 0; w < W; ++w) { P[w] = gs.P[w]; Cm[w] = gs.Cm[w]; }
          const double gx = sc.gx, gy = sc.gy;
          double rx = lx, ry = ly;
          int target = -1;  // vertex of layer t that receives loc (replace) or -1
          bool add_edges = false, resolved = false;
          const double h_loc = dist2(lx, ly, gx, gy);
#pragma unroll
          for (int w = 0; w < W; ++w) {
            if (!resolved && Mg[w]) {
              bool eq = true, sup = true, sub = true;
#pragma unroll
              for (int q = 0; q < W; ++q) {
                const uint32_t pu = mpu[w][q], cu = mcu[w][q];
                eq = eq && pu == P[q] && cu == Cm[q];
                sup = sup && (P[q] & ~pu) == 0u && (Cm[q] & ~cu) == 0u;
                sub = sub && (pu & ~P[q]) == 0u && (cu & ~Cm[q]) == 0u;
              }
              const bool cand = (Mg[w] >> lane) & 1u;
              const uint32_t hitm = __ballot_sync(0xFFFFFFFFu, cand && (eq || sup || sub));
              if (hitm) {
                const int src = __ffs(hitm) - 1, u = w * 32 + src;
                const int code = __shfl_sync(0xFFFFFFFFu, eq ? 1 : (sup ? 2 : 3), src);
                const double ux = __shfl_sync(0xFFFFFFFFu, px[w], src), uy = __shfl_sync(0xFFFFFFFFu, py[w], src);
                if (code == 1) {
                  if (h_loc < dist2(ux, uy, gx, gy)) target = u;  // replace u by loc
                  else { rx = ux; ry = uy; }                       // abandon loc
                } else if (code == 2) {
                  rx = ux; ry = uy;
                } else {
                  target = u;
                  add_edges = true;
                }
                resolved = true;
              }
            }
          }
          if (!resolved) {  // trm.append_sample(loc, t, parents, children)   utils.py:136, timed_roadmap.py:67-104
            target = n_t;
            add_edges = true;
            if (target >= bt.S) target = -1;  // cannot happen: one insert per (k_traj, t, agent) and S >= N_traj + 1
          }
          if (target >= 0) {
            const size_t v = vslot(bt, a, t, target);
            if (lane == 0) {
              bt.vpos[2 * v] = lx;
              bt.vpos[2 * v + 1] = ly;
              if (!resolved) {
                bv.vcnt[(size_t)a * bt.L + t] = n_t + 1;
                if (nl < t + 1) bv.nlayers[a] = t + 1;
              }
            }
            if (add_edges) {
              const uint32_t tbit = 1u << (target & 31);
              const int tw = target >> 5;
#pragma unroll
              for (int w = 0; w < W; ++w) {
                const uint32_t pu = resolved ? bt.pmask[v * W + w] : 0u;
                const uint32_t cu = resolved ? bt.cmask[v * W + w] : 0u;
                const int s = w * 32 + lane;
                // reductions without a return value: the group is the only writer of its agent's roadmap
                if ((P[w] & ~pu) >> lane & 1u) atomicOr(&bt.cmask[vslot(bt, a, t - 1, s) * W + tw], tbit);   // E[t-1][p].append(u)
                if ((Cm[w] & ~cu) >> lane & 1u) atomicOr(&bt.pmask[vslot(bt, a, t + 1, s) * W + tw], tbit);  // mirror of E[t][u] += ...
                __syncwarp();
                if (lane == 0) {
                  bt.pmask[v * W + w] = pu | P[w];
                  bt.cmask[v * W + w] = cu | Cm[w];
                }
              }
            }
          }
          // goal test: valid_edge(loc_next, goal)   learned_sampler.py:182-183
          bool arr = false;
          if (rx == lx && ry == ly) {
            count += (unsigned)gs.spec_tested;
            arr = gs.spec_arr != 0;
          } else {
            const int res = goal_test(bt, sc, rx, ry, lane);
            count += (unsigned)(res & 1);
            arr = (res >> 1) != 0;
          }
          // next position and arrival flag into every CTA of the cluster
          const bool arr_new = arr || arrived;
          if (lane < C) {
            double* dn = cluster.map_shared_rank(s_nxt, (unsigned)lane) + (size_t)pb * 2 * IK_NMAX;
            *reinterpret_cast<double2*>(dn + 2 * i) = make_double2(rx, ry);
            cluster.map_shared_rank(s_arrn, (unsigned)lane)[pb * IK_NMAX + i] = arr_new ? 1 : 0;
          }
          if (lane == 0 && bt.tr_loc_next != nullptr) {
            const size_t si = step_index(bt, kt, t);
            bt.tr_loc_next[2 * (si * bt.A + a)] = rx;
            bt.tr_loc_next[2 * (si * bt.A + a) + 1] = ry;
          }
        }
        cnt_acc += count;
        IK_MARK(13)
      }
      if (io + IK_GROUPS < n_own) group_bar(g);  // the group's scratch and obstacles are free for its next agent
    }
    if (wide_step) {
      // many agents per CTA (one or two CTAs per instance): sixteen agents at a time, select + merge_samples of an agent on one
      // warp (the batch path's merge_agent) — four groups would take them four at a time
      __syncthreads();  // every agent's samples are in place
      Batch bw = bv;    // next positions / arrival flags land in this CTA's copy of the step's exchange buffers
      bw.nxt = s_nxt + (size_t)pb * 2 * IK_NMAX - 2 * (size_t)a0;
      bw.arrived = s_arrn + pb * IK_NMAX - a0;
      const bool obs_w = nO <= (IK_GROUPS * IK_OBS_MAX) / IK_WARPS;
      double* w_obs = s_obs + warp * 3 * ((IK_GROUPS * IK_OBS_MAX) / IK_WARPS);
      for (int i = own_lo + warp; i < own_hi; i += IK_WARPS) {
        const int a = a0 + i;
        const double* ap = s_pack + 6 * (i - own_lo);
        StepCtx sc;
        sc.b = b; sc.t = t; sc.kt = kt; sc.makespan = makespan;
        sc.a0 = a0; sc.N = n; sc.o0 = o0; sc.nO = nO;
        sc.inv_nO = (nO >= 1 && nO <= 128) ? c_pair_inv.v[nO] : 0u;
        sc.gx = ap[0]; sc.gy = ap[1]; sc.speed = ap[2]; sc.speed2 = ap[3]; sc.merge2 = ap[4]; sc.rad = ap[5];
        sc.sobs = nullptr;
        __syncwarp();
        if (obs_w) {
          for (int o = lane; o < nO; o += 32) {
            const double* gp = bt.obs_xyr + 3 * (size_t)(o0 + o);
            w_obs[3 * o] = __ldg(gp);
            w_obs[3 * o + 1] = __ldg(gp + 1);
            w_obs[3 * o + 2] = f64add(sc.rad, __ldg(gp + 2));  // rad1 + rad2, rounded once (collision_check.cpp:56)
          }
          sc.sobs = w_obs;
        }
        const bool arrived = s_arrived[i] != 0;
        if (lane == 0) s_arrn[pb * IK_NMAX + i] = arrived ? 1 : 0;  // merge_agent only ever sets the flag
        __syncwarp();
        MergePrefetch<W> pf;
        merge_prefetch<W>(bw, a, t, lane, pf);
        double lx, ly;
        unsigned int count = select_agent(bw, sc, a, lane, s_cur[2 * i], s_cur[2 * i + 1], arrived, &lx, &ly, (s_learned[i] >> 1) & 1);
        count += merge_agent<W>(bw, sc, a, lane, lx, ly, pf);
        cnt_acc += count;
        __syncwarp();
        if (lane < C && lane != r) {  // the other CTAs' copies
          double* dn = cluster.map_shared_rank(s_nxt, (unsigned)lane) + (size_t)pb * 2 * IK_NMAX;
          *reinterpret_cast<double2*>(dn + 2 * i) = *reinterpret_cast<const double2*>(s_nxt + (size_t)pb * 2 * IK_NMAX + 2 * i);
          cluster.map_shared_rank(s_arrn, (unsigned)lane)[pb * IK_NMAX + i] = s_arrn[pb * IK_NMAX + i];
        }
      }
    }
    IK_CLUSTER_SYNC();
    IK_MARK(14)
    // ================= advance (learned_sampler.py:186-196, :65-73): every CTA keeps its own copy of the instance state ====
    bool arr_ok = true;
    double2 nx = make_double2(0.0, 0.0);
    if (tid < n) {
      arr_ok = s_arrn[pb * IK_NMAX + tid] != 0;
      nx = *reinterpret_cast<const double2*>(s_nxt + (size_t)pb * 2 * IK_NMAX + 2 * tid);
      s_arrived[tid] = arr_ok ? 1 : 0;
    }
    const bool all = __syncthreads_and(arr_ok) != 0;
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
      }
      t = 1;
    }
    IK_MARK(15)
  }
  __syncthreads();
  if (s_vcnt != nullptr) {
    for (int i = tid; i < n_own * bt.L; i += IK_THREADS) bt.vcnt[(size_t)(a0 + own_lo) * bt.L + i] = s_vcnt[i];
    for (int i = tid; i < n_own; i += IK_THREADS) bt.nlayers[a0 + own_lo + i] = s_nl[i];
  }
  if (lane == 0 && cnt_acc) atomicAdd(&bt.cnt_collide[b], (unsigned long long)cnt_acc);
  if (prof && tid == 0)
    for (int i = 0; i < IK_NPROF; ++i) ia.prof[(size_t)blockIdx.x * IK_NPROF + i] = s_prof[i];
  if (r == 0 && tid == 0) {
    bt.k_traj[b] = kt; bt.t[b] = t; bt.makespan[b] = makespan; bt.done[b] = 1;
    bt.steps_done[b] = steps;
    int* ip = bt.inst_pack + 8 * b;
    ip[0] = 1; ip[1] = t; ip[2] = kt; ip[3] = makespan;
    atomicAdd(bt.n_done, 1);
  }
  cluster.sync();  // no CTA leaves while a neighbour may still store into its shared memory
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t instance_w0_image_bytes() { return (size_t)IK_BYTES * 256 * 64 * sizeof(float); }

int launch_instance_w0_image(const float* d_blob, const ctrm_weight_offsets_t& off, int F, uint8_t* d_image, cudaStream_t st) {
  const int total = IK_BYTES * 256 * 64;
  instance_w0_image_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_blob + off.fov_w0, 2 * F * F, reinterpret_cast<float*>(d_image));
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

// shared memory of one CTA at cluster size C; fills the layout-dependent fields of `ia`
static size_t instance_smem_bytes(InstArgs& ia, int n_max, int K, int L, int C) {
  auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const int own = (n_max + C - 1) / C;
  ia.own_max = own;
  ia.vc_smem = (size_t)own * L * 4 <= 12288 ? 1 : 0;
  ia.l0_split = own <= 16 ? 1 : 0;  // a warp per quarter of an agent's inputs while the quarter sums (1 KB per agent) stay small
  size_t s = al((size_t)ia.n_w * 4);
  auto take = [&](size_t bytes) { const size_t at = s; s += al(bytes); return (int)at; };
  IkLayout& y = ia.lay;
  y.cur = take(IK_NMAX * 2 * 8);
  y.prev = take(IK_NMAX * 2 * 8);
  y.goal = take(IK_NMAX * 2 * 8);
  y.nxt = take(2 * IK_NMAX * 2 * 8);
  y.vp = take(IK_NMAX * 32 * 4);
  y.eself = take((size_t)own * 32 * 4);
  y.l0p = take((size_t)own * (ia.l0_split ? 4 : 1) * 64 * 4);
  y.radf = take(IK_NMAX * 4);
  y.speedf = take(IK_NMAX * 4);
  y.flags = take(4 * IK_NMAX);
  y.pack = take((size_t)own * 6 * 8);
  y.samples = take((size_t)own * K * 3 * 4);
  y.nl = take((size_t)own * 4);
  y.vcnt = ia.vc_smem ? take((size_t)own * L * 4) : 0;
  y.obs = take((size_t)IK_GROUPS * 3 * IK_OBS_MAX * 8);
  y.prof = take(IK_NPROF * 8);
  y.wvf = take(33 * 32 * 4);
  {  // the PreAgent records if they still fit
    const size_t u_bytes = std::max((size_t)IK_GROUPS * sizeof(GroupScratch), (size_t)own * 128 * 4);
    ia.pre_rank = s + al((size_t)own * sizeof(PreAgent)) + al(u_bytes) <= (size_t)227 * 1024 ? 1 : 0;
    y.pre = ia.pre_rank ? take((size_t)own * sizeof(PreAgent)) : 0;
  }
  y.U = take(std::max((size_t)IK_GROUPS * sizeof(GroupScratch), (size_t)own * 128 * 4));
  y.total = (int)s;
  return s;
}

// can the per-instance kernel run this batch?  (everything else takes the batch path)
int instance_kernel_supports(const DevModel& m, const Batch& bt) {
  const ctrm_model_desc& d = m.desc;
  if (bt.n_max > IK_NMAX || bt.K > IK_KMAX || bt.randomize_indicator) return 0;
  if (d.temp != 2.0f || d.dim_ind != 3) return 0;
  if (2 * d.fov_size * d.fov_size > 32 * IK_WORDS) return 0;
  const int k_nb = d.use_k_neighbor ? (d.num_neighbors < bt.n_max - 1 ? d.num_neighbors : bt.n_max - 1) : bt.n_max - 1;
  if (k_nb + 1 > IK_PMAX) return 0;
  if (bt.tab_uniforms != nullptr) return 0;
  return 1;
}

// `narrow`: a group of four warps per agent in the step (at most two agents per group), `wide`: one warp per agent
template <typename Kernel>
static int launch_instance_with(Kernel narrow, Kernel wide, SmemAttrCache& attr_n, SmemAttrCache& attr_w, const Batch& bt,
                                InstArgs& ia, int C_req, cudaStream_t st) {
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
    CTRM_CUDA_OK(attr_n.ensure(narrow, most));
    CTRM_CUDA_OK(attr_w.ensure(wide, most));
  }
  int C = 1;
  for (int c = C_req > 0 ? C_req : 8; c >= 1; c >>= 1) {
    cfg.gridDim = dim3((unsigned)(bt.B * c), 1, 1);
    cfg.dynamicSmemBytes = instance_smem_bytes(ia, bt.n_max, bt.K, bt.L, c);
    at[0].val.clusterDim.x = (unsigned)c;
    int n_clusters = 0;
    CTRM_CUDA_OK(cudaOccupancyMaxActiveClusters(&n_clusters, narrow, &cfg));
    C = c;
    if (n_clusters >= bt.B || C_req > 0) break;
  }
  ia.C = C;
  cfg.gridDim = dim3((unsigned)(bt.B * C), 1, 1);
  cfg.dynamicSmemBytes = instance_smem_bytes(ia, bt.n_max, bt.K, bt.L, C);
  at[0].val.clusterDim.x = (unsigned)C;
  const bool use_wide = ia.own_max > 2 * IK_GROUPS;
  CTRM_CUDA_OK(cudaLaunchKernelEx(&cfg, use_wide ? wide : narrow, bt, ia));
  return 0;
}

// C = 0: as many CTAs per instance (8, 4, 2, 1) as keep every cluster of the batch resident at once
int launch_instance_rollout(const DevModel& m, const Batch& bt, int nn, int C, cudaStream_t st) {
  if (bt.B <= 0) return 0;
  if (C != 0 && C != 1 && C != 2 && C != 4 && C != 8) return ctrm_fail(-1, "instance kernel: cluster size must be 0 (auto), 1, 2, 4 or 8");
  if (m.inst_w0_img == nullptr) return ctrm_fail(-1, "instance kernel: table of the FOV first layer missing");
  InstArgs ia;
  ia.blob = m.blob;
  ia.lut = reinterpret_cast<const float*>(m.inst_w0_img);
  ia.off = m.off;
  ia.F = m.desc.fov_size;
  ia.num_neighbors = m.desc.num_neighbors;
  ia.use_k_neighbor = m.desc.use_k_neighbor;
  ia.include_self = m.desc.include_self_attention;
  ia.nn = nn;
  ia.C = 1;
  ia.n_w = m.off.total - m.off.fov_b0;
  ia.l0_split = 0;
  ia.pre_rank = 0;
  ia.vc_smem = 0;
  ia.own_max = bt.n_max;
  ia.lay = IkLayout{};
  ia.prof = nullptr;
  const bool want_prof = getenv("CTRM_INSTANCE_PROF") != nullptr;
  if (want_prof) CTRM_CUDA_OK(cudaMalloc(&ia.prof, (size_t)bt.B * 8 * IK_NPROF * sizeof(long long)));
  static SmemAttrCache attr[8];
  switch (bt.W) {
    case 1: CHECK_RC(launch_instance_with(instance_rollout_kernel<1, false>, instance_rollout_kernel<1, true>, attr[0], attr[4], bt, ia, C, st)); break;
    case 2: CHECK_RC(launch_instance_with(instance_rollout_kernel<2, false>, instance_rollout_kernel<2, true>, attr[1], attr[5], bt, ia, C, st)); break;
    case 3: CHECK_RC(launch_instance_with(instance_rollout_kernel<3, false>, instance_rollout_kernel<3, true>, attr[2], attr[6], bt, ia, C, st)); break;
    case 4: CHECK_RC(launch_instance_with(instance_rollout_kernel<4, false>, instance_rollout_kernel<4, true>, attr[3], attr[7], bt, ia, C, st)); break;
    default: return ctrm_fail(-1, "unsupported adjacency mask width");
  }
  if (want_prof) {  // debugging aid: cycles per phase of the first and the last CTA, thread 0's view (group 0, warp 0)
    const int Cu = ia.C;
    std::vector<long long> h((size_t)bt.B * Cu * IK_NPROF);
    CTRM_CUDA_OK(cudaStreamSynchronize(st));
    CTRM_CUDA_OK(cudaMemcpy(h.data(), ia.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(ia.prof);
    static const char* names[IK_NPROF] = {"gate", "raster+fov_l0", "#learned", "sync", "fov_tail", "bar1", "rank", "pair_mlp", "softmax",
                                          "cvae_l0", "prior", "decode", "select", "merge", "bar2", "advance"};
    for (size_t cta : {(size_t)0, (size_t)bt.B * Cu - 1}) {
      fprintf(stderr, "[instance kernel] C=%d CTA %zu Mcycles:", Cu, cta);
      for (int i = 0; i < IK_NPROF; ++i) fprintf(stderr, " %s %.2f", names[i], 1e-6 * (double)h[cta * IK_NPROF + i]);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}
