// The per-step roll-out kernels behind get_timed_roadmaps_multiple_paths_with_learned_indicator
// (roadmap_learned/learned_sampler.py:65-196):
//   select  : candidate validation, gate logic and random walk (learned_sampler.py:114-168; valid_edge closure
//             :127-132; valid_move roadmap/utils.py:18-42; collisions static_objects.py:108-138)
//   merge   : merge_samples (roadmap_learned/utils.py:21-137) on a slot/bitmask timed roadmap, then the goal test
//             (learned_sampler.py:182-183)
//   advance : roll-out bookkeeping (learned_sampler.py:186-196, :65-73)
// One warp per agent (select, merge) / per instance (advance).  Every instance carries its own (k_traj, t).
// All geometry is fp64 without FMA in the reference's operation order: verdicts and topology are bit-exact.
#include "common.cuh"
#include "geometry.cuh"
#include <cstdlib>

#include "kernels.h"

#include "rollout.cuh"


// ------------------------------------------------------------------------------------------------
__global__ void rollout_init_kernel(Batch bt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < bt.A) {
    double sx = bt.starts[2 * i], sy = bt.starts[2 * i + 1];
    bt.cur[2 * i] = sx; bt.cur[2 * i + 1] = sy;
    bt.prev[2 * i] = sx; bt.prev[2 * i + 1] = sy;
    bt.nxt[2 * i] = sx; bt.nxt[2 * i + 1] = sy;
    bt.arrived[i] = 0;
    if (bt.live_agent != nullptr) {
      bt.live_agent[i] = i;
      bt.live_inst[i] = bt.agent_inst[i];
    }
    if (bt.learn_agent != nullptr && bt.N_traj > 0) {  // first step (k_traj 0, t 1); n_learn was zeroed by the host
      const int b = bt.agent_inst[i];
      if (gate_is_learned(bt, b, i, i - bt.agent_off[b], 0, 1, 0, false)) {
        const int slot = atomicAdd(bt.n_learn, 1);
        bt.learn_agent[slot] = i;
        bt.learn_inst[slot] = b;
      }
    }
    double* ap = bt.agent_pack + 6 * (size_t)i;  // {goal x, goal y, speed, speed^2, merge_distance^2, rad}
    ap[0] = bt.goals[2 * i]; ap[1] = bt.goals[2 * i + 1]; ap[2] = bt.speeds[i]; ap[3] = bt.speeds2[i];
    ap[4] = bt.merge2[i]; ap[5] = bt.rads[i];
    bt.nlayers[i] = 1;  // TimedRoadmap(start): V = [[start]], E = [[[]]]  (timed_roadmap.py:59-65)
    bt.vcnt[(size_t)i * bt.L] = 1;
    size_t s = vslot(bt, i, 0, 0);
    bt.vpos[2 * s] = sx;
    bt.vpos[2 * s + 1] = sy;
  }
  if (i < bt.B) {
    bt.k_traj[i] = 0;
    bt.t[i] = 1;
    bt.makespan[i] = 0;
    bt.done[i] = (bt.N_traj <= 0) ? 1 : 0;
    bt.cnt_collide[i] = 0ULL;
    bt.steps_done[i] = 0;
    bt.arrive_cnt[i] = 0;
    int* ip = bt.inst_pack + 8 * i;  // {done, t, k_traj, makespan, agent_off, n_agents, obs_off, n_obs}
    ip[0] = (bt.N_traj <= 0) ? 1 : 0; ip[1] = 1; ip[2] = 0; ip[3] = 0;
    ip[4] = bt.agent_off[i]; ip[5] = bt.agent_off[i + 1] - bt.agent_off[i];
    ip[6] = bt.obs_off[i]; ip[7] = bt.obs_off[i + 1] - bt.obs_off[i];
  }
  if (i == 0) {
    *bt.n_done = (bt.N_traj <= 0) ? bt.B : 0;
    if (bt.n_live != nullptr) *bt.n_live = bt.A;
  }
}

// ------------------------------------------------------------------------------------------------
// live rows: the agents of the unfinished instances, in instance order.  Every roll-out ends at its own time, so from
// about two thirds of the way a lock-step batch is a thinning, scattered set of instances; the per-step kernels index
// their rows through this list, which keeps their tiles full.  Rebuilt between graph chunks (two tiny launches).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) live_scan_kernel(Batch bt) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < bt.B; base += 1024) {
    const int b = base + threadIdx.x;
    const int n = (b < bt.B && !bt.done[b]) ? bt.agent_off[b + 1] - bt.agent_off[b] : 0;
    int x = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, w, d);
        if (lane >= d) w += y;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int excl = x - n + (warp > 0 ? wsum[warp - 1] : 0) + carry;
    if (b < bt.B) bt.live_off[b] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + n;
    __syncthreads();
  }
  if (threadIdx.x == 0) *bt.n_live = carry;
}
__global__ void live_scatter_kernel(Batch bt) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= bt.A) return;
  const int b = bt.agent_inst[a];
  if (!bt.done[b]) {
    const int row = bt.live_off[b] + (a - bt.agent_off[b]);
    bt.live_agent[row] = a;
    bt.live_inst[row] = b;
  }
}
int launch_compact_live(const Batch& bt, cudaStream_t st) {
  if (bt.A <= 0 || bt.live_agent == nullptr) return 0;
  live_scan_kernel<<<1, 1024, 0, st>>>(bt);
  CTRM_CUDA_OK(cudaGetLastError());
  live_scatter_kernel<<<(bt.A + 255) / 256, 256, 0, st>>>(bt);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------
// advance: roll-out bookkeeping of one instance (advance_kernel, one warp per instance, after the step kernel)
// ------------------------------------------------------------------------------------------------
// append the agents of instance b whose next step (kt, t) takes the learned branch to the learned-row list
__device__ __forceinline__ void list_learned(const Batch& bt, int b, int a0, int N, int kt, int t, int makespan, bool fresh,
                                             int lane) {
  for (int i0 = 0; i0 < N; i0 += 32) {
    const int i = i0 + lane;
    bool f = false;
    if (i < N) f = gate_is_learned(bt, b, a0 + i, i, kt, t, makespan, fresh ? false : (__ldcg(&bt.arrived[a0 + i]) != 0));
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, f);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(bt.n_learn, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (f) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      bt.learn_agent[slot] = a0 + i;
      bt.learn_inst[slot] = b;
    }
  }
}

__device__ __forceinline__ void advance_instance(const Batch& bt, const int b, const int lane) {
  const int a0 = bt.agent_off[b], N = bt.agent_off[b + 1] - a0;
  bool all = true;
  for (int i = lane; i < N; i += 32) all = all && (__ldcg(&bt.arrived[a0 + i]) != 0);
  all = __all_sync(0xFFFFFFFFu, all);
  const int t = bt.t[b];
  if (lane == 0) bt.steps_done[b] += 1;
  bool end_roll = false;
  if (all) {  // learned_sampler.py:186-191
    if (lane == 0) {
      bt.makespan[b] = bt.makespan[b] ? max(bt.makespan[b], t) : t;
      bt.inst_pack[8 * b + 3] = bt.makespan[b];
    }
    end_roll = true;
  } else if (t >= bt.max_T) {
    end_roll = true;
  }
  if (!end_roll) {  // learned_sampler.py:194-196
    for (int i = lane; i < N; i += 32) {
      const int a = a0 + i;
      bt.prev[2 * a] = bt.cur[2 * a];
      bt.prev[2 * a + 1] = bt.cur[2 * a + 1];
      bt.cur[2 * a] = __ldcg(&bt.nxt[2 * a]);
      bt.cur[2 * a + 1] = __ldcg(&bt.nxt[2 * a + 1]);
    }
    if (lane == 0) {
      bt.t[b] = t + 1;
      bt.inst_pack[8 * b + 1] = t + 1;
    }
    if (bt.learn_agent != nullptr) list_learned(bt, b, a0, N, bt.k_traj[b], t + 1, bt.makespan[b], false, lane);
  } else {
    const int kt = bt.k_traj[b] + 1;
    if (kt >= bt.N_traj) {
      if (lane == 0) {
        bt.k_traj[b] = kt;
        bt.done[b] = 1;
        bt.inst_pack[8 * b + 0] = 1;
        bt.inst_pack[8 * b + 2] = kt;
        atomicAdd(bt.n_done, 1);
      }
    } else {  // next roll-out restarts from the starts (learned_sampler.py:68-73)
      for (int i = lane; i < N; i += 32) {
        const int a = a0 + i;
        const double sx = bt.starts[2 * a], sy = bt.starts[2 * a + 1];
        bt.cur[2 * a] = sx; bt.cur[2 * a + 1] = sy;
        bt.prev[2 * a] = sx; bt.prev[2 * a + 1] = sy;
        bt.arrived[a] = 0;
      }
      if (lane == 0) {
        bt.k_traj[b] = kt;
        bt.t[b] = 1;
        bt.inst_pack[8 * b + 1] = 1;
        bt.inst_pack[8 * b + 2] = kt;
      }
      // makespan: lane 0 may have just updated it (all arrived) — read it back after the warp has converged
      __syncwarp();
      if (bt.learn_agent != nullptr) list_learned(bt, b, a0, N, kt, 1, __ldcg(&bt.makespan[b]), true, lane);
    }
  }
}

int launch_rollout_init(const Batch& bt, cudaStream_t st) {
  int n = bt.A > bt.B ? bt.A : bt.B;
  if (n <= 0) return 0;
  rollout_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(bt);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
// ------------------------------------------------------------------------------------------------
// step: select -> merge, one warp per agent; then advance, one warp per instance
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(RO_WARPS * 32, 8) step_kernel(Batch bt) {
  __shared__ double s_obs[RO_WARPS][3 * RO_OBS_MAX];
  const int warp = threadIdx.x >> 5;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  // the learned-row list of this step has been consumed (the network kernels ran before this one); the advance kernel that
  // follows refills it for the next step
  if (blockIdx.x == 0 && threadIdx.x == 0 && bt.n_learn != nullptr) *bt.n_learn = 0;
  if (blockIdx.x == 0 && threadIdx.x == 1 && bt.n_retry != nullptr) *bt.n_retry = 0;
  if (row >= (bt.live_agent != nullptr ? *bt.n_live : bt.A)) return;
  const int a = bt.live_agent != nullptr ? bt.live_agent[row] : row;
  StepCtx sc;
  sc.b = bt.live_agent != nullptr ? bt.live_inst[row] : bt.agent_inst[a];
  const int4 p0 = *reinterpret_cast<const int4*>(bt.inst_pack + 8 * sc.b);
  if (p0.x) return;  // instance finished
  const int4 p1 = *reinterpret_cast<const int4*>(bt.inst_pack + 8 * sc.b + 4);
  sc.t = p0.y; sc.kt = p0.z; sc.makespan = p0.w;
  sc.a0 = p1.x; sc.N = p1.y; sc.o0 = p1.z; sc.nO = p1.w;
  sc.inv_nO = (sc.nO >= 1 && sc.nO <= 128) ? c_pair_inv.v[sc.nO] : 0u;
  const double2 q0 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a);
  const double2 q1 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 2);
  const double2 q2 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 4);
  const double2 cur = *reinterpret_cast<const double2*>(bt.cur + 2 * (size_t)a);
  const bool arrived = bt.arrived[a] != 0;
  sc.gx = q0.x; sc.gy = q0.y; sc.speed = q1.x; sc.speed2 = q1.y; sc.merge2 = q2.x; sc.rad = q2.y;
  MergePrefetch<W> pf;
  merge_prefetch<W>(bt, a, sc.t, lane, pf);
  sc.sobs = nullptr;
  if (sc.nO <= RO_OBS_MAX) {
    for (int o = lane; o < sc.nO; o += 32) {
      const double* g = bt.obs_xyr + 3 * (size_t)(sc.o0 + o);
      s_obs[warp][3 * o] = __ldg(g);
      s_obs[warp][3 * o + 1] = __ldg(g + 1);
      s_obs[warp][3 * o + 2] = f64add(sc.rad, __ldg(g + 2));  // rad1 + rad2, rounded once (collision_check.cpp:56)
    }
    __syncwarp();
    sc.sobs = s_obs[warp];
  }
  double lx, ly;
  unsigned int count = select_agent(bt, sc, a, lane, cur.x, cur.y, arrived, &lx, &ly);
  count += merge_agent<W>(bt, sc, a, lane, lx, ly, pf);
  if (lane == 0 && count) atomicAdd(&bt.cnt_collide[sc.b], (unsigned long long)count);
}

// ------------------------------------------------------------------------------------------------
// step, ONE THREAD PER AGENT (large batches).  The warp-per-agent kernel above spends most of its ~1,150 warp instructions
// per agent-step on work that is uniform across the warp (context, addressing, rule evaluation, loop control) with 6-26 useful
// lanes in the parallel parts; with tens of thousands of agents per step there is enough parallelism ACROSS agents, so here
// a thread walks its agent's candidates, layers and obstacles serially, in the reference's own loop order
// (learned_sampler.py:123-183, roadmap_learned/utils.py:21-137).  Same device functions for every fp64 operation and the same
// counting rules, hence the same bits; the warp kernel stays the choice for small batches, where the serial walk of one
// thread would be the latency of the whole step.
// ------------------------------------------------------------------------------------------------
// StaticObjects.collide_continuous_sphere for one segment, one thread walking the instance's obstacles.  Far obstacles are
// rejected in fp32 from a float copy of the obstacle (Batch::obs_f4, one 16-byte load): the centre lies outside the segment's
// bounding box grown by rad_sum + 1e-5.  That is conservative — coordinates live in the unit square, where the float copies and
// the float box are each within 1.2e-7 of the doubles, so a rejected obstacle is at least 9e-6 further away than rad_sum and the
// exact test (closest approach on the segment) could only say "no hit" — and it keeps the common case off the fp64 pipe.
__device__ __forceinline__ bool hits_any_serial(const Batch& bt, int o0, int nO, double rad, double fx, double fy, double tx,
                                                double ty) {
  bool hit = false;
  const float lox = (float)fmin(fx, tx), hix = (float)fmax(fx, tx), loy = (float)fmin(fy, ty), hiy = (float)fmax(fy, ty);
  const float radf = (float)rad + 1e-5f;
  for (int o = 0; o < nO; o += 4) {  // four obstacle loads in flight: the walk is bound by their latency, not by the tests
    float4 q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = __ldg(&bt.obs_f4[o0 + min(o + j, nO - 1)]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (o + j >= nO) break;
      const float R = radf + q[j].z;
      if (q[j].x < lox - R || q[j].x > hix + R || q[j].y < loy - R || q[j].y > hiy + R) continue;
      const double* g = bt.obs_xyr + 3 * (size_t)(o0 + o + j);
      const double ox = __ldg(g), oy = __ldg(g + 1), rs = f64add(rad, __ldg(g + 2));
      hit = hit || sweep_static(fx, fy, tx, ty, ox, oy, rs);
    }
  }
  return hit;
}

template <int W>
__global__ void __launch_bounds__(128, 4) step_thread_kernel(Batch bt) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row == 0 && bt.n_learn != nullptr) *bt.n_learn = 0;  // see step_kernel
  if (row == 1 && bt.n_retry != nullptr) *bt.n_retry = 0;
  if (row >= (bt.live_agent != nullptr ? *bt.n_live : bt.A)) return;
  const int a = bt.live_agent != nullptr ? bt.live_agent[row] : row;
  const int b = bt.live_agent != nullptr ? bt.live_inst[row] : bt.agent_inst[a];
  const int4 p0 = *reinterpret_cast<const int4*>(bt.inst_pack + 8 * b);
  if (p0.x) return;  // instance finished
  const int4 p1 = *reinterpret_cast<const int4*>(bt.inst_pack + 8 * b + 4);
  const int t = p0.y, kt = p0.z, makespan = p0.w, a0 = p1.x, o0 = p1.z, nO = p1.w;
  const int il = a - a0;
  const double2 q0 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a);
  const double2 q1 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 2);
  const double2 q2 = *reinterpret_cast<const double2*>(bt.agent_pack + 6 * (size_t)a + 4);
  const double2 cur = *reinterpret_cast<const double2*>(bt.cur + 2 * (size_t)a);
  const bool arrived = bt.arrived[a] != 0;
  const double gx = q0.x, gy = q0.y, speed = q1.x, speed2 = q1.y, merge2 = q2.x, rad = q2.y;
  const double cx = cur.x, cy = cur.y;
  const size_t si = step_index(bt, kt, t);
  unsigned int count = 0;
  // layer sizes first: their latency overlaps the candidate scan
  const int nl = bt.nlayers[a];
  const int32_t* vc = bt.vcnt + (size_t)a * bt.L;
  const int n_prev = vc[t - 1];
  const int n_t = (nl > t) ? vc[t] : 0;
  const int n_next = (nl > t + 1) ? vc[t + 1] : 0;
  const bool has_next = nl > t + 1;

  // ---- select (learned_sampler.py:148-168): candidates in order, the first valid one wins
  double lx = cx, ly = cy;  // fallback of sample_uniform_i: stay (learned_sampler.py:145)
  {
    const bool learned = gate_is_learned(bt, b, a, il, kt, t, makespan, arrived);
    const int nL = learned ? bt.K : 0, C = nL + bt.max_attempt;
    const uint32_t c0 = bt.inst_base + (uint32_t)b, c1 = rng_ctr1(kt, t);
    for (int c = 0; c < C; ++c) {
      double x, y;
      if (c < nL) {
        const float* s3 = (bt.tab_teacher != nullptr) ? bt.tab_teacher + ((si * bt.A + a) * (size_t)bt.K + c) * 3
                                                      : bt.samples + ((size_t)a * bt.K + c) * 3;
        x = f64add(cx, (double)__fmul_rn(s3[0], s3[2]));  // Y = cur + SAMPLES[:, :2] * SAMPLES[:, 2]   :114-121
        y = f64add(cy, (double)__fmul_rn(s3[1], s3[2]));
      } else {
        const int att = c - nL;
        double u_mag, sn, cs;
        if (bt.rng_mode == 1) {
          const double* r = bt.tab_rw + ((si * bt.A + a) * (size_t)bt.max_attempt + att) * 3;
          u_mag = r[0]; sn = r[1]; cs = r[2];
        } else {
          const uint4 w = philox4x32(c0, c1, (uint32_t)il, rng_ctr3(STREAM_RW, (uint32_t)att), bt.seed_lo, bt.seed_hi);
          u_mag = u64_to_f64(w.x, w.y);
          det_sincos_2pi(u64_to_f64(w.z, w.w), &sn, &cs);
        }
        const double mag = f64mul(speed, u_mag);  // learned_sampler.py:137-142
        x = f64add(f64mul(sn, mag), cx);
        y = f64add(f64mul(cs, mag), cy);
      }
      if (!(bounds_ok(x, y, rad) && speed_ok(cx, cy, x, y, speed, speed2))) continue;
      ++count;
      if (!hits_any_serial(bt, o0, nO, rad, cx, cy, x, y)) {
        lx = x;
        ly = y;
        break;
      }
    }
    if (bt.tr_loc_pred != nullptr) {
      bt.tr_loc_pred[2 * (si * bt.A + a)] = lx;
      bt.tr_loc_pred[2 * (si * bt.A + a) + 1] = ly;
    }
    if (bt.tr_samples != nullptr && bt.samples != nullptr)
      for (int i = 0; i < bt.K * 3; ++i) bt.tr_samples[(si * bt.A + a) * (size_t)bt.K * 3 + i] = bt.samples[(size_t)a * bt.K * 3 + i];
  }

  // ---- merge_samples (roadmap_learned/utils.py:21-137)
  uint32_t P[W], Cm[W], Mg[W];
#pragma unroll
  for (int w = 0; w < W; ++w) { P[w] = 0u; Cm[w] = 0u; Mg[w] = 0u; }
  // pass 1: distances to V[t-1], V[t+1], V[t] (get_dist_arr, utils.py:50-56) — nothing but independent loads and compares, so
  // the loads of a layer are in flight together; a serial walk that tested each vertex as it arrived paid one memory round
  // trip per vertex and made a thread's latency, not the SM's throughput, the time of the kernel
  uint32_t Pc[W], Cc[W];
#pragma unroll
  for (int w = 0; w < W; ++w) { Pc[w] = 0u; Cc[w] = 0u; }
  const double2* vp_prev = reinterpret_cast<const double2*>(&bt.vpos[2 * vslot(bt, a, t - 1, 0)]);
  const double2* vp_t = reinterpret_cast<const double2*>(&bt.vpos[2 * vslot(bt, a, t, 0)]);
  const double2* vp_next = reinterpret_cast<const double2*>(&bt.vpos[2 * vslot(bt, a, t + 1, 0)]);
  {
    // the three layers side by side, three slots of each per round: nine independent 16-byte loads in flight, so the pass costs
    // ceil(max layer size / 3) memory round trips instead of one per four vertices of every layer in turn
    const int n_max3 = max(n_prev, max(n_next, n_t));
    for (int s0 = 0; s0 < n_max3; s0 += 3) {
      double2 pp[3], pn[3], pt[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int sl = s0 + j;
        pp[j] = sl < n_prev ? vp_prev[sl] : make_double2(0.0, 0.0);
        pn[j] = sl < n_next ? vp_next[sl] : make_double2(0.0, 0.0);
        pt[j] = sl < n_t ? vp_t[sl] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int sl = s0 + j;
        const uint32_t bit = 1u << (sl & 31);
        if (sl < n_prev && dist2(pp[j].x, pp[j].y, lx, ly) <= speed2) Pc[sl >> 5] |= bit;
        if (sl < n_next && dist2(pn[j].x, pn[j].y, lx, ly) <= speed2) Cc[sl >> 5] |= bit;
        if (sl < n_t && dist2(pt[j].x, pt[j].y, lx, ly) <= merge2) Mg[sl >> 5] |= bit;  // merge candidates   utils.py:86-89
      }
    }
  }
  // pass 2: swept tests of the vertices within reach; parents as (parent -> loc) utils.py:58-69, children as (child -> loc) :72-84
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t m = Pc[w];
    count += __popc(m);
    while (m) {
      const int sl = w * 32 + __ffs(m) - 1;
      m &= m - 1;
      const double2 p = vp_prev[sl];
      if (!hits_any_serial(bt, o0, nO, rad, p.x, p.y, lx, ly)) P[w] |= 1u << (sl & 31);
    }
    m = Cc[w];
    count += __popc(m);
    while (m) {
      const int sl = w * 32 + __ffs(m) - 1;
      m &= m - 1;
      const double2 p = vp_next[sl];
      if (!hits_any_serial(bt, o0, nO, rad, p.x, p.y, lx, ly)) Cm[w] |= 1u << (sl & 31);
    }
  }
  double rx = lx, ry = ly;
  int target = -1;
  bool add_edges = false, resolved = false;
  const double h_loc = dist2(lx, ly, gx, gy);
#pragma unroll
  for (int w = 0; w < W; ++w) {  // rules (utils.py:94-133), candidates in index order
    uint32_t m = Mg[w];
    while (m && !resolved) {
      const int u = w * 32 + __ffs(m) - 1;
      m &= m - 1;
      const size_t v = vslot(bt, a, t, u);
      bool eq = true, sup = true, sub = true;
#pragma unroll
      for (int q = 0; q < W; ++q) {
        const uint32_t pu = bt.pmask[v * W + q], cu = bt.cmask[v * W + q];
        eq = eq && pu == P[q] && cu == Cm[q];
        sup = sup && (P[q] & ~pu) == 0u && (Cm[q] & ~cu) == 0u;
        sub = sub && (pu & ~P[q]) == 0u && (cu & ~Cm[q]) == 0u;
      }
      const double ux = bt.vpos[2 * v], uy = bt.vpos[2 * v + 1];
      if (eq) {
        if (h_loc < dist2(ux, uy, gx, gy)) target = u;
        else { rx = ux; ry = uy; }
        resolved = true;
      } else if (sup) {
        rx = ux; ry = uy;
        resolved = true;
      } else if (sub) {
        target = u;
        add_edges = true;
        resolved = true;
      }
    }
  }
  if (!resolved) {  // trm.append_sample(loc, t, parents, children)   utils.py:136
    target = n_t;
    add_edges = true;
    if (target >= bt.S) target = -1;
  }
  if (target >= 0) {
    const size_t v = vslot(bt, a, t, target);
    bt.vpos[2 * v] = lx;
    bt.vpos[2 * v + 1] = ly;
    if (!resolved) {
      bt.vcnt[(size_t)a * bt.L + t] = n_t + 1;
      if (nl < t + 1) bt.nlayers[a] = t + 1;
    }
    if (add_edges) {
      const uint32_t tbit = 1u << (target & 31);
      const int tw = target >> 5;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const uint32_t pu = resolved ? bt.pmask[v * W + w] : 0u;
        const uint32_t cu = resolved ? bt.cmask[v * W + w] : 0u;
        uint32_t m = P[w] & ~pu;  // E[t-1][p].append(u): this thread is the only writer of its agent's roadmap
        while (m) {
          const int sl = w * 32 + __ffs(m) - 1;
          m &= m - 1;
          bt.cmask[vslot(bt, a, t - 1, sl) * W + tw] |= tbit;
        }
        m = Cm[w] & ~cu;          // mirror of E[t][u] += children
        while (m) {
          const int sl = w * 32 + __ffs(m) - 1;
          m &= m - 1;
          bt.pmask[vslot(bt, a, t + 1, sl) * W + tw] |= tbit;
        }
        bt.pmask[v * W + w] = pu | P[w];
        bt.cmask[v * W + w] = cu | Cm[w];
      }
    }
  }
  // goal test: valid_edge(loc_next, goal)   learned_sampler.py:182-183
  if (bounds_ok(gx, gy, rad) && speed_ok(rx, ry, gx, gy, speed, speed2)) {
    ++count;
    if (!hits_any_serial(bt, o0, nO, rad, rx, ry, gx, gy)) bt.arrived[a] = 1;
  }
  bt.nxt[2 * a] = rx;
  bt.nxt[2 * a + 1] = ry;
  if (bt.tr_loc_next != nullptr) {
    bt.tr_loc_next[2 * (si * bt.A + a)] = rx;
    bt.tr_loc_next[2 * (si * bt.A + a) + 1] = ry;
  }
  if (count) atomicAdd(&bt.cnt_collide[b], (unsigned long long)count);
  (void)has_next;
}

// roll-out bookkeeping, one warp per instance, after every agent of the step has been processed (a kernel boundary is
// cheaper than electing the instance's last warp with __threadfence + atomics, which was 12% of the step kernel's stalls)
__global__ void __launch_bounds__(RO_WARPS * 32) advance_kernel(Batch bt) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= bt.B || bt.done[b]) return;
  advance_instance(bt, b, lane);
}

// batches from this many agents on take the thread-per-agent kernel unless ctrm_set_step_kernel pins one of the two
// (CTRM_STEP_THREAD_MIN_AGENTS moves the threshold)
static int step_thread_min_agents() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTRM_STEP_THREAD_MIN_AGENTS");
    v = e ? atoi(e) : 20000;  // measured crossover: profiles/r2_step_kernel_ab.md (warp 52 us vs thread 44 us at 25 k agents, 34 vs 41 at 13 k)
    if (v < 0) v = 0;
  }
  return v;
}

int launch_step(const Batch& bt, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  if (bt.step_mode == 2 || (bt.step_mode == 0 && bt.A >= step_thread_min_agents())) {
    const int tb = (bt.A + 127) / 128;
    switch (bt.W) {
      case 1: step_thread_kernel<1><<<tb, 128, 0, st>>>(bt); break;
      case 2: step_thread_kernel<2><<<tb, 128, 0, st>>>(bt); break;
      case 3: step_thread_kernel<3><<<tb, 128, 0, st>>>(bt); break;
      case 4: step_thread_kernel<4><<<tb, 128, 0, st>>>(bt); break;
      default: return ctrm_fail(-1, "unsupported adjacency mask width");
    }
    CTRM_CUDA_OK(cudaGetLastError());
    advance_kernel<<<(bt.B + RO_WARPS - 1) / RO_WARPS, RO_WARPS * 32, 0, st>>>(bt);
    CTRM_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int blocks = (bt.A + RO_WARPS - 1) / RO_WARPS;
  switch (bt.W) {
    case 1: step_kernel<1><<<blocks, RO_WARPS * 32, 0, st>>>(bt); break;
    case 2: step_kernel<2><<<blocks, RO_WARPS * 32, 0, st>>>(bt); break;
    case 3: step_kernel<3><<<blocks, RO_WARPS * 32, 0, st>>>(bt); break;
    case 4: step_kernel<4><<<blocks, RO_WARPS * 32, 0, st>>>(bt); break;
    default: return ctrm_fail(-1, "unsupported adjacency mask width");
  }
  CTRM_CUDA_OK(cudaGetLastError());
  advance_kernel<<<(bt.B + RO_WARPS - 1) / RO_WARPS, RO_WARPS * 32, 0, st>>>(bt);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
