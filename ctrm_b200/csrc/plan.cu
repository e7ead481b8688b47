// Prioritized planning over the timed roadmaps: the consumer of the roadmaps, batched.
//   reference: PrioritizedPlanning._solve (planner/prioritized_planning.py:40-101), Planner.get_single_agent_path
//   (planner/planner.py:311-384: space-time A*, CLOSE set on generation, OPEN ordered by (f, t)),
//   Planner.collide_dynamic_agents (planner.py:216-263) -> continuousCollideSpheres (collision_check.cpp:16-58).
//
// One CTA per instance; agents in priority order, as the reference.  The A* of one agent keeps OPEN as an f-value per
// roadmap vertex in shared memory (+inf = not open): "pop" is a block-wide arg-min (ties: smaller (t, index), the
// vertex-id order), "push" is a plain store by whichever thread validated the successor — there is no sequential heap.
// A pop's successors are validated against every earlier agent's move of that time step at once: (successor, earlier
// agent) pairs are flattened over the CTA, verdicts OR-ed into a shared bitmask.  fp64, the reference's operation order,
// no FMA: verdicts, expansion order, paths and counters equal the reference's (exact (f, t) ties excepted: the reference
// breaks them with np.random.rand()).
#include "common.cuh"
#include "kernels.h"

#define PL_THREADS 128
#define PL_MAXDEG 128

struct PlanArgs {
  int B, L;
  const int32_t* agent_off;   // [B+1]
  const int32_t* n_layers;    // [A]
  const int32_t* layer_ptr;   // [A*L+1] global vertex id of (agent, t, 0)
  const double* pos;          // [nv][2]
  const int64_t* eptr;        // [nv+1]
  const int32_t* eidx;        // [ne] local indices into the next layer
  const double* goals;        // [A][2]
  const double* speeds;       // [A]
  const double* rads;         // [A]
  int32_t* paths;             // [A][L] chosen vertex index per time step, -1 beyond the horizon / unsolved
  int32_t* solved;            // [B]
  int32_t* failed_agent;      // [B] local index of the agent that found no path, -1 if solved
  long long* counters;        // [B][3] collision checks, expanded, explored
  double* sol_pos;            // [A][L][2] scratch: positions along the chosen paths
  int max_nv;                 // largest vertex count of one agent (shared-memory sizing)
};

__device__ __forceinline__ double pl_f(int t, double px, double py, double gx, double gy, double speed) {
  const double d0 = f64sub(gx, px), d1 = f64sub(gy, py);
  const double n = __dsqrt_rn(f64add(f64mul(d0, d0), f64mul(d1, d1)));
  return f64add((double)t, __ddiv_rn(n, speed));  // v.t + norm(goal - pos) / max_speed   (prioritized_planning.py:54)
}

__global__ void __launch_bounds__(PL_THREADS) plan_prioritized_kernel(PlanArgs p) {
  extern __shared__ __align__(16) uint8_t pl_smem[];
  double* fval = reinterpret_cast<double*>(pl_smem);                       // [max_nv]
  int16_t* par = reinterpret_cast<int16_t*>(fval + p.max_nv);              // [max_nv] parent index, -1 = not generated
  uint8_t* tl = reinterpret_cast<uint8_t*>(par + p.max_nv);                // [max_nv] layer of the vertex
  __shared__ double red_f[PL_THREADS / 32];
  __shared__ int red_i[PL_THREADS / 32];
  __shared__ uint32_t hitmask[PL_MAXDEG / 32];
  __shared__ int s_best, s_newly, s_tfin;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = p.agent_off[b], N = p.agent_off[b + 1] - a0;
  if (N <= 0) {
    if (tid == 0) { p.solved[b] = 1; p.failed_agent[b] = -1; }
    return;
  }
  const int T = p.n_layers[a0] - 1;
  int required = 1;
  long long cnt = 0, expanded = 0, explored = 0;  // kept by thread 0
  bool ok_all = true;
  for (int ag = 0; ag < N && ok_all; ++ag) {
    const int a = a0 + ag;
    const int32_t* lp = p.layer_ptr + (size_t)a * p.L;
    const int first = lp[0], nv = lp[T + 1] - first;
    const double gx = p.goals[2 * a], gy = p.goals[2 * a + 1], speed = p.speeds[a], rad = p.rads[a];
    for (int v = tid; v < nv; v += PL_THREADS) {
      fval[v] = INFINITY;
      par[v] = -1;
    }
    for (int t = warp; t <= T; t += PL_THREADS / 32)
      for (int v = lp[t] - first + lane; v < lp[t + 1] - first; v += 32) tl[v] = (uint8_t)t;
    __syncthreads();
    if (tid == 0) {
      fval[0] = pl_f(0, p.pos[2 * (size_t)first], p.pos[2 * (size_t)first + 1], gx, gy, speed);
      expanded += 1;
      s_tfin = -1;
    }
    __syncthreads();
    int fin_v = -1;
    while (true) {
      // ---- pop: arg-min of f over the open vertices, ties to the smaller vertex id (= smaller t, then smaller index)
      double bf = INFINITY;
      int bi = 0x7FFFFFFF;
      for (int v = tid; v < nv; v += PL_THREADS) {
        const double f = fval[v];
        if (f < bf) { bf = f; bi = v; }  // v ascending per thread: strict < keeps the smallest id
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double of = __shfl_xor_sync(0xFFFFFFFFu, bf, d);
        const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, d);
        if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
      }
      if (lane == 0) { red_f[warp] = bf; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < PL_THREADS / 32; ++w)
          if (red_f[w] < bf || (red_f[w] == bf && red_i[w] < bi)) { bf = red_f[w]; bi = red_i[w]; }
        s_best = (bf < INFINITY) ? bi : -1;
        if (s_best >= 0) {
          fval[s_best] = INFINITY;
          explored += 1;
        }
        hitmask[0] = hitmask[1] = hitmask[2] = hitmask[3] = 0u;
        s_newly = 0;
      }
      __syncthreads();
      const int v = s_best;
      if (v < 0) break;  // OPEN ran dry: no path (planner.py:384)
      const int t = tl[v], idx = v - (lp[t] - first);
      if (t >= required && idx == lp[t + 1] - lp[t] - 1) {  // check_fin: the last vertex of a layer is the goal
        fin_v = v;
        break;
      }
      if (t == T) continue;  // no successors
      const size_t gv = (size_t)first + v;
      const long long e0 = p.eptr[gv];
      const int deg = (int)(p.eptr[gv + 1] - e0);
      if (deg > PL_MAXDEG) {  // roadmaps built here have at most 128 slots per layer; a foreign CSR may not: report, do not overrun
        if (tid == 0) p.failed_agent[b] = -2;
        fin_v = -1;
        break;
      }
      const int nfirst = lp[t + 1] - first;  // local id of (t + 1, 0)
      const double vx = p.pos[2 * gv], vy = p.pos[2 * gv + 1];
      // ---- valid_successor for all (successor, earlier agent) pairs at once
      if (ag > 0) {
        const int npairs = deg * ag;
        for (int q = tid; q < npairs; q += PL_THREADS) {
          const int s = q / ag, k = q - s * ag;
          const int j = p.eidx[e0 + s];
          if (par[nfirst + j] >= 0) continue;  // CLOSE
          const size_t gu = (size_t)first + nfirst + j;
          const double f1[2] = {vx, vy}, t1[2] = {p.pos[2 * gu], p.pos[2 * gu + 1]};
          const double* sp = p.sol_pos + ((size_t)(a0 + k) * p.L + t) * 2;
          const double f2[2] = {sp[0], sp[1]}, t2[2] = {sp[2], sp[3]};
          if (sweep_pair(f1, t1, rad, f2, t2, p.rads[a0 + k], 2)) atomicOr(&hitmask[s >> 5], 1u << (s & 31));
        }
      }
      __syncthreads();
      // ---- generate the valid successors: CLOSE, PARENT, insert
      int fresh = 0, gen = 0;
      for (int s = tid; s < deg; s += PL_THREADS) {
        const int j = p.eidx[e0 + s];
        const int u = nfirst + j;
        if (par[u] >= 0) continue;
        fresh += 1;
        if ((hitmask[s >> 5] >> (s & 31)) & 1u) continue;
        const size_t gu = (size_t)first + u;
        par[u] = (int16_t)idx;
        fval[u] = pl_f(t + 1, p.pos[2 * gu], p.pos[2 * gu + 1], gx, gy, speed);
        gen += 1;
      }
      if (fresh | gen) atomicAdd(&s_newly, fresh | (gen << 16));
      __syncthreads();
      if (tid == 0) {
        cnt += (long long)(s_newly & 0xFFFF) * ag;  // every (successor, earlier agent) pair is one collide_dynamic_agents call
        expanded += s_newly >> 16;
      }
    }
    // ---- backtrack / failure
    if (fin_v < 0) {
      ok_all = false;
      if (tid == 0) {
        p.solved[b] = 0;
        if (p.failed_agent[b] != -2) p.failed_agent[b] = ag;
      }
      break;
    }
    if (tid == 0) {
      int v = fin_v, t = tl[v];
      s_tfin = t;
      int32_t* path = p.paths + (size_t)a * p.L;
      while (true) {
        const int idx = v - (lp[t] - first);
        path[t] = idx;
        if (t == 0) break;
        v = (lp[t - 1] - first) + par[v];
        t -= 1;
      }
    }
    __syncthreads();
    const int tfin = s_tfin;
    required = max(tfin, required);
    {  // wait at the goal until T (prioritized_planning.py:96), then publish the positions along the path
      int32_t* path = p.paths + (size_t)a * p.L;
      for (int t = tfin + 1 + tid; t <= T; t += PL_THREADS) path[t] = lp[t + 1] - lp[t] - 1;
      __syncthreads();
      for (int t = tid; t <= T; t += PL_THREADS) {
        const size_t g = (size_t)lp[t] + path[t];
        p.sol_pos[((size_t)a * p.L + t) * 2] = p.pos[2 * g];
        p.sol_pos[((size_t)a * p.L + t) * 2 + 1] = p.pos[2 * g + 1];
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (tid == 0) {
    if (ok_all) {
      p.solved[b] = 1;
      p.failed_agent[b] = -1;
    }
    p.counters[3 * b] = cnt;
    p.counters[3 * b + 1] = expanded;
    p.counters[3 * b + 2] = explored;
  }
  if (!ok_all) {  // self.solution.clear()
    for (int i = tid; i < N * p.L; i += PL_THREADS) p.paths[(size_t)a0 * p.L + i] = -1;
  }
}

int launch_plan_prioritized(int B, int L, const int32_t* d_agent_off, const int32_t* d_n_layers, const int32_t* d_layer_ptr,
                            const double* d_pos, const int64_t* d_eptr, const int32_t* d_eidx, const double* d_goals,
                            const double* d_speeds, const double* d_rads, int max_nv, int A, int32_t* d_paths, int32_t* d_solved,
                            int32_t* d_failed_agent, long long* d_counters, double* d_sol_pos, cudaStream_t st) {
  if (B <= 0) return 0;
  if (L > 256) return ctrm_fail(-1, "plan_prioritized: more than 256 layers are not supported");
  if (max_nv > 32767 * 4) return ctrm_fail(-1, "plan_prioritized: roadmap of one agent too large");
  const size_t smem = (size_t)max_nv * (sizeof(double) + sizeof(int16_t) + 1) + 16;
  if (smem > 200 * 1024) return ctrm_fail(-1, "plan_prioritized: roadmap of one agent does not fit shared memory");
  static SmemAttrCache attr;
  if (smem > 48 * 1024) CTRM_CUDA_OK(attr.ensure(plan_prioritized_kernel, smem));
  CTRM_CUDA_OK(cudaMemsetAsync(d_paths, 0xFF, (size_t)A * L * sizeof(int32_t), st));
  PlanArgs p;
  p.B = B; p.L = L; p.agent_off = d_agent_off; p.n_layers = d_n_layers; p.layer_ptr = d_layer_ptr; p.pos = d_pos;
  p.eptr = d_eptr; p.eidx = d_eidx; p.goals = d_goals; p.speeds = d_speeds; p.rads = d_rads; p.paths = d_paths;
  p.solved = d_solved; p.failed_agent = d_failed_agent; p.counters = d_counters; p.sol_pos = d_sol_pos; p.max_nv = max_nv;
  plan_prioritized_kernel<<<B, PL_THREADS, smem, st>>>(p);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
