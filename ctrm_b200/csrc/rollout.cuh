// Device functions of one agent-step shared by the batch kernels (rollout.cu) and the per-instance kernel (instance.cu):
// gate, candidate selection and merge_samples for ONE WARP PER AGENT.  See rollout.cu for the reference citations.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

#define RO_WARPS 4

__device__ __forceinline__ size_t step_index(const Batch& bt, int kt, int t) { return (size_t)kt * bt.max_T + (t - 1); }

// the gate of learned_sampler.py:148-157 for agent a (local index il) of instance b at (kt, t): learned branch iff
// (not arrived and u < p_rw) or (arrived and u > prob_after_goal) or k_traj < dim_ind.  The draw is keyed, so the decision can
// be taken ahead of the step (advance kernel: which rows need the networks) and again inside it (select) with the same result.
__device__ __forceinline__ bool gate_is_learned(const Batch& bt, int b, int a, int il, int kt, int t, int makespan, bool arrived) {
  if (kt < bt.dim_ind) return true;
  const int D = makespan ? makespan : bt.max_T;
  const double p_rw = bt.p_rw[(size_t)t * (bt.max_T + 1) + D];  // prob_random_walk (learned_sampler.py:84-91), host table
  double u_gate;
  if (bt.rng_mode == 1) {
    u_gate = bt.tab_gate[step_index(bt, kt, t) * bt.A + a];
  } else {
    const uint4 w = philox4x32(bt.inst_base + (uint32_t)b, rng_ctr1(kt, t), (uint32_t)il, rng_ctr3(STREAM_GATE, 0), bt.seed_lo,
                               bt.seed_hi);
    u_gate = u64_to_f64(w.x, w.y);
  }
  return (!arrived && u_gate < p_rw) || (arrived && u_gate > bt.prob_after_goal);
}

// Per-warp context of one agent-step: the instance state and the agent's constants arrive in a few wide loads
// (inst_pack / agent_pack) instead of a chain of dependent scalar loads, and the instance's obstacles are staged once in
// shared memory as (x, y, rad_agent + rad_obstacle) so that the swept-circle loops never wait on global memory.
#define RO_OBS_MAX 64
struct StepCtx {
  int b, t, kt, makespan, a0, N, o0, nO;
  uint32_t inv_nO;  // 2^20 / nO + 1 for nO in 1..128, else 0
  double gx, gy, speed, speed2, merge2, rad;
  const double* sobs;  // shared memory [nO][3] when nO <= RO_OBS_MAX, else nullptr (global fallback)
};

__device__ __forceinline__ void obstacle_at(const Batch& bt, const StepCtx& cx, int o, double* ox, double* oy, double* rs) {
  if (cx.sobs != nullptr) {
    *ox = cx.sobs[3 * o];
    *oy = cx.sobs[3 * o + 1];
    *rs = cx.sobs[3 * o + 2];
  } else {
    const double* g = bt.obs_xyr + 3 * (size_t)(cx.o0 + o);
    *ox = __ldg(g);
    *oy = __ldg(g + 1);
    *rs = f64add(cx.rad, __ldg(g + 2));
  }
}
// 2^20 / n + 1 for n = 1..128 (entry 0 unused): reciprocal table of pair_div, in constant memory
struct PairInvTable {
  uint32_t v[129];
  constexpr PairInvTable() : v() {
    v[0] = 0u;
    for (int i = 1; i < 129; ++i) v[i] = (1u << 20) / (uint32_t)i + 1u;
  }
};
static __constant__ PairInvTable c_pair_inv = PairInvTable();

// p / nO for pair indices p <= 32 * nO (exact: 32 * nO^2 < 2^20 for nO <= 128) without an integer division
__device__ __forceinline__ int pair_div(const StepCtx& cx, int p) {
  return cx.inv_nO ? (int)(((uint32_t)p * cx.inv_nO) >> 20) : p / cx.nO;
}

// StaticObjects.collide_continuous_sphere (any over the instance's obstacles) for the lanes set in `cand` at once (lane l
// holds the segment start (vx, vy), the end (tx, ty) is uniform): the (candidate lane, obstacle) pairs are flattened over the warp, so a handful of candidates costs
// ceil(candidates * nO / 32) passes instead of nO passes with most lanes idle.  Returns the lanes whose segment collides.
__device__ __forceinline__ uint32_t hits_any_flat(const Batch& bt, const StepCtx& cx, uint32_t cand, double vx, double vy,
                                                  double tx, double ty, int lane) {
  const int nO = cx.nO;
  const int npairs = __popc(cand) * nO;
  uint32_t coll = 0u;
  for (int p0 = 0; p0 < npairs; p0 += 32) {
    const int p = p0 + lane;
    const bool act = p < npairs;
    const int ci = act ? pair_div(cx, p) : 0;
    uint32_t pm = cand;
    for (int q = 0; q < ci; ++q) pm &= pm - 1;  // drop the ci lowest set bits
    const int src = act ? (__ffs(pm) - 1) : 0;
    const double fx = __shfl_sync(0xFFFFFFFFu, vx, src), fy = __shfl_sync(0xFFFFFFFFu, vy, src);
    if (act) {
      double ox, oy, rs;
      obstacle_at(bt, cx, p - ci * nO, &ox, &oy, &rs);
      if (!sweep_far(fx, fy, tx, ty, ox, oy, rs) && sweep_static(fx, fy, tx, ty, ox, oy, rs)) coll |= 1u << src;
    }
  }
  return __reduce_or_sync(0xFFFFFFFFu, coll);
}

// ------------------------------------------------------------------------------------------------
// select: loc_pred for every agent
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int select_agent(const Batch& bt, const StepCtx& sc, const int a, const int lane,
                                                     const double cx, const double cy, const bool arrived, double* out_x,
                                                     double* out_y, const int gate_hint = -1) {
  const int b = sc.b, t = sc.t, kt = sc.kt;
  const int a0 = sc.a0, il = a - a0, N = sc.N, nO = sc.nO;
  const double speed = sc.speed, speed2 = sc.speed2, rad = sc.rad;
  const size_t si = step_index(bt, kt, t);
  const uint32_t c0 = bt.inst_base + (uint32_t)b, c1 = rng_ctr1(kt, t);
  // gate_hint: the caller has already evaluated the gate of this agent-step (0 / 1)
  const bool learned = gate_hint >= 0 ? gate_hint != 0 : gate_is_learned(bt, b, a, il, kt, t, sc.makespan, arrived);
  const int nL = learned ? bt.K : 0;
  const int C = nL + bt.max_attempt;
  (void)N;
  bool found = false;
  double lx = cx, ly = cy;  // fallback of sample_uniform_i: stay (learned_sampler.py:145)
  unsigned int count = 0;
  // the reference scans the candidates in order and stops at the first valid one: the learned candidates are evaluated
  // first and the random-walk ones (Philox + sin/cos in fp64) only when none of them survives — the draws are keyed, not
  // streamed, so skipping them changes nothing
  for (int phase = 0; phase < 2 && !found; ++phase) {
   const int c_lo = phase == 0 ? 0 : nL, c_hi = phase == 0 ? nL : C;
   for (int g0 = c_lo; g0 < c_hi && !found; g0 += 32) {
    const int c = g0 + lane;
    double x = cx, y = cy;
    bool pre = false;
    if (c < c_hi) {
      if (c < nL) {
        const float* s3 = (bt.tab_teacher != nullptr) ? bt.tab_teacher + ((si * bt.A + a) * (size_t)bt.K + c) * 3
                                                      : bt.samples + ((size_t)a * bt.K + c) * 3;
        // Y = cur (f64) + SAMPLES[:, :2] * SAMPLES[:, 2] (f32 product)   learned_sampler.py:114-121
        x = f64add(cx, (double)__fmul_rn(s3[0], s3[2]));
        y = f64add(cy, (double)__fmul_rn(s3[1], s3[2]));
      } else {
        const int att = c - nL;
        double u_mag, sn, cs;
        if (bt.rng_mode == 1) {
          const double* r = bt.tab_rw + ((si * bt.A + a) * (size_t)bt.max_attempt + att) * 3;
          u_mag = r[0]; sn = r[1]; cs = r[2];
        } else {
          uint4 w = philox4x32(c0, c1, (uint32_t)il, rng_ctr3(STREAM_RW, (uint32_t)att), bt.seed_lo, bt.seed_hi);
          u_mag = u64_to_f64(w.x, w.y);
          det_sincos_2pi(u64_to_f64(w.z, w.w), &sn, &cs);
        }
        // mag = speed * U ; loc = [sin, cos] * mag + cur   (learned_sampler.py:137-142)
        double mag = f64mul(speed, u_mag);
        x = f64add(f64mul(sn, mag), cx);
        y = f64add(f64mul(cs, mag), cy);
      }
      pre = bounds_ok(x, y, rad) && speed_ok(cx, cy, x, y, speed, speed2);
    }
    const uint32_t pre_mask = __ballot_sync(0xFFFFFFFFu, pre);
    // swept tests of the pre-checked candidates, flattened over (candidate, obstacle) pairs across the warp
    const int npairs = __popc(pre_mask) * nO;
    uint32_t coll = 0u;
    for (int p0 = 0; p0 < npairs; p0 += 32) {
      const int p = p0 + lane;
      const bool act = p < npairs;
      const int ci = act ? pair_div(sc, p) : 0;
      uint32_t pm = pre_mask;
      for (int q = 0; q < ci; ++q) pm &= pm - 1;  // drop the ci lowest set bits
      const int src = act ? (__ffs(pm) - 1) : 0;
      const double tx = __shfl_sync(0xFFFFFFFFu, x, src), ty = __shfl_sync(0xFFFFFFFFu, y, src);
      if (act) {
        double ox, oy, rs;
        obstacle_at(bt, sc, p - ci * nO, &ox, &oy, &rs);
        if (!sweep_far(cx, cy, tx, ty, ox, oy, rs) && sweep_static(cx, cy, tx, ty, ox, oy, rs)) coll |= 1u << src;
      }
    }
    coll = __reduce_or_sync(0xFFFFFFFFu, coll);
    const uint32_t valid = pre_mask & ~coll;
    if (valid) {
      const int first = __ffs(valid) - 1;
      count += __popc(pre_mask & ((2u << first) - 1u));
      lx = __shfl_sync(0xFFFFFFFFu, x, first);
      ly = __shfl_sync(0xFFFFFFFFu, y, first);
      found = true;
    } else {
      count += __popc(pre_mask);
    }
   }
  }
  if (lane == 0 && bt.tr_loc_pred != nullptr) {
    bt.tr_loc_pred[2 * (si * bt.A + a)] = lx;
    bt.tr_loc_pred[2 * (si * bt.A + a) + 1] = ly;
  }
  if (bt.tr_samples != nullptr && bt.samples != nullptr) {
    for (int i = lane; i < bt.K * 3; i += 32)
      bt.tr_samples[(si * bt.A + a) * (size_t)bt.K * 3 + i] = bt.samples[(size_t)a * bt.K * 3 + i];
  }
  *out_x = lx;
  *out_y = ly;
  return count;
}

// ------------------------------------------------------------------------------------------------
// merge: insert loc_pred into the agent's timed roadmap or merge it with a compatible vertex
// ------------------------------------------------------------------------------------------------
// layer sizes and this lane's candidate positions of V[t-1], V[t], V[t+1], loaded BEFORE the candidate selection runs so
// that their latency overlaps it (the kernel is bound by dependent global loads, not by arithmetic)
template <int W>
struct MergePrefetch {
  int nl, n_prev, n_t, n_next;
  double px[3][W], py[3][W];
};
template <int W>
__device__ __forceinline__ void merge_prefetch(const Batch& bt, int a, int t, int lane, MergePrefetch<W>& pf) {
  pf.nl = bt.nlayers[a];
  const int32_t* vc = bt.vcnt + (size_t)a * bt.L;
  pf.n_prev = vc[t - 1];
  pf.n_t = (pf.nl > t) ? vc[t] : 0;
  pf.n_next = (pf.nl > t + 1) ? vc[t + 1] : 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int n = l == 0 ? pf.n_prev : (l == 1 ? pf.n_t : pf.n_next);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int sl = w * 32 + lane;
      pf.px[l][w] = 0.0;
      pf.py[l][w] = 0.0;
      if (sl < n) {
        const size_t v = vslot(bt, a, t - 1 + l, sl);
        const double2 p = *reinterpret_cast<const double2*>(&bt.vpos[2 * v]);
        pf.px[l][w] = p.x;
        pf.py[l][w] = p.y;
      }
    }
  }
}

template <int W>
__device__ __forceinline__ unsigned int merge_agent(const Batch& bt, const StepCtx& sc, const int a, const int lane,
                                                    const double lx, const double ly, const MergePrefetch<W>& pf) {
  const int t = sc.t, kt = sc.kt;
  const double speed = sc.speed, speed2 = sc.speed2, merge2 = sc.merge2, rad = sc.rad;
  const double gx = sc.gx, gy = sc.gy;
  const int nl = pf.nl;
  const bool has_next = nl > t + 1;
  unsigned int count = 0;
  uint32_t P[W], Cm[W], Mg[W];
#pragma unroll
  for (int w = 0; w < W; ++w) { P[w] = 0u; Cm[w] = 0u; Mg[w] = 0u; }
  // parents: V[t-1] within speed and collision free, tested as (parent -> loc)   utils.py:58-69
#pragma unroll
  for (int w = 0; w < W; ++w) {
    if (w * 32 < pf.n_prev) {
      const bool in = w * 32 + lane < pf.n_prev;
      const bool cand = in && dist2(pf.px[0][w], pf.py[0][w], lx, ly) <= speed2;
      const uint32_t cm = __ballot_sync(0xFFFFFFFFu, cand);
      count += __popc(cm);
      P[w] = cm & ~hits_any_flat(bt, sc, cm, pf.px[0][w], pf.py[0][w], lx, ly, lane);
    }
  }
  // children: V[t+1], tested as (child -> loc)   utils.py:72-84
  if (has_next) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
      if (w * 32 < pf.n_next) {
        const bool in = w * 32 + lane < pf.n_next;
        const bool cand = in && dist2(pf.px[2][w], pf.py[2][w], lx, ly) <= speed2;
        const uint32_t cm = __ballot_sync(0xFFFFFFFFu, cand);
        count += __popc(cm);
        Cm[w] = cm & ~hits_any_flat(bt, sc, cm, pf.px[2][w], pf.py[2][w], lx, ly, lane);
      }
    }
  }
  const int n_t = pf.n_t;
  // merge candidates: V[t] within merge_distance   utils.py:86-89
#pragma unroll
  for (int w = 0; w < W; ++w) {
    if (w * 32 < n_t) {
      const bool cand = (w * 32 + lane < n_t) && dist2(pf.px[1][w], pf.py[1][w], lx, ly) <= merge2;
      Mg[w] = __ballot_sync(0xFFFFFFFFu, cand);
    }
  }
  // rules (utils.py:94-133), candidates in index order; everything below is warp-uniform
  double rx = lx, ry = ly;
  int target = -1;       // vertex of layer t that receives loc (replace) or -1
  bool add_edges = false, resolved = false;
  const double h_loc = dist2(lx, ly, gx, gy);
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t m = Mg[w];
    while (m && !resolved) {
      const int u = w * 32 + __ffs(m) - 1;
      m &= m - 1;
      const size_t v = vslot(bt, a, t, u);
      bool eq = true, sup = true, sub = true;
      for (int q = 0; q < W; ++q) {
        const uint32_t pu = bt.pmask[v * W + q], cu = bt.cmask[v * W + q];
        eq = eq && pu == P[q] && cu == Cm[q];
        sup = sup && (P[q] & ~pu) == 0u && (Cm[q] & ~cu) == 0u;
        sub = sub && (pu & ~P[q]) == 0u && (cu & ~Cm[q]) == 0u;
      }
      const double ux = bt.vpos[2 * v], uy = bt.vpos[2 * v + 1];
      if (eq) {
        if (h_loc < dist2(ux, uy, gx, gy)) target = u;  // replace u by loc
        else { rx = ux; ry = uy; }                       // abandon loc
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
        bt.vcnt[(size_t)a * bt.L + t] = n_t + 1;
        if (nl < t + 1) bt.nlayers[a] = t + 1;
      }
    }
    if (add_edges) {
      const uint32_t tbit = 1u << (target & 31);
      const int tw = target >> 5;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        {
          const uint32_t pu = resolved ? bt.pmask[v * W + w] : 0u;
          const uint32_t cu = resolved ? bt.cmask[v * W + w] : 0u;
          const int s = w * 32 + lane;
          // reductions without a return value: the warp does not wait for the old word (it is the only writer of its roadmap)
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
  }
  // goal test: valid_edge(loc_next, goal)   learned_sampler.py:182-183
  bool arr = false;
  if (bounds_ok(gx, gy, rad) && speed_ok(rx, ry, gx, gy, speed, speed2)) {
    count += 1;
    bool hit = false;
    for (int o = lane; o < sc.nO; o += 32) {
      double ox, oy, rs;
      obstacle_at(bt, sc, o, &ox, &oy, &rs);
      hit = hit || (!sweep_far(rx, ry, gx, gy, ox, oy, rs) && sweep_static(rx, ry, gx, gy, ox, oy, rs));
    }
    arr = !__any_sync(0xFFFFFFFFu, hit);
  }
  if (lane == 0) {
    bt.nxt[2 * a] = rx;
    bt.nxt[2 * a + 1] = ry;
    if (arr) bt.arrived[a] = 1;
    if (bt.tr_loc_next != nullptr) {
      const size_t si = step_index(bt, kt, t);
      bt.tr_loc_next[2 * (si * bt.A + a)] = rx;
      bt.tr_loc_next[2 * (si * bt.A + a) + 1] = ry;
    }
  }
  return count;
}
