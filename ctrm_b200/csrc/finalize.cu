// Finalisation of the timed roadmaps and CSR export.
//   reference: format_trms (roadmap_learned/utils.py:140-156) -> TimedRoadmap.extend_until
//   (roadmap/timed_roadmap.py:248-272); append_goals (utils.py:159-174) -> TimedRoadmap.append_sample /
//   add_parents / add_children (timed_roadmap.py:67-150).  Adjacency is held as bitmasks over the next layer's
//   slots; the CSR (V positions, edge offsets, edge targets in ascending index order) comes out of a popcount
//   prefix scan.
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

#define FZ_WARPS 4

// valid_move (roadmap/utils.py:18-42) without the bounds test; *count gets +1 when the swept check is reached
__device__ __forceinline__ bool valid_move_dev(const Batch& bt, double ax, double ay, double bx, double by, double speed,
                                               double speed2, double rad, int o0, int o1, unsigned int* count) {
  if (!speed_ok(ax, ay, bx, by, speed, speed2)) return false;
  *count += 1;
  return !any_obstacle_hit_f4(ax, ay, bx, by, rad, bt.obs_xyr, bt.obs_f4, o0, o1);
}

// T = max_i len(trm_i.V) - 1 per instance   (utils.py:147)
__global__ void __launch_bounds__(FZ_WARPS * 32) tfin_kernel(Batch bt, int32_t* __restrict__ tfin) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= bt.B) return;
  const int a0 = bt.agent_off[b], N = bt.agent_off[b + 1] - a0;
  int m = 0;
  for (int i = lane; i < N; i += 32) m = max(m, bt.nlayers[a0 + i]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
  if (lane == 0) tfin[b] = m - 1;
}

// extend_until(T + 1): all-pairs valid_move inside the last layer (self loop + symmetric neighbours), then the
// last layer is replicated until the roadmap has T + 2 layers.  One warp per agent.
__global__ void __launch_bounds__(FZ_WARPS * 32) extend_kernel(Batch bt, const int32_t* __restrict__ tfin) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= bt.A) return;
  const int b = bt.agent_inst[a];
  const int o0 = bt.obs_off[b], o1 = bt.obs_off[b + 1];
  const double speed = bt.speeds[a], speed2 = bt.speeds2[a], rad = bt.rads[a];
  const int W = bt.W;
  const int last = bt.nlayers[a] - 1;
  const int n = bt.vcnt[(size_t)a * bt.L + last];
  const int target = tfin[b] + 1;
  unsigned int count = 0;
  for (int i = 0; i < n; ++i) {
    const size_t vi = vslot(bt, a, last, i);
    const double ix = bt.vpos[2 * vi], iy = bt.vpos[2 * vi + 1];
    if (lane == 0) bt.cmask[vi * W + (i >> 5)] |= 1u << (i & 31);  // adjs[i] = [i]
    for (int j0 = i + 1; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      bool ok = false;
      size_t vj = 0;
      if (j < n) {
        vj = vslot(bt, a, last, j);
        ok = valid_move_dev(bt, ix, iy, bt.vpos[2 * vj], bt.vpos[2 * vj + 1], speed, speed2, rad, o0, o1, &count);
      }
      // i -> j bits (several lanes may share a word: combine by ballot, lane 0 of each word writes)
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
      if (ok) bt.cmask[vj * W + (i >> 5)] |= 1u << (i & 31);  // adjs[j].append(i): one lane per j
      if (m) {
        // bits of j = j0 + l for l in mask; they span at most two words of adjs[i]
        const int w0 = j0 >> 5, sh = j0 & 31;
        if (lane == 0) {
          bt.cmask[vi * W + w0] |= m << sh;
          if (sh && (m >> (32 - sh)) && w0 + 1 < W) bt.cmask[vi * W + w0 + 1] |= m >> (32 - sh);
        }
      }
      __syncwarp();
    }
    __syncwarp();
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) count += __shfl_xor_sync(0xFFFFFFFFu, count, d);
  if (lane == 0 && count) atomicAdd(&bt.cnt_collide[b], (unsigned long long)count);
  __syncwarp();
  // replicate: V[l] = V[last], E[l-1] = adjs, E[l] = [[]...]   (timed_roadmap.py:263-272)
  for (int l = last + 1; l <= target; ++l) {
    for (int s = lane; s < n; s += 32) {
      const size_t src = vslot(bt, a, last, s), dst = vslot(bt, a, l, s), prv = vslot(bt, a, l - 1, s);
      bt.vpos[2 * dst] = bt.vpos[2 * src];
      bt.vpos[2 * dst + 1] = bt.vpos[2 * src + 1];
      for (int w = 0; w < W; ++w) {
        const uint32_t adj = bt.cmask[src * W + w];
        if (l - 1 != last) bt.cmask[prv * W + w] = adj;
        bt.pmask[dst * W + w] = adj;  // adjacency is symmetric: parents of copy s are adjs[s]
      }
    }
    if (lane == 0) bt.vcnt[(size_t)a * bt.L + l] = n;
  }
  if (lane == 0) bt.nlayers[a] = target + 1;
}

// append_goals: one warp per (agent, layer t >= 1).  The goal becomes the LAST vertex of V[t]; parents are every
// valid u in V[t-1] (including the goal appended at t-1), children every valid u in V[t+1] (before its goal).
__global__ void __launch_bounds__(FZ_WARPS * 32) goals_kernel(Batch bt) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int a = (int)(wid / bt.L), t = (int)(wid % bt.L);
  if (a >= bt.A) return;
  const int nl = bt.nlayers[a];
  if (t < 1 || t >= nl) return;
  const int b = bt.agent_inst[a];
  const int o0 = bt.obs_off[b], o1 = bt.obs_off[b + 1];
  const double speed = bt.speeds[a], speed2 = bt.speeds2[a], rad = bt.rads[a];
  const double gx = bt.goals[2 * a], gy = bt.goals[2 * a + 1];
  const int W = bt.W;
  const int idx = bt.vcnt[(size_t)a * bt.L + t];
  if (idx >= bt.S) return;
  const size_t vg = vslot(bt, a, t, idx);
  const uint32_t gbit = 1u << (idx & 31);
  const int gw = idx >> 5;
  unsigned int count = 0;
  if (lane == 0) {
    bt.vpos[2 * vg] = gx;
    bt.vpos[2 * vg + 1] = gy;
  }
  {  // parents
    const int n_prev = bt.vcnt[(size_t)a * bt.L + (t - 1)];
    const int n_all = n_prev + ((t - 1 >= 1) ? 1 : 0);  // + the goal appended at t-1
    for (int s0 = 0; s0 < n_all; s0 += 32) {
      const int s = s0 + lane;
      bool ok = false;
      if (s < n_all) {
        double ux = gx, uy = gy;
        if (s < n_prev) {
          const size_t v = vslot(bt, a, t - 1, s);
          ux = bt.vpos[2 * v];
          uy = bt.vpos[2 * v + 1];
        }
        ok = valid_move_dev(bt, ux, uy, gx, gy, speed, speed2, rad, o0, o1, &count);
        if (ok) atomicOr(&bt.cmask[vslot(bt, a, t - 1, s) * W + gw], gbit);
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
      if (lane == 0 && m) atomicOr(&bt.pmask[vg * W + (s0 >> 5)], m);
    }
  }
  if (t + 1 < nl) {  // children
    const int n_next = bt.vcnt[(size_t)a * bt.L + (t + 1)];
    for (int s0 = 0; s0 < n_next; s0 += 32) {
      const int s = s0 + lane;
      bool ok = false;
      if (s < n_next) {
        const size_t v = vslot(bt, a, t + 1, s);
        ok = valid_move_dev(bt, gx, gy, bt.vpos[2 * v], bt.vpos[2 * v + 1], speed, speed2, rad, o0, o1, &count);
        if (ok) atomicOr(&bt.pmask[v * W + gw], gbit);
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
      if (lane == 0 && m) atomicOr(&bt.cmask[vg * W + (s0 >> 5)], m);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) count += __shfl_xor_sync(0xFFFFFFFFu, count, d);
  if (lane == 0 && count) atomicAdd(&bt.cnt_collide[b], (unsigned long long)count);
}

__global__ void goals_count_kernel(Batch bt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int a = (int)(i / bt.L), t = (int)(i % bt.L);
  if (a >= bt.A) return;
  if (t >= 1 && t < bt.nlayers[a] && bt.vcnt[(size_t)a * bt.L + t] < bt.S) bt.vcnt[(size_t)a * bt.L + t] += 1;
}

// ------------------------------------------------------------------------------------------------
// CSR: per-agent counts -> exclusive scan over agents -> per-agent fill
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FZ_WARPS * 32) csr_count_kernel(Batch bt, int64_t* __restrict__ agent_nv,
                                                                  int64_t* __restrict__ agent_ne) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= bt.A) return;
  const int nl = bt.nlayers[a], W = bt.W;
  long long nv = 0, ne = 0;
  for (int t = 0; t < nl; ++t) {
    const int n = bt.vcnt[(size_t)a * bt.L + t];
    nv += n;
    for (int s = lane; s < n; s += 32) {
      const size_t v = vslot(bt, a, t, s);
      for (int w = 0; w < W; ++w) ne += __popc(bt.cmask[v * W + w]);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ne += __shfl_xor_sync(0xFFFFFFFFu, ne, d);
  if (lane == 0) {
    agent_nv[a] = nv;
    agent_ne[a] = ne;
  }
}

// single-block exclusive scan of two int64 arrays of length n; totals[0..1] = sums, totals[2] = sum of nlayers
__global__ void __launch_bounds__(1024) csr_scan_kernel(int64_t* __restrict__ nv, int64_t* __restrict__ ne, int n,
                                                        const int32_t* __restrict__ nlayers, int64_t* __restrict__ totals) {
  __shared__ long long wsum[2][32];
  __shared__ long long carry[3];
  if (threadIdx.x < 3) carry[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    long long v0 = (i < n) ? nv[i] : 0, v1 = (i < n) ? ne[i] : 0, v2 = (i < n) ? nlayers[i] : 0;
    long long x0 = v0, x1 = v1, x2 = v2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      long long y0 = __shfl_up_sync(0xFFFFFFFFu, x0, d), y1 = __shfl_up_sync(0xFFFFFFFFu, x1, d);
      if (lane >= d) { x0 += y0; x1 += y1; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x2 += __shfl_xor_sync(0xFFFFFFFFu, x2, d);
    if (lane == 31) { wsum[0][warp] = x0; wsum[1][warp] = x1; }
    if (lane == 0 && x2) atomicAdd((unsigned long long*)&carry[2], (unsigned long long)x2);
    __syncthreads();
    if (warp == 0) {
      long long w0 = wsum[0][lane], w1 = wsum[1][lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        long long y0 = __shfl_up_sync(0xFFFFFFFFu, w0, d), y1 = __shfl_up_sync(0xFFFFFFFFu, w1, d);
        if (lane >= d) { w0 += y0; w1 += y1; }
      }
      wsum[0][lane] = w0;
      wsum[1][lane] = w1;
    }
    __syncthreads();
    const long long e0 = x0 - v0 + (warp ? wsum[0][warp - 1] : 0) + carry[0];
    const long long e1 = x1 - v1 + (warp ? wsum[1][warp - 1] : 0) + carry[1];
    if (i < n) { nv[i] = e0; ne[i] = e1; }
    __syncthreads();
    if (threadIdx.x == 1023) { carry[0] = e0 + v0; carry[1] = e1 + v1; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; totals[2] = carry[2]; }
}

// COMPACT = false: layer_ptr (i32, stride L per agent), eptr (i64), eidx (i32) — what the planner kernel and the reference-shaped
// views read.  COMPACT = true: vertices per layer, out-degree per vertex and successor index as single BYTES (a layer holds at
// most S <= 128 vertices), which is what travels to the host: 0.75 MB instead of 1.46 MB per aamas22 instance.
template <bool COMPACT>
__global__ void __launch_bounds__(FZ_WARPS * 32) csr_fill_kernel(Batch bt, const int64_t* __restrict__ agent_v0,
                                                                 const int64_t* __restrict__ agent_e0,
                                                                 const int64_t* __restrict__ totals,
                                                                 void* __restrict__ layer_out, double* __restrict__ pos,
                                                                 void* __restrict__ eptr_out, void* __restrict__ eidx_out) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= bt.A) return;
  int32_t* layer_ptr = reinterpret_cast<int32_t*>(layer_out);
  uint8_t* layer_cnt = reinterpret_cast<uint8_t*>(layer_out);
  int64_t* eptr = reinterpret_cast<int64_t*>(eptr_out);
  uint8_t* deg8 = reinterpret_cast<uint8_t*>(eptr_out);
  int32_t* eidx = reinterpret_cast<int32_t*>(eidx_out);
  uint8_t* eidx8 = reinterpret_cast<uint8_t*>(eidx_out);
  const int nl = bt.nlayers[a], W = bt.W;
  long long vb = agent_v0[a], eb = agent_e0[a];
  for (int t = 0; t < bt.L; ++t) {
    const int n = t < nl ? bt.vcnt[(size_t)a * bt.L + t] : 0;
    if (lane == 0) {
      if (COMPACT) layer_cnt[(size_t)a * bt.L + t] = (uint8_t)n;
      else layer_ptr[(size_t)a * bt.L + t] = (int32_t)vb;
    }
    if (t >= nl) continue;
    for (int s0 = 0; s0 < n; s0 += 32) {
      const int s = s0 + lane;
      int deg = 0;
      size_t v = 0;
      if (s < n) {
        v = vslot(bt, a, t, s);
        for (int w = 0; w < W; ++w) deg += __popc(bt.cmask[v * W + w]);
      }
      int x = deg;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
      }
      const int tot = __shfl_sync(0xFFFFFFFFu, x, 31);
      if (s < n) {
        long long e = eb + (x - deg);
        *reinterpret_cast<double2*>(&pos[2 * (vb + s)]) = *reinterpret_cast<const double2*>(&bt.vpos[2 * v]);
        if (COMPACT) deg8[vb + s] = (uint8_t)deg;
        else eptr[vb + s] = e;
        for (int w = 0; w < W; ++w) {
          uint32_t m = bt.cmask[v * W + w];
          while (m) {
            const int j = w * 32 + __ffs(m) - 1;
            if (COMPACT) eidx8[e++] = (uint8_t)j;
            else eidx[e++] = j;
            m &= m - 1;
          }
        }
      }
      eb += tot;
    }
    vb += n;
  }
  if (!COMPACT && a == bt.A - 1 && lane == 0) {
    layer_ptr[(size_t)bt.A * bt.L] = (int32_t)totals[0];
    eptr[totals[0]] = totals[1];
  }
}

int launch_finalize(const Batch& bt, int32_t* d_tfin, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  tfin_kernel<<<(bt.B + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, 0, st>>>(bt, d_tfin);
  CTRM_CUDA_OK(cudaGetLastError());
  extend_kernel<<<(bt.A + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, 0, st>>>(bt, d_tfin);
  CTRM_CUDA_OK(cudaGetLastError());
  long long warps = (long long)bt.A * bt.L;
  goals_kernel<<<(unsigned)((warps + FZ_WARPS - 1) / FZ_WARPS), FZ_WARPS * 32, 0, st>>>(bt);
  CTRM_CUDA_OK(cudaGetLastError());
  goals_count_kernel<<<(unsigned)((warps + 255) / 256), 256, 0, st>>>(bt);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_csr_count(const Batch& bt, int64_t* d_agent_nv, int64_t* d_agent_ne, int64_t* d_totals, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  csr_count_kernel<<<(bt.A + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, 0, st>>>(bt, d_agent_nv, d_agent_ne);
  CTRM_CUDA_OK(cudaGetLastError());
  csr_scan_kernel<<<1, 1024, 0, st>>>(d_agent_nv, d_agent_ne, bt.A, bt.nlayers, d_totals);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_csr_fill(const Batch& bt, const int64_t* d_agent_v0, const int64_t* d_agent_e0, const int64_t* d_totals,
                    int32_t* d_layer_ptr, double* d_pos, int64_t* d_eptr, int32_t* d_eidx, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  csr_fill_kernel<false><<<(bt.A + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, 0, st>>>(bt, d_agent_v0, d_agent_e0, d_totals,
                                                                                  d_layer_ptr, d_pos, d_eptr, d_eidx);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_csr_fill_compact(const Batch& bt, const int64_t* d_agent_v0, const int64_t* d_agent_e0, const int64_t* d_totals,
                            uint8_t* d_layer_cnt, double* d_pos, uint8_t* d_deg, uint8_t* d_eidx, cudaStream_t st) {
  if (bt.A <= 0) return 0;
  if (bt.S > 255) return ctrm_fail(-1, "compact export needs at most 255 vertices per layer");
  csr_fill_kernel<true><<<(bt.A + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, 0, st>>>(bt, d_agent_v0, d_agent_e0, d_totals,
                                                                                 d_layer_cnt, d_pos, d_deg, d_eidx);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
