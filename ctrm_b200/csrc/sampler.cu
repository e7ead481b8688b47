// Baseline samplers on the collision / edge kernels:
//   collide_points   StaticObjects.collide_sphere (environment/static_objects.py:44-57, :93-106): unit-box test, then a
//                    static circle-vs-circle test against every obstacle of the instance.  The reference delegates that
//                    test to FCL (python-fcl over FCL 0.5.0, Dockerfile:63), which is absent here; restated from FCL's
//                    published sphereSphereIntersect: collide <=> !(||c1 - c2|| > r1 + r2), norm in fp64.
//   pair_adjacency   the all-pairs valid_move loops of TimedRoadmap.append_fixed_strucutre (roadmap/timed_roadmap.py:186-193),
//                    get_common_roadmaps (roadmap/utils.py:114-120), TimedRoadmap.extend_until (timed_roadmap.py:259-265) and
//                    add_parents (timed_roadmap.py:116-131): valid_move (roadmap/utils.py:18-42) = L-inf pre-check, exact L2
//                    check, swept circle against the obstacles (collision_check.cpp:16-58) — one bit per (row, column).
//   adjacency_csr    the adjacency lists in the reference's order ([i] first, then the neighbours ascending) by popcount,
//                    exclusive prefix scan and a ballot-ranked scatter.
// fp64 in the reference's operation order; verdicts, edge lists and the collision counter are bit-exact.
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// static point test
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) collide_points_kernel(const double2* __restrict__ pos, const double* __restrict__ rad,
                                                             const int32_t* __restrict__ pt_inst, long long n,
                                                             const double* __restrict__ obs_xyr,
                                                             const int32_t* __restrict__ obs_off, uint8_t* __restrict__ verdict) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 p = pos[i];
  const double r = rad[i];
  // static_objects.py:103-104: max(pos + rad) > 1 or min(pos - rad) < 0
  bool hit = fmax(f64add(p.x, r), f64add(p.y, r)) > 1.0 || fmin(f64sub(p.x, r), f64sub(p.y, r)) < 0.0;
  if (!hit) {
    const int b = pt_inst ? pt_inst[i] : 0;
    for (int o = obs_off[b]; o < obs_off[b + 1]; ++o) {
      const double dx = f64sub(__ldg(&obs_xyr[3 * o]), p.x), dy = f64sub(__ldg(&obs_xyr[3 * o + 1]), p.y);
      const double len = __dsqrt_rn(f64add(f64add(f64mul(dx, dx), f64mul(dy, dy)), 0.0));  // 2-D points padded with z = 0
      hit = hit || !(len > f64add(r, __ldg(&obs_xyr[3 * o + 2])));
    }
  }
  verdict[i] = hit ? 1 : 0;
}

int launch_collide_points(const double* d_pos, const double* d_rad, const int32_t* d_pt_inst, int64_t n, const double* d_obs_xyr,
                          const int32_t* d_obs_off, uint8_t* d_verdict, cudaStream_t st) {
  if (n <= 0) return 0;
  collide_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const double2*>(d_pos), d_rad, d_pt_inst, n,
                                                                    d_obs_xyr, d_obs_off, d_verdict);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// all-pairs valid_move
// ------------------------------------------------------------------------------------------------
// One warp per (set, row, chunk of PA_WORDS 32-column words); lane = column.  The row position stays in registers, the column
// positions are read coalesced (16 B per lane) and are shared through L1 by the warps of neighbouring rows.  The L-inf / L2
// pre-checks reject all but the ~(2 speed)^2 fraction of the pairs; the survivors run the swept test over the obstacles.
// symmetric: rows == columns, the pair {i, j} is always tested as (min -> max) like the reference's `for j in range(i + 1, n)`
// loops, the diagonal stays clear, and the collision counter counts each pair once (j > i).
#define PA_WORDS 8
__global__ void __launch_bounds__(256) pair_adjacency_kernel(const double2* __restrict__ rows, const int32_t* __restrict__ row_off,
                                                             const double2* __restrict__ cols, const int32_t* __restrict__ col_off,
                                                             const int32_t* __restrict__ set_inst, const double* __restrict__ speed,
                                                             const double* __restrict__ speed2, const double* __restrict__ rad,
                                                             const double* __restrict__ obs_xyr, const int32_t* __restrict__ obs_off,
                                                             int symmetric, const int64_t* __restrict__ bits_off,
                                                             const int64_t* __restrict__ item_off, int n_sets,
                                                             uint32_t* __restrict__ bits, unsigned long long* __restrict__ cnt) {
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= item_off[n_sets]) return;
  // the set of this work item: binary search over the per-set item prefix
  int lo = 0, hi = n_sets - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (item_off[mid] <= item) lo = mid; else hi = mid - 1;
  }
  const int s = lo;
  const int r0 = row_off[s], nr = row_off[s + 1] - r0, c0 = col_off[s], nc = col_off[s + 1] - c0;
  (void)nr;
  const int W = (nc + 31) >> 5, chunks = (W + PA_WORDS - 1) / PA_WORDS;
  const long long local = item - item_off[s];
  const int i = (int)(local / chunks), ch = (int)(local - (long long)i * chunks);
  const double2 pi = rows[r0 + i];
  const double sp = speed[s], sp2 = speed2[s], ra = rad[s];
  const int b = set_inst ? set_inst[s] : 0;
  const int o0 = obs_off[b], o1 = obs_off[b + 1];
  uint32_t* out = bits + bits_off[s] + (size_t)i * W;
  unsigned int counted = 0;
  for (int w = ch * PA_WORDS; w < min(W, (ch + 1) * PA_WORDS); ++w) {
    const int j = w * 32 + lane;
    bool ok = false;
    if (j < nc && !(symmetric && j == i)) {
      const double2 pj = cols[c0 + j];
      if (speed_ok(pi.x, pi.y, pj.x, pj.y, sp, sp2)) {  // |p1 - p2| is symmetric: the pre-checks do not depend on the order
        const bool fwd = !symmetric || i < j;
        const double fx = fwd ? pi.x : pj.x, fy = fwd ? pi.y : pj.y, tx = fwd ? pj.x : pi.x, ty = fwd ? pj.y : pi.y;
        if (fwd) ++counted;  // one collide_continuous_sphere call per pair (static_objects.py:72-75)
        ok = !any_obstacle_hit(fx, fy, tx, ty, ra, obs_xyr, o0, o1);
      }
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
    if (lane == 0) out[w] = m;
  }
  counted = __reduce_add_sync(0xFFFFFFFFu, counted);
  if (lane == 0 && counted && cnt != nullptr) atomicAdd(&cnt[s], (unsigned long long)counted);
}

int launch_pair_adjacency(const double* d_rows, const int32_t* d_row_off, const double* d_cols, const int32_t* d_col_off,
                          int n_sets, const int32_t* d_set_inst, const double* d_speed, const double* d_speed2, const double* d_rad,
                          const double* d_obs_xyr, const int32_t* d_obs_off, int symmetric, const int64_t* d_bits_off,
                          const int64_t* d_item_off, int64_t n_items, uint32_t* d_bits, int64_t* d_cnt, cudaStream_t st) {
  if (n_sets <= 0 || n_items <= 0) return 0;
  const long long blocks = (n_items + 7) / 8;  // 8 warps per block
  if (blocks > 0x7FFFFFFFLL) return ctrm_fail(-1, "pair_adjacency: too many rows");
  pair_adjacency_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const double2*>(d_rows), d_row_off,
                                                         reinterpret_cast<const double2*>(d_cols), d_col_off, d_set_inst, d_speed,
                                                         d_speed2, d_rad, d_obs_xyr, d_obs_off, symmetric, d_bits_off, d_item_off,
                                                         n_sets, d_bits, reinterpret_cast<unsigned long long*>(d_cnt));
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bit rows -> CSR
// ------------------------------------------------------------------------------------------------
// degree of every row: popcount of its words (+ 1 for the leading self entry of the reference's `adjs = [[i] ...]`)
__global__ void __launch_bounds__(256) adj_degree_kernel(const uint32_t* __restrict__ bits, const int64_t* __restrict__ bits_off,
                                                         const int32_t* __restrict__ row_off, const int32_t* __restrict__ col_off,
                                                         int n_sets, int self_first, int64_t* __restrict__ deg) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= row_off[n_sets]) return;
  int lo = 0, hi = n_sets - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (row_off[mid] <= row) lo = mid; else hi = mid - 1;
  }
  const int s = lo, W = (col_off[s + 1] - col_off[s] + 31) >> 5;
  const uint32_t* r = bits + bits_off[s] + (size_t)(row - row_off[s]) * W;
  int c = 0;
  for (int w = lane; w < W; w += 32) c += __popc(r[w]);
  c = __reduce_add_sync(0xFFFFFFFFu, c);
  if (lane == 0) deg[row] = c + (self_first ? 1 : 0);
}

// in-place exclusive scan of n 64-bit counts by one block; the total lands in v[n]
__global__ void __launch_bounds__(1024) scan64_kernel(int64_t* __restrict__ v, int n) {
  __shared__ long long wsum[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const long long x0 = i < n ? v[i] : 0;
    long long x = x0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(0xFFFFFFFFu, w, d);
        if (lane >= d) w += y;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const long long excl = x - x0 + (warp > 0 ? wsum[warp - 1] : 0) + carry;
    if (i < n) v[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + x0;
    __syncthreads();
  }
  if (threadIdx.x == 0) v[n] = carry;
}

// one warp per row: [row (local index)] first, then the set bits in ascending order
__global__ void __launch_bounds__(256) adj_fill_kernel(const uint32_t* __restrict__ bits, const int64_t* __restrict__ bits_off,
                                                       const int32_t* __restrict__ row_off, const int32_t* __restrict__ col_off,
                                                       int n_sets, int self_first, const int64_t* __restrict__ eptr,
                                                       int32_t* __restrict__ eidx) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= row_off[n_sets]) return;
  int lo = 0, hi = n_sets - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (row_off[mid] <= row) lo = mid; else hi = mid - 1;
  }
  const int s = lo, W = (col_off[s + 1] - col_off[s] + 31) >> 5, il = row - row_off[s];
  const uint32_t* r = bits + bits_off[s] + (size_t)il * W;
  int64_t at = eptr[row];
  if (self_first) {
    if (lane == 0) eidx[at] = il;
    ++at;
  }
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    const uint32_t m = w < W ? r[w] : 0u;
    int c = __popc(m), x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= d) x += y;
    }
    int64_t p = at + (x - c);
    uint32_t mm = m;
    while (mm) {
      eidx[p++] = w * 32 + (__ffs(mm) - 1);
      mm &= mm - 1;
    }
    at += __shfl_sync(0xFFFFFFFFu, x, 31);
  }
}

int launch_adjacency_csr(const uint32_t* d_bits, const int64_t* d_bits_off, const int32_t* d_row_off, const int32_t* d_col_off,
                         int n_sets, int n_rows, int self_first, int64_t* d_eptr, int32_t* d_eidx, cudaStream_t st) {
  if (n_sets <= 0 || n_rows <= 0) return 0;
  const unsigned blocks = (unsigned)((n_rows + 7) / 8);
  if (d_eidx == nullptr) {  // pass 1: degrees -> exclusive scan (eptr[n_rows] = number of entries)
    adj_degree_kernel<<<blocks, 256, 0, st>>>(d_bits, d_bits_off, d_row_off, d_col_off, n_sets, self_first, d_eptr);
    CTRM_CUDA_OK(cudaGetLastError());
    scan64_kernel<<<1, 1024, 0, st>>>(d_eptr, n_rows);
    CTRM_CUDA_OK(cudaGetLastError());
  } else {  // pass 2: scatter
    adj_fill_kernel<<<blocks, 256, 0, st>>>(d_bits, d_bits_off, d_row_off, d_col_off, n_sets, self_first, d_eptr, d_eidx);
    CTRM_CUDA_OK(cudaGetLastError());
  }
  return 0;
}
