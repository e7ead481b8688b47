// Batched swept-circle collision checks.
//   reference: continuousCollideSpheres (cpp_helper/sphere_collision_check_wrapper/src/collision_check.cpp:16-58)
//   driven by StaticObjects.collide_continuous_sphere (environment/static_objects.py:59-75, :108-138) and
//   continuous_collide_spheres (environment/fcl_utils.py:185-212).
// fp64, the reference's operation order, every * and + rounded separately (no FMA) -> verdicts are bit-exact.
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

// one thread per swept circle; obstacles of the segment's instance are walked from L1/L2 (24 B each, shared by
// every segment of the instance).  Survivor compaction: pass 1 writes verdicts and per-block survivor counts,
// pass 2 scans the block counts, pass 3 scatters survivor indices in ascending order (ballot + popc prefix).
__global__ void __launch_bounds__(256) collide_segments_kernel(const double2* __restrict__ from, const double2* __restrict__ to,
                                                               const double* __restrict__ rad,
                                                               const int32_t* __restrict__ seg_inst, long long n,
                                                               const double* __restrict__ obs_xyr,
                                                               const int32_t* __restrict__ obs_off,
                                                               uint8_t* __restrict__ verdict,
                                                               int32_t* __restrict__ block_cnt) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false, in = s < n;
  if (in) {
    double2 f = from[s], t = to[s];
    double r = rad[s];
    int b = seg_inst ? seg_inst[s] : 0;
    int o0 = obs_off[b], o1 = obs_off[b + 1];
    hit = any_obstacle_hit(f.x, f.y, t.x, t.y, r, obs_xyr, o0, o1);
    verdict[s] = hit ? 1 : 0;
  }
  if (block_cnt != nullptr) {
    int c = __syncthreads_count(in && !hit);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = c;
  }
}

// exclusive scan of the per-block survivor counts (single block, decoupled from the data pass)
__global__ void __launch_bounds__(1024) scan_block_counts_kernel(int32_t* __restrict__ block_cnt, int nblocks,
                                                                 int32_t* __restrict__ total) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nblocks) ? block_cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xFFFFFFFFu, w, d);
        if (lane >= d) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = x - v + (warp > 0 ? warp_sums[warp - 1] : 0) + carry;
    if (i < nblocks) block_cnt[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(256) scatter_survivors_kernel(const uint8_t* __restrict__ verdict, long long n,
                                                                const int32_t* __restrict__ block_off,
                                                                int32_t* __restrict__ survivors) {
  __shared__ int warp_cnt[8];
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool keep = (s < n) && (verdict[s] == 0);
  uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(m);
  __syncthreads();
  int off = block_off[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += warp_cnt[w];
  if (keep) survivors[off + __popc(m & ((1u << lane) - 1u))] = (int32_t)s;
}

__global__ void __launch_bounds__(256) collide_pairs_kernel(const double* __restrict__ f1, const double* __restrict__ t1,
                                                            const double* __restrict__ r1, const double* __restrict__ f2,
                                                            const double* __restrict__ t2, const double* __restrict__ r2,
                                                            int dim, long long n, uint8_t* __restrict__ verdict) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double a[3], b[3], c[3], d[3];
  for (int i = 0; i < dim; ++i) {
    a[i] = f1[s * dim + i];
    b[i] = t1[s * dim + i];
    c[i] = f2[s * dim + i];
    d[i] = t2[s * dim + i];
  }
  verdict[s] = sweep_pair(a, b, r1[s], c, d, r2[s], dim) ? 1 : 0;
}

int launch_collide_segments(const double* d_from, const double* d_to, const double* d_rad, const int32_t* d_seg_inst,
                            int64_t n, const double* d_obs_xyr, const int32_t* d_obs_off, uint8_t* d_verdict,
                            int32_t* d_survivors, int32_t* d_n_survivors, cudaStream_t st) {
  if (n <= 0) {
    if (d_n_survivors) CTRM_CUDA_OK(cudaMemsetAsync(d_n_survivors, 0, sizeof(int32_t), st));
    return 0;
  }
  long long nb = (n + 255) / 256;
  if (nb > 0x7FFFFFFFLL) return ctrm_fail(-1, "too many segments");
  int32_t* d_block_cnt = nullptr;
  if (d_survivors && keep_async_pool_warm()) return -2;
  if (d_survivors) CTRM_CUDA_OK(cudaMallocAsync(&d_block_cnt, (size_t)nb * sizeof(int32_t), st));
  collide_segments_kernel<<<(unsigned)nb, 256, 0, st>>>(reinterpret_cast<const double2*>(d_from),
                                                       reinterpret_cast<const double2*>(d_to), d_rad, d_seg_inst, n,
                                                       d_obs_xyr, d_obs_off, d_verdict, d_block_cnt);
  CTRM_CUDA_OK(cudaGetLastError());
  if (d_survivors) {
    scan_block_counts_kernel<<<1, 1024, 0, st>>>(d_block_cnt, (int)nb, d_n_survivors);
    CTRM_CUDA_OK(cudaGetLastError());
    scatter_survivors_kernel<<<(unsigned)nb, 256, 0, st>>>(d_verdict, n, d_block_cnt, d_survivors);
    CTRM_CUDA_OK(cudaGetLastError());
    CTRM_CUDA_OK(cudaFreeAsync(d_block_cnt, st));
  }
  return 0;
}

int launch_collide_pairs(const double* f1, const double* t1, const double* r1, const double* f2, const double* t2,
                         const double* r2, int dim, int64_t n, uint8_t* verdict, cudaStream_t st) {
  if (n <= 0) return 0;
  if (dim != 2 && dim != 3) return ctrm_fail(-1, "dim must be 2 or 3");
  collide_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f1, t1, r1, f2, t2, r2, dim, n, verdict);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
