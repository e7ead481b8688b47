// Training-time pairwise features: numba get_arr_others_info + get_normed_vec_mag
// (learning/compiled.py:11-97) and get_self_info (compiled.py:100-125) for all T time steps of a demonstration at once,
// as Dataset.map_instance_to_tensor builds them (learning/dataset.py:135-160).  float64 arithmetic in the reference's
// order (square, add, sqrt, divide - each rounded once, no FMA), one final cast to float32 (torch .float()): bit-exact.
#include "common.cuh"
#include "kernels.h"

__device__ __forceinline__ void normed_vec_mag64(double x, double y, float* o) {
  const double mag = __dsqrt_rn(f64add(f64mul(x, x), f64mul(y, y)));
  const double d = (mag == 0.0) ? 1.0 : mag;
  o[0] = (float)__ddiv_rn(x, d);
  o[1] = (float)__ddiv_rn(y, d);
  o[2] = (float)mag;
}

// thread = (t, i, j): others_info[t][i*N + j][0:11] = {cur_j - cur_i, goal_j - cur_i, prev_j - cur_i (unit vector, norm
// each), rad_j, speed_j}; self_info[t][i] = columns 3..10 of the (i, i) row
__global__ void __launch_bounds__(256) others_info_kernel(const double* __restrict__ cur, const double* __restrict__ goals,
                                                          const double* __restrict__ prev, const double* __restrict__ rads,
                                                          const double* __restrict__ speeds, int T, int N,
                                                          float* __restrict__ info, float* __restrict__ self_info) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)T * N * N;
  if (q >= total) return;
  const int j = (int)(q % N), i = (int)((q / N) % N), t = (int)(q / ((long long)N * N));
  const double cx = cur[((size_t)t * N + i) * 2], cy = cur[((size_t)t * N + i) * 2 + 1];
  float o[11];
  normed_vec_mag64(f64sub(cur[((size_t)t * N + j) * 2], cx), f64sub(cur[((size_t)t * N + j) * 2 + 1], cy), o + 0);
  normed_vec_mag64(f64sub(goals[2 * j], cx), f64sub(goals[2 * j + 1], cy), o + 3);
  normed_vec_mag64(f64sub(prev[((size_t)t * N + j) * 2], cx), f64sub(prev[((size_t)t * N + j) * 2 + 1], cy), o + 6);
  o[9] = (float)rads[j];
  o[10] = (float)speeds[j];
  float* dst = info + (size_t)q * 11;
#pragma unroll
  for (int c = 0; c < 11; ++c) dst[c] = o[c];
  if (self_info != nullptr && i == j) {
    float* s = self_info + ((size_t)t * N + i) * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) s[c] = o[3 + c];
  }
}

int launch_others_info(const double* d_cur, const double* d_goals, const double* d_prev, const double* d_rads,
                       const double* d_speeds, int T, int N, float* d_info, float* d_self, cudaStream_t st) {
  const long long total = (long long)T * N * N;
  if (total <= 0) return 0;
  others_info_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_cur, d_goals, d_prev, d_rads, d_speeds, T, N, d_info, d_self);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
