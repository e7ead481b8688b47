// Shared device helpers of the ctrm_b200 kernels: fp64 arithmetic that never contracts to FMA,
// the reference's swept-circle test, validity predicates, grid-cell mapping and counter-based RNG.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CTRM_UNREACH 65535

// ---- fp64 with one rounding per operation (the reference is generic x86-64 without FMA contraction;
//      SURVEY §7.2 item 3).  __dadd_rn/__dmul_rn are never fused by nvcc. ----
__device__ __forceinline__ double f64add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double f64sub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double f64mul(double a, double b) { return __dmul_rn(a, b); }

// cost_to_go.cpp:41-46 / :121-126: v < 1 ? floor(v*M) : M-1, stored in a uint16
__device__ __forceinline__ int grid_cell(double v, int M) {
  int c = (v < 1.0) ? (int)floor(f64mul(v, (double)M)) : (M - 1);
  return c & 0xFFFF;
}

// continuousCollideSpheres (collision_check.cpp:16-58) for a moving circle (f -> t, radius folded into rad_sum)
// against a STATIC circle at o (from2 == to2 == o, static_objects.py:126-133).  Operation order is the
// reference's: val = ((-f + t) + o) - o ; a += val*val ; b += val*(f - o) ; closest approach at t* ;
// dist^2 <= (rad1+rad2)^2.  rad_sum = rad1 + rad2 (one rounded add, collision_check.cpp:56).
__device__ __forceinline__ bool sweep_static(double fx, double fy, double tx, double ty, double ox, double oy,
                                             double rad_sum) {
  double vx = f64sub(f64add(f64add(-fx, tx), ox), ox);
  double vy = f64sub(f64add(f64add(-fy, ty), oy), oy);
  double dx = f64sub(fx, ox), dy = f64sub(fy, oy);
  double a = f64add(f64add(0.0, f64mul(vx, vx)), f64mul(vy, vy));
  double b = f64add(f64add(0.0, f64mul(vx, dx)), f64mul(vy, dy));
  double tt;
  if (b >= 0.0) tt = 0.0;
  else if (f64add(a, b) <= 0.0) tt = 1.0;
  else tt = (-b) / a;
  double omt = f64sub(1.0, tt);
  double px = f64sub(f64add(f64mul(omt, fx), f64mul(tt, tx)), f64add(f64mul(omt, ox), f64mul(tt, ox)));
  double py = f64sub(f64add(f64mul(omt, fy), f64mul(tt, ty)), f64add(f64mul(omt, oy), f64mul(tt, oy)));
  double tot = f64add(f64add(0.0, f64mul(px, px)), f64mul(py, py));
  return tot <= f64mul(rad_sum, rad_sum);
}

// the general two-moving-spheres form, dim 2 or 3 (collision_check.cpp:16-58 verbatim semantics)
__device__ __forceinline__ bool sweep_pair(const double* f1, const double* t1, double r1, const double* f2,
                                           const double* t2, double r2, int dim) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < dim; ++i) {
    double val = f64sub(f64add(f64add(-f1[i], t1[i]), f2[i]), t2[i]);
    a = f64add(a, f64mul(val, val));
    b = f64add(b, f64mul(val, f64sub(f1[i], f2[i])));
  }
  double tt;
  if (b >= 0.0) tt = 0.0;
  else if (f64add(a, b) <= 0.0) tt = 1.0;
  else tt = (-b) / a;
  double omt = f64sub(1.0, tt);
  double tot = 0.0;
  for (int i = 0; i < dim; ++i) {
    double val = f64sub(f64add(f64mul(omt, f1[i]), f64mul(tt, t1[i])), f64add(f64mul(omt, f2[i]), f64mul(tt, t2[i])));
    tot = f64add(tot, f64mul(val, val));
  }
  double rad = f64add(r1, r2);
  return tot <= f64mul(rad, rad);
}

// valid_move pre-checks (roadmap/utils.py:37-41): L-inf then exact L2, `sum(p**2)` = (0 + px^2) + py^2
__device__ __forceinline__ bool speed_ok(double ax, double ay, double bx, double by, double speed, double speed2) {
  double px = fabs(f64sub(ax, bx)), py = fabs(f64sub(ay, by));
  if (px > speed || py > speed) return false;
  double s = f64add(f64add(0.0, f64mul(px, px)), f64mul(py, py));
  return !(s > speed2);
}

// bounds test of the valid_edge closure (learned_sampler.py:128-129) on the destination
__device__ __forceinline__ bool bounds_ok(double bx, double by, double rad) {
  double mn = fmin(bx, by), mx = fmax(bx, by);
  return !(mn < rad || f64sub(1.0, rad) < mx);
}

// get_dist_arr (roadmap_learned/utils.py:16-18): sum((c - loc)**2) over 2 components
__device__ __forceinline__ double dist2(double cx, double cy, double lx, double ly) {
  double dx = f64sub(cx, lx), dy = f64sub(cy, ly);
  return f64add(f64add(0.0, f64mul(dx, dx)), f64mul(dy, dy));
}

// ---- cost-to-go field layouts ----
// layout 0: row-major [x][y] (what getCostToGo2D returns, cost_to_go.cpp:105-106).
// layout 1: 8 x 4 cell micro-tiles of 64 bytes (one DRAM burst): a 19 x 19 FOV crop then touches ~18 bursts (1.1 KB) instead
//           of 19 row segments that each straddle 2-3 sectors (measured 3.7 KB of DRAM reads per crop with layout 0).
__host__ __device__ __forceinline__ size_t ctg_field_elems(int M, int layout) {
  return layout ? (size_t)((M + 7) >> 3) * ((M + 3) >> 2) * 32 : (size_t)M * M;
}
__device__ __forceinline__ int ctg_index(int x, int y, int M, int layout) {
  return layout ? ((((x >> 3) * ((M + 3) >> 2) + (y >> 2)) << 5) + ((x & 7) << 2) + (y & 3)) : (x * M + y);
}

// ---- Philox4x32-10, counter layout of oracle/rng_tables.py ----
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
#define STREAM_GATE 0u
#define STREAM_RW 1u
#define STREAM_STEP 2u
#define STREAM_RANDIND 3u
#define STREAM_GUMBEL 8u

__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
    uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t rng_ctr1(int k_traj, int t) { return ((uint32_t)(k_traj & 0xFFFF) << 16) | (uint32_t)(t & 0xFFFF); }
__device__ __forceinline__ uint32_t rng_ctr3(uint32_t stream, uint32_t sub) { return ((stream & 0xFFu) << 8) | (sub & 0xFFu); }
__device__ __forceinline__ double u64_to_f64(uint32_t w0, uint32_t w1) {
  unsigned long long x = ((unsigned long long)w0 << 32) | (unsigned long long)w1;
  return (double)(x >> 11) * 1.1102230246251565e-16;  // 2^-53, exact
}
__device__ __forceinline__ float u32_to_f32(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }  // 2^-24

// (sin(2*pi*u), cos(2*pi*u)), bit-identical to oracle/rng_tables.det_sincos_2pi: exact quadrant reduction,
// one rounded multiply by pi/2, Taylor polynomials with individually rounded * and +.
__device__ __forceinline__ void det_sincos_2pi(double u, double* s_out, double* c_out) {
  const unsigned long long SB[9] = {0x3FF0000000000000ULL, 0xBFC5555555555555ULL, 0x3F81111111111111ULL,
                                    0xBF2A01A01A01A01AULL, 0x3EC71DE3A556C734ULL, 0xBE5AE64567F544E4ULL,
                                    0x3DE6124613A86D09ULL, 0xBD6AE7F3E733B81FULL, 0x3CE952C77030AD4AULL};
  const unsigned long long CB[10] = {0x3FF0000000000000ULL, 0xBFE0000000000000ULL, 0x3FA5555555555555ULL,
                                     0xBF56C16C16C16C17ULL, 0x3EFA01A01A01A01AULL, 0xBE927E4FB7789F5CULL,
                                     0x3E21EED8EFF8D898ULL, 0xBDA93974A8C07C9DULL, 0x3D2AE7F3E733B81FULL,
                                     0xBCA6827863B97D97ULL};
  double a = f64mul(u, 4.0);
  int q = (int)a;
  double f = f64sub(a, (double)q);
  bool swap = f > 0.5;
  double g = swap ? f64sub(1.0, f) : f;
  double x = f64mul(g, __longlong_as_double(0x3FF921FB54442D18LL));
  double x2 = f64mul(x, x);
  double s = __longlong_as_double((long long)SB[8]);
#pragma unroll
  for (int k = 7; k >= 0; --k) s = f64add(f64mul(s, x2), __longlong_as_double((long long)SB[k]));
  s = f64mul(s, x);
  double c = __longlong_as_double((long long)CB[9]);
#pragma unroll
  for (int k = 8; k >= 0; --k) c = f64add(f64mul(c, x2), __longlong_as_double((long long)CB[k]));
  if (swap) { double tmp = s; s = c; c = tmp; }
  if (q == 0) { *s_out = s; *c_out = c; }
  else if (q == 1) { *s_out = c; *c_out = -s; }
  else if (q == 2) { *s_out = -s; *c_out = -c; }
  else { *s_out = -c; *c_out = s; }
}

#define CTRM_CUDA_OK(call)                                   \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return ctrm_fail_cuda(_e, #call); \
  } while (0)

int ctrm_fail_cuda(cudaError_t e, const char* what);
int ctrm_fail(int code, const char* what);
