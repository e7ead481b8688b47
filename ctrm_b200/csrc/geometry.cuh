// Obstacle loops built on sweep_static: StaticObjects.collide_continuous_sphere (static_objects.py:108-138).
#pragma once
#include "common.cuh"

// Conservative reject that can never change a verdict: the closest-approach point lies on the segment
// (t* in [0,1]), so if the obstacle centre is outside the segment's bounding box grown by rad_sum + 1e-9 the
// exact distance exceeds rad_sum by >= 1e-9, ten million times the fp64 rounding error of the test below.
__device__ __forceinline__ bool sweep_far(double fx, double fy, double tx, double ty, double ox, double oy, double rad_sum) {
  double R = rad_sum + 1e-9;
  return ox < fmin(fx, tx) - R || ox > fmax(fx, tx) + R || oy < fmin(fy, ty) - R || oy > fmax(fy, ty) + R;
}

// any over obstacles [o0, o1) of the instance (the reference builds the full list then `any`; the result is the
// same boolean).  One thread walks all obstacles.
__device__ __forceinline__ bool any_obstacle_hit(double fx, double fy, double tx, double ty, double rad,
                                                 const double* __restrict__ obs_xyr, int o0, int o1) {
  bool hit = false;
  // the segment's bounding box once; no early exit, so the obstacle loads of the following iterations are already in
  // flight while one is tested
  const double lox = fmin(fx, tx), hix = fmax(fx, tx), loy = fmin(fy, ty), hiy = fmax(fy, ty);
#pragma unroll 2
  for (int o = o0; o < o1; ++o) {
    const double ox = __ldg(&obs_xyr[3 * o + 0]), oy = __ldg(&obs_xyr[3 * o + 1]), orad = __ldg(&obs_xyr[3 * o + 2]);
    const double rs = f64add(rad, orad);
    const double R = rs + 1e-9;  // sweep_far, with the box hoisted
    if (ox < lox - R || ox > hix + R || oy < loy - R || oy > hiy + R) continue;
    hit = hit || sweep_static(fx, fy, tx, ty, ox, oy, rs);
  }
  return hit;
}

// The same `any over obstacles`, with the far obstacles rejected in fp32 from a float copy (x, y, rad, 0) of each obstacle:
// one 16-byte load and eight fp32 operations per far obstacle instead of three 8-byte loads and ten fp64 operations.
// Conservative: the box is grown by rad_sum + 1e-5 while the float copies and the float box are each within 1.2e-7 of the
// doubles for coordinates in the unit square (|coordinate| <= 16 keeps the total rounding below 1e-5), so a rejected obstacle
// is farther than rad_sum from every point of the segment and the exact test could only answer "no hit".
__device__ __forceinline__ bool any_obstacle_hit_f4(double fx, double fy, double tx, double ty, double rad,
                                                    const double* __restrict__ obs_xyr, const float4* __restrict__ obs_f4,
                                                    int o0, int o1) {
  bool hit = false;
  const float lox = (float)fmin(fx, tx), hix = (float)fmax(fx, tx), loy = (float)fmin(fy, ty), hiy = (float)fmax(fy, ty);
  const float radf = (float)rad + 1e-5f;
  for (int o = o0; o < o1; o += 4) {
    float4 q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = __ldg(&obs_f4[min(o + j, o1 - 1)]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (o + j >= o1) break;
      const float R = radf + q[j].z;
      if (q[j].x < lox - R || q[j].x > hix + R || q[j].y < loy - R || q[j].y > hiy + R) continue;
      const double* g = obs_xyr + 3 * (size_t)(o + j);
      hit = hit || sweep_static(fx, fy, tx, ty, __ldg(g), __ldg(g + 1), f64add(rad, __ldg(g + 2)));
    }
  }
  return hit;
}
