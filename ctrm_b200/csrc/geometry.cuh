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
