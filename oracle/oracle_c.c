/* ORACLE (test infrastructure) — plain C restatement of the integer / fp64 pieces of the reference's timed-roadmap
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this; the product never does.
 * Build: make -C oracle oracle_c   (gcc -O2 -ffp-contract=off: every fp64 product is rounded before the add, like the
 * reference's generic x86-64 build).  Each function cites the reference file:line it follows (paths relative to the
 * reference checkout). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cost_to_go.cpp:41-46, :121-126 — v < 1 ? floor(v * M) : M - 1, kept in a uint16 */
static int oc_cell(double v, int M) {
  int c = (v < 1.0) ? (int)floor(v * (double)M) : (M - 1);
  return c & 0xFFFF;
}

/* collision_check.cpp:16-58 with from2 == to2 == (ox, oy) (static_objects.py:126-133) */
static int oc_sweep(double fx, double fy, double tx, double ty, double r1, double ox, double oy, double r2) {
  double f[2] = {fx, fy}, t[2] = {tx, ty}, o[2] = {ox, oy};
  double a = 0, b = 0;
  for (int i = 0; i < 2; ++i) {
    double val = -f[i] + t[i] + o[i] - o[i];
    a += val * val;
    b += val * (f[i] - o[i]);
  }
  double tt = (b >= 0) ? 0.0 : ((a + b <= 0) ? 1.0 : -b / a);
  double tot = 0;
  for (int i = 0; i < 2; ++i) {
    double val = ((1 - tt) * f[i] + tt * t[i]) - ((1 - tt) * o[i] + tt * o[i]);
    tot += val * val;
  }
  double rad = r1 + r2;
  return tot <= rad * rad;
}

/* StaticObjects.collide_continuous_sphere (static_objects.py:108-138): any over the obstacles, n segments */
void oc_collide_segments(int64_t n, const double* from, const double* to, const double* rad, int n_obs,
                         const double* obs_xyr, uint8_t* out) {
  for (int64_t s = 0; s < n; ++s) {
    int hit = 0;
    for (int o = 0; o < n_obs; ++o)
      hit |= oc_sweep(from[2 * s], from[2 * s + 1], to[2 * s], to[2 * s + 1], rad[s], obs_xyr[3 * o], obs_xyr[3 * o + 1],
                      obs_xyr[3 * o + 2]);
    out[s] = (uint8_t)hit;
  }
}

/* ObstacleSphere.draw_2d (obstacle.py:63-81) + StaticObjects.draw_2d (static_objects.py:173-187); skimage.draw.disk
 * 0.18.1: bounding box ceil(c - R)..floor(c + R) clipped, ((r - r0)/R)^2 + ((c - c0)/R)^2 < 1 in float64 */
void oc_occupancy(int M, int n_obs, const double* obs_xyr, uint8_t* occ) {
  memset(occ, 0, (size_t)M * M);
  for (int o = 0; o < n_obs; ++o) {
    int X = (int)(M * obs_xyr[3 * o]), Y = (int)(M * obs_xyr[3 * o + 1]), R = (int)(M * obs_xyr[3 * o + 2]);
    double Rd = (double)R;
    for (int x = X - R; x <= X + R; ++x)
      for (int y = Y - R; y <= Y + R; ++y) {
        if (x < 0 || x >= M || y < 0 || y >= M) continue;
        double dx = (double)(x - X) / Rd, dy = (double)(y - Y) / Rd; /* R == 0 -> nan/inf -> comparison false */
        if (dx * dx + dy * dy < 1.0) occ[x * M + y] = 1;
      }
  }
}

/* getCostToGo2D (cost_to_go.cpp:25-107): FIFO breadth-first search from the goal cell; a neighbour is entered when it is
 * inside, unvisited, free, and no footprint cell (stencil == 1) is outside the map or occupied.  stencil is (2r+1)^2. */
void oc_cost_to_go(int M, const uint8_t* occ, const int32_t* stencil, int r, const double* goal, uint16_t* ctg) {
  const int DX[4] = {-1, 1, 0, 0}, DY[4] = {0, 0, -1, 1};
  for (int i = 0; i < M * M; ++i) ctg[i] = 65535;
  int gx = oc_cell(goal[0], M), gy = oc_cell(goal[1], M);
  if (gx >= M || gy >= M) return;
  int* queue = (int*)malloc(sizeof(int) * (size_t)M * M + 4);
  int head = 0, tail = 0;
  ctg[gx * M + gy] = 0;
  queue[tail++] = gx * M + gy;
  while (head < tail) {
    int cur = queue[head++], x = cur / M, y = cur % M;
    for (int m = 0; m < 4; ++m) {
      int nx = x + DX[m], ny = y + DY[m];
      if (nx < 0 || nx >= M || ny < 0 || ny >= M) continue;
      if (ctg[nx * M + ny] != 65535) continue;
      if (occ[nx * M + ny] == 1) continue;
      int blocked = 0;
      for (int k = -r; k <= r && !blocked; ++k)
        for (int l = -r; l <= r; ++l) {
          if (stencil[(k + r) * (2 * r + 1) + (l + r)] != 1) continue;
          int xk = nx + k, yl = ny + l;
          if (xk < 0 || xk >= M || yl < 0 || yl >= M || occ[xk * M + yl] != 0) { blocked = 1; break; }
        }
      if (blocked) continue;
      ctg[nx * M + ny] = (uint16_t)(ctg[cur] + 1);
      queue[tail++] = nx * M + ny;
    }
  }
  free(queue);
}

/* getFov2dOccupancyCost (cost_to_go.cpp:109-186), flatten layout; ctg 65535 = unreachable (the reference sees INT_MIN) */
void oc_fov_crop(int M, int F, const uint8_t* occ, const uint16_t* ctg, const double* pos, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * 2 * (size_t)F * F);
  int cx = oc_cell(pos[0], M), cy = oc_cell(pos[1], M), b = (F - 1) / 2;
  if (cx >= M || cy >= M) return;
  int c = ctg[cx * M + cy];
  for (int x = (cx - b > 0 ? cx - b : 0); x <= (cx + b < M - 1 ? cx + b : M - 1); ++x)
    for (int y = (cy - b > 0 ? cy - b : 0); y <= (cy + b < M - 1 ? cy + b : M - 1); ++y) {
      int i = (x - cx + b) * F + (y - cy + b), _c = ctg[x * M + y];
      out[i] = 1 - occ[x * M + y];
      out[F * F + i] = (_c != 65535) && (c == 65535 || _c < c);
    }
}
