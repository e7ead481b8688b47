// Occupancy raster, footprint passability masks, batched cost-to-go wavefront and FOV raster.
//   reference: StaticObjects.draw_2d (environment/static_objects.py:173-187), ObstacleSphere.draw_2d
//   (environment/obstacle.py:63-81), get_approx_cost_to_go_matrix_2d (learning/fov.py:34-63),
//   getCostToGo2D (cpp_helper/cost_to_go_wrapper/src/cost_to_go.cpp:25-107),
//   getFov2dOccupancyCost (cost_to_go.cpp:109-186).
// All integer/bit work: results are bit-exact by construction.
#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// obstacle -> integer disk (X, Y, R):  X = int(M*x), Y = int(M*y), R = int(M*rad)   (obstacle.py:73-77)
// ------------------------------------------------------------------------------------------------
__global__ void obs_to_cells_kernel(const double* __restrict__ obs_xyr, int O, int M, int4* __restrict__ cells) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  double m = (double)M;
  int X = (int)f64mul(m, obs_xyr[3 * o + 0]);  // int() truncates toward zero
  int Y = (int)f64mul(m, obs_xyr[3 * o + 1]);
  int R = (int)f64mul(m, obs_xyr[3 * o + 2]);
  cells[o] = make_int4(X, Y, R * R, R);
}

// occ[b][x][y] = any_o ( (x-X)^2 + (y-Y)^2 < R^2 ) — the integer form of skimage.draw.disk's float test
// ((r-r0)/R)^2 + ((c-c0)/R)^2 < 1 (exhaustively identical for R <= 40, SURVEY §8a note (i); R = 0 is empty).
// The same raster as O(output) + O(disk areas): the maps are cleared with one memset at the HBM write rate and every
// obstacle then writes the cells of ITS disk (a few hundred bytes), instead of every 16-byte chunk of every map testing every
// obstacle of its instance (ncu: that kernel was ALU-bound, 330 instructions per chunk, 12 % of the DRAM rate).  occ = max
// over obstacles, and every writer writes 1, so the order does not matter.  One warp per obstacle; lanes = the y cells of a row.
__global__ void __launch_bounds__(256) raster_disks_kernel(const int4* __restrict__ cells, const int32_t* __restrict__ obs_off,
                                                           int B, int O, int M, uint8_t* __restrict__ occ) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= O) return;
  int lo = 0, hi = B - 1;  // instance of obstacle o: last b with obs_off[b] <= o
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (obs_off[mid] <= o) lo = mid; else hi = mid - 1;
  }
  const int4 d = __ldg(&cells[o]);  // (X, Y, R^2, R)
  if (d.w <= 0) return;             // R = 0: empty disk
  uint8_t* out = occ + (size_t)lo * M * M;
  const int x0 = max(0, d.x - d.w + 1), x1 = min(M - 1, d.x + d.w - 1);
  const int y0 = max(0, d.y - d.w + 1), y1 = min(M - 1, d.y + d.w - 1);
  for (int x = x0; x <= x1; ++x) {
    const int dx = x - d.x, rem2 = d.z - dx * dx;  // dy^2 < rem2 <=> inside
    for (int y = y0 + lane; y <= y1; y += 32) {
      const int dy = y - d.y;
      if (dy * dy < rem2) out[x * M + y] = 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// passability masks: bit (x, y) of mask p set iff the agent footprint centred on (x,y) is collision-free
// (cost_to_go.cpp:68-97): occ[x][y]==0 and no stencil cell with agent==1 is outside the map or occupied.
// Bit-parallel: the instance's "blocked" grid (occupied, or outside the map) is packed one bit per cell; the footprint test
// is then a dilation — for every set stencil cell (k, l) OR the blocked row x + k shifted by l — 32 cells per word operation
// instead of one byte load and compare per (cell, stencil cell) (ncu: the per-cell kernel was ALU-bound at 1 ms per 1,024
// masks).  MW = ceil(M/32).  stencils: [n_st][st_ld*st_ld] uint8, the (2r+1)^2 footprint with leading dimension st_ld.
// ------------------------------------------------------------------------------------------------
// blocked[b][x][w]: bit y of word w = occ[b][x][32 w + y] != 0; bits at y >= M are set (outside the map counts as blocked)
__global__ void __launch_bounds__(256) pack_blocked_kernel(const uint8_t* __restrict__ occ, int B, int M, int MW,
                                                           uint32_t* __restrict__ blocked) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * M * MW) return;
  const int w = (int)(warp % MW);
  const long long t = warp / MW;
  const int x = (int)(t % M), b = (int)(t / M);
  const int y = w * 32 + lane;
  const bool blk = (y >= M) || (occ[((size_t)b * M + x) * M + y] != 0);
  const uint32_t bits = __ballot_sync(0xFFFFFFFFu, blk);
  if (lane == 0) blocked[((size_t)b * M + x) * MW + w] = bits;
}

// one thread per (mask p, row x, word w)
__global__ void __launch_bounds__(256) passable_mask_kernel(const uint32_t* __restrict__ blocked, int M, int MW,
                                                            const uint8_t* __restrict__ stencils,
                                                            const int32_t* __restrict__ stencil_r, int st_ld,
                                                            const int32_t* __restrict__ pm_inst,
                                                            const int32_t* __restrict__ pm_stencil, int P,
                                                            uint32_t* __restrict__ pmask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * M * MW) return;
  const int w = (int)(i % MW);
  const long long t = i / MW;
  const int x = (int)(t % M), p = (int)(t / M);
  const uint32_t* Bk = blocked + (size_t)pm_inst[p] * M * MW;
  const int st = pm_stencil[p], r = stencil_r[st];
  const uint8_t* s = stencils + (size_t)st * st_ld * st_ld;
  uint32_t bad = Bk[(size_t)x * MW + w];  // the cell itself must be free (cost_to_go.cpp:68)
  for (int k = -r; k <= r; ++k) {
    const int xk = x + k;
    const bool row_in = xk >= 0 && xk < M;
    const uint32_t* row = Bk + (size_t)(row_in ? xk : 0) * MW;
    for (int l = -r; l <= r; ++l) {
      if (s[(k + r) * st_ld + (l + r)] == 0) continue;
      // bit y of the result reads blocked(xk, 32 w + y + l): a 32-bit window of the row starting at bit 32 w + l; words
      // outside the row (and the whole row when xk is outside the map) read as all ones
      const int pos = 32 * w + l;
      const int q = pos >> 5, sft = pos & 31;  // arithmetic shift = floor for negative positions
      const uint32_t w0 = (row_in && q >= 0 && q < MW) ? row[q] : 0xFFFFFFFFu;
      const uint32_t w1 = (row_in && q + 1 >= 0 && q + 1 < MW) ? row[q + 1] : 0xFFFFFFFFu;
      bad |= sft ? (w0 >> sft) | (w1 << (32 - sft)) : w0;
    }
  }
  pmask[((size_t)p * M + x) * MW + w] = ~bad;
}

// ------------------------------------------------------------------------------------------------
// cost-to-go wavefront: one warp per goal, the whole M x M bit grid lives in registers.
// Lane l owns rows [l*RPL, (l+1)*RPL), each row MW words.  Per level:
//   new = (N|S|E|W neighbours of the frontier) & passable & ~visited ; ctg[new cells] = level.
// BFS hop counts are independent of the expansion order, so this equals the reference's std::queue BFS.
// The goal cell gets 0 unconditionally (cost_to_go.cpp:50) and is expanded even if not passable itself.
// ------------------------------------------------------------------------------------------------
template <int RPL, int MW>
__global__ void __launch_bounds__(128) wavefront_kernel(const uint32_t* __restrict__ pmask,
                                                        const int32_t* __restrict__ agent_pm,
                                                        const double* __restrict__ goals, int A, int M,
                                                        int layout, uint16_t* __restrict__ ctg) {
  int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (a >= A) return;
  const int mw = (M + 31) >> 5;  // words per row of the mask as written by passable_mask_kernel
  const uint32_t* pm = pmask + (size_t)agent_pm[a] * M * mw;
  uint16_t* out = ctg + (size_t)a * ctg_field_elems(M, layout);
  // open[r][w] = passable and not yet visited (one bit grid instead of two: the kernel is bound by registers -> occupancy)
  uint32_t open[RPL][MW], fr[RPL][MW];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    int x = lane * RPL + r;
#pragma unroll
    for (int w = 0; w < MW; ++w) {
      open[r][w] = (x < M && w < mw) ? __ldg(&pm[x * mw + w]) : 0u;
      fr[r][w] = 0u;
    }
  }
  int gx = grid_cell(goals[2 * a + 0], M), gy = grid_cell(goals[2 * a + 1], M);
  if (gx >= M || gy >= M) return;  // asserts are compiled out in the reference; field stays unreachable
  // selects, not an indexed store: a run-time index would push both bit grids into local memory for the whole kernel
  const int g_r = gx - lane * RPL, g_w = gy >> 5;
  const uint32_t g_bit = 1u << (gy & 31);
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
#pragma unroll
    for (int w = 0; w < MW; ++w) {
      const uint32_t m = (g_r == r && g_w == w) ? g_bit : 0u;
      fr[r][w] = m;
      open[r][w] &= ~m;
    }
  }
  if (lane == gx / RPL) out[ctg_index(gx, gy, M, layout)] = 0;
  for (int level = 1; level < 65535; ++level) {
    uint32_t nw[RPL][MW];
    uint32_t any = 0u;
    // rows across lane boundaries
    uint32_t up_in[MW], dn_in[MW];
#pragma unroll
    for (int w = 0; w < MW; ++w) {
      up_in[w] = __shfl_up_sync(0xFFFFFFFFu, fr[RPL - 1][w], 1);  // row x-1 of my first row
      dn_in[w] = __shfl_down_sync(0xFFFFFFFFu, fr[0][w], 1);      // row x+1 of my last row
      if (lane == 0) up_in[w] = 0u;
      if (lane == 31) dn_in[w] = 0u;
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
#pragma unroll
      for (int w = 0; w < MW; ++w) {
        uint32_t f = fr[r][w];
        uint32_t lo = (w > 0) ? fr[r][w - 1] : 0u;
        uint32_t hi = (w < MW - 1) ? fr[r][w + 1] : 0u;
        uint32_t nb = (f << 1) | (lo >> 31) | (f >> 1) | (hi << 31);
        nb |= (r > 0) ? fr[r - 1][w] : up_in[w];
        nb |= (r < RPL - 1) ? fr[r + 1][w] : dn_in[w];
        uint32_t n = nb & open[r][w];
        nw[r][w] = n;
        any |= n;
      }
    }
    if (!__any_sync(0xFFFFFFFFu, any != 0u)) break;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
#pragma unroll
      for (int w = 0; w < MW; ++w) {
        uint32_t n = nw[r][w];
        open[r][w] &= ~n;
        fr[r][w] = n;
        while (n) {
          int bit = __ffs(n) - 1;
          n &= n - 1;
          out[ctg_index(lane * RPL + r, w * 32 + bit, M, layout)] = (uint16_t)level;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FOV raster (cost_to_go.cpp:109-186, flatten layout) for all agents: fov[a][x_fov*F + y_fov] = 1 - occ,
// fov[a][F^2 + ...] = reach(x,y) && (!reach(c) || ctg[x,y] < ctg[c]); out-of-map cells stay 0; float32.
// One warp rasterises one agent's 2F^2 values into shared memory; the block (APB agents) then streams its
// contiguous APB*2F^2 floats out with coalesced 16-byte stores (APB even keeps every block base 16B aligned).
// ------------------------------------------------------------------------------------------------
template <int APB>
__global__ void __launch_bounds__(APB * 32) fov_raster_kernel(const uint8_t* __restrict__ occ,
                                                              const uint16_t* __restrict__ ctg,
                                                              const int32_t* __restrict__ agent_inst,
                                                              const double* __restrict__ pos,
                                                              const uint8_t* __restrict__ skip_inst, int A, int M,
                                                              int F, int layout, float* __restrict__ fov) {
  extern __shared__ __align__(16) float tile[];  // [APB][2F^2]
  const int FF = F * F, ROW = 2 * FF;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int a0 = blockIdx.x * APB;
  int a = a0 + warp;
  bool live = false;
  if (a < A) {
    int b = agent_inst[a];
    live = (skip_inst == nullptr) || (skip_inst[b] == 0);
    float* t = tile + warp * ROW;
    if (live) {
      const uint8_t* o = occ + (size_t)b * M * M;
      const uint16_t* c = ctg + (size_t)a * ctg_field_elems(M, layout);
      int cx = grid_cell(pos[2 * a + 0], M), cy = grid_cell(pos[2 * a + 1], M);
      int buf = (F - 1) / 2;
      // a position outside [0,1] wraps through uint16 in the reference; such crops are entirely outside the map
      bool centre_in = cx < M && cy < M;
      int cc = centre_in ? (int)c[ctg_index(cx, cy, M, layout)] : CTRM_UNREACH;
      bool c_reach = cc != CTRM_UNREACH;
      if ((M & 3) == 0 && centre_in) {
        // the crop as (row xf, aligned group of four y cells): one 4-byte occupancy load and one 8-byte load of the cost
        // field per work item instead of eight scattered 1-2 byte loads (the kernel was bound by its load-request rate)
        for (int i = lane; i < ROW; i += 32) t[i] = 0.f;
        __syncwarp();
        const int ylo = cy - buf;
        const int g0 = ylo >> 2;                    // floor(ylo / 4), also for negative ylo
        const int NG = ((cy + buf) >> 2) - g0 + 1;  // groups per crop row
        for (int it = lane; it < F * NG; it += 32) {
          const int xf = it / NG, gy = it - xf * NG;
          const int x = cx - buf + xf, yb = (g0 + gy) << 2;
          if (x < 0 || x >= M || yb < 0 || yb >= M) continue;  // yb is a multiple of 4 and so is M: all four cells inside or none
          const uint32_t ow = __ldg(reinterpret_cast<const uint32_t*>(&o[x * M + yb]));
          uint32_t cv[4];
          // both layouts keep 4 aligned y cells of a row contiguous and 8-byte aligned (row-major: M % 4 == 0)
          const uint2 cw = __ldg(reinterpret_cast<const uint2*>(&c[ctg_index(x, yb, M, layout)]));
          cv[0] = cw.x & 0xFFFFu; cv[1] = cw.x >> 16; cv[2] = cw.y & 0xFFFFu; cv[3] = cw.y >> 16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int yf = yb - ylo + j;
            if (yf < 0 || yf >= F) continue;
            const int _c = (int)cv[j];
            t[xf * F + yf] = ((ow >> (8 * j)) & 0xFFu) ? 0.f : 1.f;
            t[FF + xf * F + yf] = (_c != CTRM_UNREACH && (!c_reach || _c < cc)) ? 1.f : 0.f;
          }
        }
      } else
      for (int i = lane; i < FF; i += 32) {
        int xf = i / F, yf = i - xf * F;
        int x = cx - buf + xf, y = cy - buf + yf;
        float v0 = 0.f, v1 = 0.f;
        if (centre_in && x >= 0 && x < M && y >= 0 && y < M) {
          v0 = (float)(1 - (int)o[x * M + y]);
          int _c = (int)c[ctg_index(x, y, M, layout)];
          v1 = (_c != CTRM_UNREACH && (!c_reach || _c < cc)) ? 1.f : 0.f;
        }
        t[i] = v0;
        t[FF + i] = v1;
      }
    } else {
      for (int i = lane; i < ROW; i += 32) t[i] = 0.f;  // finished instance: row is never consumed
    }
  }
  // a block whose agents all belong to finished instances writes nothing
  if (!__syncthreads_or(live)) return;
  int n_rows = min(APB, A - a0);
  size_t base = (size_t)a0 * ROW;  // floats; APB even -> base*4 bytes is a multiple of 16
  int n_f = n_rows * ROW;
  int n_v = n_f >> 2;
  float4* dst = reinterpret_cast<float4*>(fov + base);
  const float4* src = reinterpret_cast<const float4*>(tile);
  for (int v = threadIdx.x; v < n_v; v += blockDim.x) dst[v] = src[v];
  for (int i = n_v * 4 + threadIdx.x; i < n_f; i += blockDim.x) fov[base + i] = tile[i];
}

// ------------------------------------------------------------------------------------------------
// FOV raster for the tensor-core encoders (fov_tc.cu): the same crop, one BIT per input, stored as
// [tile of 128 rows][chunk of 96 inputs][row][12 bytes].  A warp rasterises its agent's crop into a byte row in shared
// memory (cells outside the map stay 0), packs it and stores 24 words; fov_encode expands a chunk to the canonical 8-bit UMMA
// tile inside ITS shared memory, so the FOV tensor crosses HBM as 96 bytes per agent-step in each direction instead of 768.
// ------------------------------------------------------------------------------------------------
#define RT_APB 8
#define RT_CHUNK 96
template <int FT>  // FT > 0: compile-time fov_size (constant divisions, unrolled loads); 0: run-time F
__global__ void __launch_bounds__(RT_APB * 32) fov_raster_tiles_kernel(const uint8_t* __restrict__ occ,
                                                                      const uint16_t* __restrict__ ctg,
                                                                      const int32_t* __restrict__ agent_inst,
                                                                      const double* __restrict__ pos,
                                                                      const uint8_t* __restrict__ skip_inst, int A, int M, int F_rt,
                                                                      int n_chunks, uint8_t* __restrict__ tiles,
                                                                      const int32_t* __restrict__ live_agent,
                                                                      const int32_t* __restrict__ n_live) {
  extern __shared__ __align__(16) uint8_t rt_smem[];
  const int F = FT > 0 ? FT : F_rt;
  const int FF = F * F, KD = 2 * FF, Kpad = n_chunks * RT_CHUNK;
  const int row_bytes = Kpad + 16;  // one byte per input; +16: rows land in different banks for the 16-byte read-out
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows: agents of the unfinished instances (live_agent, rebuilt between graph chunks) or all agents
  const int n_rows = live_agent != nullptr ? *n_live : A;
  const int a0 = blockIdx.x * RT_APB;  // first ROW of this block
  if (a0 >= n_rows) return;
  const int row = a0 + warp;
  const int a = row < n_rows ? (live_agent != nullptr ? live_agent[row] : row) : A;
  uint8_t* t = rt_smem + (size_t)warp * row_bytes;
  bool live = false;
  if (a < A) {
    const int b = agent_inst[a];
    live = (skip_inst == nullptr) || (skip_inst[b] == 0);
    if (live) {
      const uint8_t* o = occ + (size_t)b * M * M;
      const uint16_t* c = ctg + (size_t)a * ctg_field_elems(M, 1);
      const int cx = grid_cell(pos[2 * a + 0], M), cy = grid_cell(pos[2 * a + 1], M);
      const int buf = (F - 1) / 2;
      const bool centre_in = cx < M && cy < M;
      const int cc = centre_in ? (int)c[ctg_index(cx, cy, M, 1)] : CTRM_UNREACH;
      const bool c_reach = cc != CTRM_UNREACH;
      // zero the row (cells outside the map stay 0 in both channels), then visit the crop as (row xf, aligned group of four
      // y cells): one 8-byte load of the micro-tiled cost field and one 4-byte load of the occupancy row per work item
      for (int i = lane; i < Kpad / 16; i += 32) reinterpret_cast<uint4*>(t)[i] = make_uint4(0u, 0u, 0u, 0u);
      __syncwarp();
      if (centre_in) {
        const int ylo = cy - buf;
        const int g0 = ylo >> 2;                          // floor(ylo / 4), also for negative ylo
        const int NG = ((cy + buf) >> 2) - g0 + 1;        // groups per crop row
        const unsigned inv = 65536u / (unsigned)NG + 1u;  // it / NG == (it * inv) >> 16 for it < F * NG <= 1071
        const bool occ_vec = (M & 3) == 0;
        for (int it = lane; it < F * NG; it += 32) {
          const int xf = (int)(((unsigned)it * inv) >> 16), gy = it - xf * NG;
          const int x = cx - buf + xf, yb = (g0 + gy) << 2;
          if (x < 0 || x >= M || yb < 0 || yb >= M) continue;  // yb is a multiple of 4: negative groups are entirely outside
          const uint2 cw = __ldg(reinterpret_cast<const uint2*>(&c[ctg_index(x, yb, M, 1)]));
          uint32_t ow;
          if (occ_vec) {
            ow = __ldg(reinterpret_cast<const uint32_t*>(&o[x * M + yb]));
          } else {
            ow = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (yb + j < M) ow |= (uint32_t)__ldg(&o[x * M + yb + j]) << (8 * j);
          }
          const uint32_t cv[4] = {cw.x & 0xFFFFu, cw.x >> 16, cw.y & 0xFFFFu, cw.y >> 16};
          uint8_t* t0 = t + xf * F + (yb - ylo);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int yf = yb - ylo + j;
            if (yf < 0 || yf >= F || yb + j >= M) continue;
            const int _c = (int)cv[j];
            t0[j] = ((ow >> (8 * j)) & 0xFFu) ? 0 : 1;
            t0[FF + j] = (_c != CTRM_UNREACH && (!c_reach || _c < cc)) ? 1 : 0;
          }
        }
      }
    }
  }
  // ---- pack the row to BITS and store it: [tile of 128 rows][chunk of 96 inputs][row][12 bytes].  fov_encode expands a
  //      chunk back to the 8-bit UMMA operand layout inside shared memory (fov_tc.cu), so the crop leaves the SM as 96 bytes
  //      per agent-step instead of 768 and is read back as 96.  Word w of the row = inputs 32 w .. 32 w + 31, bit = input.
  if (live) {
    __syncwarp();
    const int nw = Kpad >> 5;  // 24 words for the trained F = 19 (n_chunks x 96 inputs)
    for (int w = lane; w < nw; w += 32) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(t + 32 * w);
      uint32_t bits = 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) bits |= (((src[q] * 0x01020408u) >> 24) & 0xFu) << (4 * q);  // four 0/1 bytes -> a nibble
      const int c = w / 3, j = w - 3 * c;
      const int r = row & 127;
      __stcs(reinterpret_cast<uint32_t*>(tiles + ((size_t)((row >> 7) * n_chunks + c) * 128 + r) * 12 + 4 * j), bits);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_obs_to_cells(const double* d_obs_xyr, int O, int M, int4* d_cells, cudaStream_t st) {
  if (O <= 0) return 0;
  obs_to_cells_kernel<<<(O + 127) / 128, 128, 0, st>>>(d_obs_xyr, O, M, d_cells);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_raster_occupancy(const int4* d_cells, const int32_t* d_obs_off, int B, int O, int M, uint8_t* d_occ, cudaStream_t st) {
  if (B <= 0) return 0;
  CTRM_CUDA_OK(cudaMemsetAsync(d_occ, 0, (size_t)B * M * M, st));
  if (O > 0) {
    raster_disks_kernel<<<(unsigned)((O + 7) / 8), 256, 0, st>>>(d_cells, d_obs_off, B, O, M, d_occ);
    CTRM_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int launch_passable_masks(const uint8_t* d_occ, int B, int M, const uint8_t* d_stencils, const int32_t* d_stencil_r, int st_ld,
                          const int32_t* d_pm_inst, const int32_t* d_pm_stencil, int P, uint32_t* d_blocked, uint32_t* d_pmask,
                          cudaStream_t st) {
  const int MW = (M + 31) / 32;
  if (P <= 0 || B <= 0) return 0;
  const long long pw = (long long)B * M * MW * 32;
  pack_blocked_kernel<<<(unsigned)((pw + 255) / 256), 256, 0, st>>>(d_occ, B, M, MW, d_blocked);
  CTRM_CUDA_OK(cudaGetLastError());
  const long long threads = (long long)P * M * MW;
  passable_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_blocked, M, MW, d_stencils, d_stencil_r, st_ld,
                                                                       d_pm_inst, d_pm_stencil, P, d_pmask);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_wavefront(const uint32_t* d_pmask, const int32_t* d_agent_pm, const double* d_goals, int A, int M, int layout,
                     uint16_t* d_ctg, cudaStream_t st) {
  if (A <= 0) return 0;
  CTRM_CUDA_OK(cudaMemsetAsync(d_ctg, 0xFF, (size_t)A * ctg_field_elems(M, layout) * sizeof(uint16_t), st));
  int blocks = (A + 3) / 4;
  if (M <= 160) wavefront_kernel<5, 5><<<blocks, 128, 0, st>>>(d_pmask, d_agent_pm, d_goals, A, M, layout, d_ctg);
  else if (M <= 256) wavefront_kernel<8, 8><<<blocks, 128, 0, st>>>(d_pmask, d_agent_pm, d_goals, A, M, layout, d_ctg);
  else return ctrm_fail(-1, "map_size > 255 unsupported (reference limit, cost_to_go.cpp:47)");
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int APB>
static int launch_fov_raster_t(const uint8_t* d_occ, const uint16_t* d_ctg, const int32_t* d_agent_inst, const double* d_pos,
                               const uint8_t* d_skip_inst, int A, int M, int F, int layout, float* d_fov, cudaStream_t st) {
  size_t smem = (size_t)APB * 2 * F * F * sizeof(float);
  if (smem > 48 * 1024)
    CTRM_CUDA_OK(cudaFuncSetAttribute(fov_raster_kernel<APB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fov_raster_kernel<APB><<<(A + APB - 1) / APB, APB * 32, smem, st>>>(d_occ, d_ctg, d_agent_inst, d_pos, d_skip_inst, A, M, F,
                                                                   layout, d_fov);
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_fov_raster(const uint8_t* d_occ, const uint16_t* d_ctg, const int32_t* d_agent_inst, const double* d_pos,
                      const uint8_t* d_skip_inst, int A, int M, int F, int layout, float* d_fov, cudaStream_t st) {
  if (A <= 0) return 0;
  if (F <= 27) return launch_fov_raster_t<8>(d_occ, d_ctg, d_agent_inst, d_pos, d_skip_inst, A, M, F, layout, d_fov, st);
  return launch_fov_raster_t<2>(d_occ, d_ctg, d_agent_inst, d_pos, d_skip_inst, A, M, F, layout, d_fov, st);
}

int launch_fov_raster_tiles(const uint8_t* d_occ, const uint16_t* d_ctg, const int32_t* d_agent_inst, const double* d_pos,
                            const uint8_t* d_skip_inst, int A, int M, int F, int n_chunks, int layout, uint8_t* d_tiles,
                            const int32_t* d_live_agent, const int32_t* d_n_live, cudaStream_t st) {
  if (layout != 1) return ctrm_fail(-1, "fov_raster_tiles reads the micro-tiled cost-to-go layout only");
  if (A <= 0) return 0;
  const size_t smem = (size_t)RT_APB * ((size_t)n_chunks * RT_CHUNK + 16);
  const int blocks = (A + RT_APB - 1) / RT_APB;
  if (F == 19) {  // the trained configuration
    fov_raster_tiles_kernel<19><<<blocks, RT_APB * 32, smem, st>>>(d_occ, d_ctg, d_agent_inst, d_pos, d_skip_inst, A, M, F, n_chunks,
                                                                 d_tiles, d_live_agent, d_n_live);
  } else {
    static SmemAttrCache attr;
    if (smem > 48 * 1024) CTRM_CUDA_OK(attr.ensure(fov_raster_tiles_kernel<0>, smem));
    fov_raster_tiles_kernel<0><<<blocks, RT_APB * 32, smem, st>>>(d_occ, d_ctg, d_agent_inst, d_pos, d_skip_inst, A, M, F, n_chunks,
                                                                d_tiles, d_live_agent, d_n_live);
  }
  CTRM_CUDA_OK(cudaGetLastError());
  return 0;
}
