// Internal declarations shared by the .cu translation units of libctrm_b200 (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/ctrm_b200.h"

#define CTRM_MAX_W 4          // adjacency bitmask words per vertex: slots S = 32*W >= N_traj + 1  (N_traj <= 127)
#define CTRM_DIM_X 72         // [self_info 8 | fov_self 32 | comm 32]   (format_input.py:354-355)
#define CTRM_DIM_H 32
#define CTRM_DIM_INFO 11      // compiled.py:30-97

// device-side view of the weights
struct DevModel {
  const float* blob;
  ctrm_weight_offsets_t off;
  ctrm_model_desc desc;
  const uint8_t* fov_w0_tiles;  // digit tiles of the FOV encoders' first layer (fov_tc.cu), chunk-major
  int fov_w0_exp;               // |fov W0| < 2^fov_w0_exp
  int dec_w0z_exp;  // |decoder W0[0:64]| < 2^dec_w0z_exp (digit scale of the exact tensor-core product, tc.cuh)
  // prebuilt shared-memory images of the tensor-core kernels' weight operands (split / digit tiles + fp32 constants), built once
  // at ctrm_load_weights and bulk-copied by every CTA
  const uint8_t* prior_img;
  const uint8_t* decode_img;
  const uint8_t* agent_img;
  const uint8_t* fov_img;
  const uint8_t* inst_w0_img;  // the FOV first layer as a table over the bytes of the binary input, for the per-instance kernel (instance.cu)
};
size_t fov_encode_image_bytes();
int launch_fov_encode_image(const float* d_blob, const ctrm_weight_offsets_t& off, uint8_t* d_image, cudaStream_t st);
size_t cvae_prior_image_bytes();
int launch_cvae_prior_image(const float* d_blob, const ctrm_weight_offsets_t& off, uint8_t* d_image, cudaStream_t st);
size_t cvae_decode_image_bytes();
int launch_cvae_decode_image(const float* d_blob, const ctrm_weight_offsets_t& off, int w0_exp, uint8_t* d_image, cudaStream_t st);
size_t agent_comm_image_bytes();
int launch_agent_comm_image(const float* d_blob, const ctrm_weight_offsets_t& off, uint8_t* d_image, cudaStream_t st);

// device-side view of one uploaded batch and its roll-out state; passed to kernels by value
struct Batch {
  int B, A, O, M, F, L, S, W, K;
  int n_max;  // largest num_agents of the batch
  int max_T, N_traj, max_attempt, dim_ind, randomize_indicator;
  int step_mode;  // select + merge kernel: 0 = by batch size, 1 = one warp per agent, 2 = one thread per agent (3: instance kernel, api.cu)
  // static problem data (flat over agents / obstacles)
  const int32_t *agent_off, *obs_off, *agent_inst;
  const double *starts, *goals, *speeds, *speeds2, *merge2, *rads, *obs_xyr;
  const float4* obs_f4;  // [O] (x, y, rad, 0) rounded to float: conservative far-obstacle rejects off the fp64 pipe
  const uint8_t* occ;   // [B][M][M]
  const uint16_t* ctg;  // [A][field]: cost-to-go in the 8 x 4 micro-tiled layout (common.cuh ctg_index, layout 1)
  // roll-out state
  double *cur, *prev, *nxt, *loc_pred;  // [A][2]
  uint8_t* arrived;                     // [A]
  int32_t *k_traj, *t, *makespan;       // [B]   makespan 0 = None
  uint8_t* done;                        // [B]
  unsigned long long* cnt_collide;      // [B]   StaticObjects.cnt_continuous_collide
  int32_t* n_done;                      // [1]
  int32_t* steps_done;                  // [B] executed roll-out steps (bench accounting)
  int32_t* arrive_cnt;                  // [B] warps of the instance that finished the current step
  int32_t* inst_pack;                   // [B][8] {done, t, k_traj, makespan, agent_off, n_agents, obs_off, n_obs}
  // rows of the per-step kernels: agents of the UNFINISHED instances, instance-contiguous, rebuilt between graph chunks so that
  // the tile kernels do not drag the dead rows of a lock-step batch's tail along (nullptr = identity over all A agents)
  int32_t* live_agent;                  // [A]
  int32_t* live_inst;                   // [A] instance of live_agent[row]: read side by side, not after it
  int32_t* n_live;                      // [1] number of rows
  int32_t* live_off;                    // [B] scratch: first row of an unfinished instance
  // rows of the kernels that only serve the LEARNED branch (agent_comm, cvae_prior, cvae_decode): the agents whose gate draw
  // of the coming step selects the learned sampler (learned_sampler.py:148-157).  An agent that random-walks never reads its
  // CVAE samples, nor its own communication vector; it still runs the FOV encoders (its neighbours read its embedding).
  // Rebuilt by the advance kernel at the end of every step (instance-contiguous); nullptr = every live row.
  int32_t* learn_agent;                 // [A]
  int32_t* learn_inst;                  // [A]
  int32_t* n_learn;                     // [1]
  // agents whose FIRST CVAE candidate failed valid_edge in this step: only they need copies 1 .. K-1 (cvae_decode, second launch)
  int32_t* retry_agent;                 // [A]
  int32_t* retry_inst;                  // [A]
  int32_t* n_retry;                     // [1]
  double* agent_pack;                   // [A][6] {goal x, goal y, speed, speed^2, merge_distance^2, rad}
  // per-step tensors
  uint8_t* fov_tiles;  // bf16 operand tiles [ceil(A/128)][n_chunks][128 x 96]
  float *e_self, *e_vain, *vain_proj, *X, *samples;
  float *cv_sqrtp, *cv_dec0;  // [A][64] sqrt of the clamped prior probabilities, [A][32] x-part of the decoder's first layer
  int32_t* ind;  // [A]
  // timed roadmaps: V[t][i] / E[t][i] as fixed slots + bitmask adjacency
  double* vpos;     // [A][L][S][2]
  int32_t* vcnt;    // [A][L]
  uint32_t* cmask;  // [A][L][S][W]  bit j: edge (t,i) -> (t+1,j)
  uint32_t* pmask;  // [A][L][S][W]  bit j: edge (t-1,j) -> (t,i)
  int32_t* nlayers; // [A] len(trm.V)
  // randomness
  int rng_mode;
  uint32_t seed_lo, seed_hi, inst_base;
  const double *tab_gate, *tab_rw, *tab_step;
  const float *tab_teacher, *tab_uniforms;
  const double* p_rw;  // [(max_T+1)][(max_T+1)]  p_rw[t][D], D = makespan or max_T
  double prob_after_goal;
  // optional trace
  float* tr_samples;
  double *tr_loc_pred, *tr_loc_next;
};

__host__ __device__ inline size_t vslot(const Batch& bt, int a, int t, int s) { return ((size_t)a * bt.L + t) * bt.S + s; }

// ---- launchers (each returns 0 or a negative ctrm_status) ----
int launch_obs_to_cells(const double* d_obs_xyr, int O, int M, int4* d_cells, cudaStream_t st);
int launch_raster_occupancy(const int4* d_cells, const int32_t* d_obs_off, int B, int O, int M, uint8_t* d_occ, cudaStream_t st);
int launch_passable_masks(const uint8_t* d_occ, int B, int M, const uint8_t* d_stencils, const int32_t* d_stencil_r, int st_ld,
                          const int32_t* d_pm_inst, const int32_t* d_pm_stencil, int P, uint32_t* d_blocked /*[B][M][ceil(M/32)]*/,
                          uint32_t* d_pmask, cudaStream_t st);
int launch_wavefront(const uint32_t* d_pmask, const int32_t* d_agent_pm, const double* d_goals, int A, int M, int layout,
                     uint16_t* d_ctg, cudaStream_t st);
int launch_fov_raster(const uint8_t* d_occ, const uint16_t* d_ctg, const int32_t* d_agent_inst, const double* d_pos,
                      const uint8_t* d_skip_inst, int A, int M, int F, int layout, float* d_fov, cudaStream_t st);

int launch_collide_segments(const double* d_from, const double* d_to, const double* d_rad, const int32_t* d_seg_inst,
                            int64_t n, const double* d_obs_xyr, const int32_t* d_obs_off, uint8_t* d_verdict,
                            int32_t* d_survivors, int32_t* d_n_survivors, cudaStream_t st);
int launch_collide_pairs(const double* f1, const double* t1, const double* r1, const double* f2, const double* t2,
                         const double* r2, int dim, int64_t n, uint8_t* verdict, cudaStream_t st);

int launch_fov_raster_tiles(const uint8_t* d_occ, const uint16_t* d_ctg, const int32_t* d_agent_inst, const double* d_pos,
                            const uint8_t* d_skip_inst, int A, int M, int F, int n_chunks, int layout, uint8_t* d_tiles,
                            const int32_t* d_live_agent, const int32_t* d_n_live, cudaStream_t st);
int fov_n_chunks(int F);
size_t fov_tiles_bytes(int A, int F);
size_t fov_w0_tiles_bytes(int F);
int launch_fov_w0_prep(const float* d_w0, int F, int w0_exp, uint8_t* d_out, cudaStream_t st);
int launch_fov_f32_to_tiles(const float* d_fov, int A, int F, uint8_t* d_tiles, cudaStream_t st);
int launch_fov_encode(const DevModel& m, const uint8_t* d_fov_tiles, const int32_t* d_agent_inst, const uint8_t* d_skip_inst,
                      int A, float* d_e_self, float* d_e_vain, float* d_vain_proj, const int32_t* d_live_agent,
                      const int32_t* d_n_live, cudaStream_t st);
// rebuild Batch::live_agent / n_live from the instances' done flags (rollout.cu)
int launch_compact_live(const Batch& bt, cudaStream_t st);
int launch_agent_comm(const DevModel& m, const Batch& bt, const double* d_cur, const double* d_prev, const float* d_e_self,
                      const float* d_vain_proj, float* d_X, cudaStream_t st);
// force_ind >= 0: evaluate the encoder / decoder-x paths with that indicator instead of the argmax (one launch per indicator
// value feeds the per-row random indicators of randomize_indicator=True); d_ind is then not written
int launch_cvae_prior(const DevModel& m, const Batch& bt, const float* d_X, float* d_sqrtp, float* d_dec0, int32_t* d_ind,
                      cudaStream_t st, int force_ind = -1);
// variants != 0: d_sqrtp / d_dec0 hold one [A]-sized slot per indicator value and every row picks its own
// (learned_sampler.py:93-97: torch.randint indicators when the step's draw exceeds p_rw, else the argmax)
int launch_cvae_decode(const DevModel& m, const Batch& bt, const float* d_sqrtp, const float* d_dec0, const float* d_uniforms,
                       int use_state_kt, int kt_fixed, int t_fixed, float* d_samples, cudaStream_t st, int variants = 0,
                       int k_first = 0, int k_count = -1, int emit_retry = 0);

int launch_step(const Batch& bt, cudaStream_t st);
// the whole roll-out loop of an instance inside one kernel, one cluster of C CTAs per instance (instance.cu): small batches
int instance_kernel_supports(const DevModel& m, const Batch& bt);
size_t instance_w0_image_bytes();
int launch_instance_w0_image(const float* d_blob, const ctrm_weight_offsets_t& off, int F, uint8_t* d_image, cudaStream_t st);
int launch_instance_rollout(const DevModel& m, const Batch& bt, int nn, int C, cudaStream_t st);
int launch_rollout_init(const Batch& bt, cudaStream_t st);

int launch_finalize(const Batch& bt, int32_t* d_tfin, cudaStream_t st);
int launch_csr_count(const Batch& bt, int64_t* d_agent_nv, int64_t* d_agent_ne, int64_t* d_totals, cudaStream_t st);
int launch_csr_fill(const Batch& bt, const int64_t* d_agent_v0, const int64_t* d_agent_e0, const int64_t* d_totals,
                    int32_t* d_layer_ptr, double* d_pos, int64_t* d_eptr, int32_t* d_eidx, cudaStream_t st);

// baseline samplers (sampler.cu)
int launch_collide_points(const double* d_pos, const double* d_rad, const int32_t* d_pt_inst, int64_t n, const double* d_obs_xyr,
                          const int32_t* d_obs_off, uint8_t* d_verdict, cudaStream_t st);
int launch_pair_adjacency(const double* d_rows, const int32_t* d_row_off, const double* d_cols, const int32_t* d_col_off,
                          int n_sets, const int32_t* d_set_inst, const double* d_speed, const double* d_speed2, const double* d_rad,
                          const double* d_obs_xyr, const int32_t* d_obs_off, int symmetric, const int64_t* d_bits_off,
                          const int64_t* d_item_off, int64_t n_items, uint32_t* d_bits, int64_t* d_cnt, cudaStream_t st);
int launch_adjacency_csr(const uint32_t* d_bits, const int64_t* d_bits_off, const int32_t* d_row_off, const int32_t* d_col_off,
                         int n_sets, int n_rows, int self_first, int64_t* d_eptr, int32_t* d_eidx, cudaStream_t st);

int launch_csr_fill_compact(const Batch& bt, const int64_t* d_agent_v0, const int64_t* d_agent_e0, const int64_t* d_totals,
                            uint8_t* d_layer_cnt, double* d_pos, uint8_t* d_deg, uint8_t* d_eidx, cudaStream_t st);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: remember what has been granted on each device
// (a process may hold contexts on several GPUs), not once per process.  Contexts are driven from several host threads
// (PipelinedBuilder): check, set and record under one lock, and record only what the driver accepted, so that no thread
// can launch (or capture) the kernel before the attribute call of another thread has taken effect.
struct SmemAttrCache {
  std::mutex mu;
  size_t granted[64] = {0};
  template <typename Kernel>
  cudaError_t ensure(Kernel kernel, size_t smem) {
    std::lock_guard<std::mutex> lock(mu);
    int d = 0;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) return e;
    const bool tracked = d >= 0 && d < 64;
    if (tracked && smem <= granted[d]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && tracked) granted[d] = smem;
    return e;
  }
};

// stream-ordered scratch of the context-free stage entry points: keep freed blocks in the device's default pool instead of
// returning them to the driver at every synchronisation (the default), which made each call pay a fresh allocation
int keep_async_pool_warm();

// prioritized planning over exported CSR roadmaps (plan.cu)
int launch_plan_prioritized(int B, int L, const int32_t* d_agent_off, const int32_t* d_n_layers, const int32_t* d_layer_ptr,
                            const double* d_pos, const int64_t* d_eptr, const int32_t* d_eidx, const double* d_goals,
                            const double* d_speeds, const double* d_rads, int max_nv, int A, int32_t* d_paths, int32_t* d_solved,
                            int32_t* d_failed_agent, long long* d_counters, double* d_sol_pos, cudaStream_t st);

// training-time pairwise features for T time steps (features.cu)
int launch_others_info(const double* d_cur, const double* d_goals, const double* d_prev, const double* d_rads,
                       const double* d_speeds, int T, int N, float* d_info, float* d_self, cudaStream_t st);
