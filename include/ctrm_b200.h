/*
 * ctrm_b200 — C-ABI of the B200-native timed-roadmap construction path.
 *
 * This header is the drop-in boundary: plain pointers and sizes, no torch / pybind11 types.
 * Every entry point cites the interface of the reference (omron-sinicx/ctrm, paths relative to the
 * reference checkout) that it replaces.  All functions return 0 on success and a negative
 * ctrm_status on failure; they never throw and never take ownership of caller memory.
 * `stream` arguments are a cudaStream_t passed as void* (NULL = the legacy default stream).
 * Pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 *
 * There is no CPU fallback anywhere behind this interface: without a CUDA device every compute entry
 * point returns CTRM_ERR_CUDA.
 */
#ifndef CTRM_B200_H
#define CTRM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTRM_B200_ABI_VERSION 1
#define CTRM_UNREACHABLE 65535u /* cost_to_go.cpp:36 max_cost */

typedef enum {
  CTRM_OK = 0,
  CTRM_ERR_ARG = -1,   /* bad argument (null pointer, size out of the supported range) */
  CTRM_ERR_CUDA = -2,  /* a CUDA call failed; see ctrm_last_error() */
  CTRM_ERR_STATE = -3, /* call order violated (e.g. build before weights / instances were loaded) */
  CTRM_ERR_ALLOC = -4
} ctrm_status;

typedef struct ctrm_ctx ctrm_ctx;

/* ---- hyper-parameters of the trained networks (learning/model/*.py, format_input.py:84-142) ---- */
typedef struct {
  int32_t fov_size;      /* F, odd; trained config 19  (fov_encoder.py:16-22)              */
  int32_t map_size;      /* M <= 255 (cost_to_go.cpp:47 uint16 cell index); trained 160    */
  int32_t dim_hidden;    /* 32: every hidden layer of every net (must be 32)               */
  int32_t dim_message;   /* 32 (agent_encoder.py:18-24)                                    */
  int32_t dim_attention; /* 10                                                             */
  int32_t dim_ind;       /* 3  (format_output.py:80-81)                                    */
  int32_t dim_latent;    /* 64 (cvae.py:23-37)                                             */
  int32_t num_neighbors; /* 15 (format_input.py:96)                                        */
  int32_t use_k_neighbor;
  int32_t include_self_attention;
  float temp;            /* 2.0 Gumbel-softmax temperature (cvae.py:133)                   */
} ctrm_model_desc;

/* ---- sampler parameters: the keyword arguments of
 *      get_timed_roadmaps_multiple_paths_with_learned_indicator (learned_sampler.py:19-31) ---- */
typedef struct {
  double prob_uniform_sampling_after_goal; /* 0.9 */
  double prob_uniform_bias;                /* 0.0 */
  double prob_uniform_gamma;               /* 5.0 */
  int32_t max_T;                           /* 64  */
  int32_t N_traj;                          /* 25  */
  int32_t max_attempt;                     /* 3   */
  int32_t randomize_indicator;             /* 0   */
  double merge_distance_rate;              /* 0.25 */
  int32_t K;                               /* candidates per agent-step; 0 -> dim_ind (reference couples them,
                                              learned_sampler.py:82) */
  int32_t rng_mode;                        /* 0 = counter-based Philox4x32-10 keyed (instance,k_traj,t,agent,..);
                                              1 = explicit host tables (replay of a recorded reference run) */
  uint64_t seed;                           /* Philox key (rng_mode 0) */
  uint32_t instance_id_base;               /* Philox counter word 0 of instance b is instance_id_base + b */
} ctrm_params;

/* explicit randomness / teacher forcing for parity runs (all optional, HOST pointers, flat over the batch's
 * agents A = sum n_agents; index [k_traj][t-1][agent(flat)]...):
 *   gate     [N_traj][max_T][A]               f64  np.random.rand() of learned_sampler.py:150-155
 *   rw       [N_traj][max_T][A][max_attempt][3] f64 (u_mag, sin(theta), cos(theta)) of :137-141
 *   step     [N_traj][max_T][B]               f64  the always-consumed draw of :93
 *   teacher  [N_traj][max_T][A][K][3]         f32  SAMPLES to use INSTEAD of the CVAE output (:107-113)
 *   uniforms [N_traj][max_T][A][K][dim_latent] f32 torch.rand of RelaxedOneHotCategorical.rsample (cvae.py:133-136)
 */
typedef struct {
  const double* h_gate;
  const double* h_rw;
  const double* h_step;
  const float* h_teacher;
  const float* h_uniforms;
  /* host-computed constants, so that libm-dependent values are the caller's (numpy / CPython) own:
   *   p_rw   [(max_T+1)][(max_T+1)] f64  prob_random_walk(t, D) of learned_sampler.py:84-91, D = makespan or max_T
   *   merge2 [A] f64  (merge_distance_rate * max_speed) ** 2 of roadmap_learned/utils.py:86-88
   * NULL -> computed by the library with C `exp` / a plain product. */
  const double* h_p_rw;
  const double* h_merge2;
} ctrm_tables;

/* ------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------- */
int ctrm_abi_version(void);
/* one context per (process, device, caller thread).  */
int ctrm_ctx_create(int device, ctrm_ctx** out);
void ctrm_ctx_destroy(ctrm_ctx* ctx);

/* scheduling priority of the context's internal stream (high != 0: the device's greatest priority, else the least).  For two
 * contexts that share one GPU and alternate batches: the kernels of the high-priority context go first, the other context's
 * fill the gaps its host phases (packing, copies) leave, so the two do not finish — and idle the GPU — in lock step.  Call
 * while the context is idle. */
int ctrm_set_stream_priority(ctrm_ctx* ctx, int high);
const char* ctrm_last_error(const ctrm_ctx* ctx); /* ctx may be NULL: process-wide last error */

/* replaces learning/utils.py:32-53 `reconstruct` + model/__init__.py:19-59 (the device half): upload the
 * five networks as one fp32 blob.  BatchNorm (eval mode) is already folded into the preceding Linear by
 * the host (ctrm_b200/weights.py); every matrix is stored [in][out_padded] row-major.  The blob layout is
 * fixed by ctrm_weight_offsets() so host and device agree. */
typedef struct {
  /* offsets in floats into the blob; *_w matrices are [in][ld], biases [ld] */
  int32_t fov_w0, fov_b0;       /* [2F^2][64]  first layers of fov_encoder_self (cols 0..31) | _vain (32..63) */
  int32_t fovs_w1, fovs_b1, fovs_w2, fovs_b2; /* self: [32][32] x2 */
  int32_t fovv_w1, fovv_b1, fovv_w2, fovv_b2; /* vain: [32][32] x2 */
  int32_t ag_w0i, ag_w0v, ag_b0; /* agent encoder layer 0 split: [11][32] (others_info part), [32][32] (fov_vain part) */
  int32_t ag_w1, ag_b1;          /* [32][32] */
  int32_t ag_w2, ag_b2;          /* [32][44] (42 = message 32 + attention 10, padded to 44) */
  int32_t ind_w0, ind_b0, ind_w1, ind_b1, ind_w2, ind_b2; /* indicator 72->32->32->3(ld 4), BN folded */
  int32_t enc_w0, enc_b0, enc_w1, enc_b1, enc_w2, enc_b2; /* encoder_input 75->32->32->64 */
  int32_t dec_w0, dec_b0, dec_w1, dec_b1, dec_w2, dec_b2; /* decoder 139->32->32->3(ld 4) */
  int32_t total;
} ctrm_weight_offsets_t;
int ctrm_weight_offsets(const ctrm_model_desc* desc, ctrm_weight_offsets_t* out);
int ctrm_load_weights(ctrm_ctx* ctx, const ctrm_model_desc* desc, const float* h_blob, size_t n_floats);

/* replaces environment/instance.py:17-46 + Format2D_CTRM_Input.set_instance (format_input.py:200-210):
 * copies B problem definitions (fp64, SoA, flat over agents / obstacles), rasterises the occupancy maps
 * and runs the batched cost-to-go wavefront for every goal.  speeds2[a] = max_speed**2 and
 * merge2 are computed by the host in Python-float arithmetic (roadmap/utils.py:40, utils.py:86-88). */
int ctrm_upload_instances(ctrm_ctx* ctx, int B, const int32_t* h_n_agents, const double* h_starts /*[A][2]*/,
                          const double* h_goals /*[A][2]*/, const double* h_speeds /*[A]*/,
                          const double* h_speeds2 /*[A]*/, const double* h_rads /*[A]*/,
                          const int32_t* h_n_obs, const double* h_obs_xyr /*[O][3]*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * stage entry points (device pointers; each is one kernel family that can be profiled alone)
 * ---------------------------------------------------------------------------------------------- */

/* replaces sphere_collision_check_wrapper.continuousCollideSpheres (collision_check.cpp:16-58) as used by
 * StaticObjects.collide_continuous_sphere (static_objects.py:108-138): for each of n_seg swept circles
 * (from,to,rad) of instance seg_inst[s], verdict[s] = any over that instance's obstacles.  fp64, the
 * reference's operation order, no FMA.  d_survivors (optional) receives the compacted indices of the
 * NON-colliding segments in ascending order (warp ballot + prefix scan), *d_n_survivors their count. */
int ctrm_collide_segments(const double* d_from /*[n][2]*/, const double* d_to /*[n][2]*/, const double* d_rad /*[n]*/,
                          const int32_t* d_seg_inst /*[n] or NULL (all instance 0)*/, int64_t n_seg,
                          const double* d_obs_xyr /*[O][3]*/, const int32_t* d_obs_off /*[B+1]*/,
                          uint8_t* d_verdict /*[n]*/, int32_t* d_survivors /*[n] or NULL*/,
                          int32_t* d_n_survivors /*[1] or NULL*/, void* stream);

/* general pairwise form of the same test (both spheres moving; dim 2 or 3) — the full signature of
 * continuousCollideSpheres(from1,to1,rad1,from2,to2,rad2) batched over n pairs. */
int ctrm_collide_sphere_pairs(const double* d_from1, const double* d_to1, const double* d_rad1, const double* d_from2,
                              const double* d_to2, const double* d_rad2, int dim, int64_t n, uint8_t* d_verdict,
                              void* stream);

/* replaces StaticObjects.draw_2d (static_objects.py:173-187) / ObstacleSphere.draw_2d (obstacle.py:63-81):
 * occ[b][x][y] in {0,1}, uint8, first axis x. */
int ctrm_raster_occupancy(const double* d_obs_xyr, const int32_t* d_obs_off, int B, int M, uint8_t* d_occ, void* stream);

/* replaces get_approx_cost_to_go_matrix_2d (learning/fov.py:34-63) + getCostToGo2D (cost_to_go.cpp:25-107),
 * batched over A goals: ctg[a][x][y] uint16 hop count, 65535 = unreachable.  Agent footprint stencil
 * r = ceil(rad*M) (r computed by the HOST in Python float arithmetic -> d_stencil_r). */
int ctrm_cost_to_go(const uint8_t* d_occ /*[B][M][M]*/, const int32_t* d_agent_inst /*[A]*/,
                    const double* d_goals /*[A][2]*/, const int32_t* d_stencil_r /*[A]*/, int A, int M,
                    uint16_t* d_ctg /*[A][M][M]*/, void* stream);

/* replaces getFov2dOccupancyCost (cost_to_go.cpp:109-186) + torch.tensor(...).float() (format_input.py:299-308)
 * for all agents at once: fov[a][0..2F^2) float32 {0,1}, flatten layout idx = x_fov*F + y_fov, channel 1 at F^2.
 * ctg_layout 0 = row-major [A][M][M] as ctrm_cost_to_go / the reference return it; 1 = the context's own 8 x 4 micro-tiled
 * fields (ctrm_device_view.d_ctg), which keep a crop's DRAM traffic near its algorithmic size. */
int ctrm_fov_raster(const uint8_t* d_occ, const uint16_t* d_ctg, int ctg_layout, const int32_t* d_agent_inst,
                    const double* d_pos /*[A][2]*/, int A, int M, int F, float* d_fov /*[A][2F^2]*/, void* stream);

/* replaces Format2D_CTRM_Input.get_feature_one_step after the FOV crop (format_input.py:311-357):
 * FOV encoders, numba get_arr_others_info (compiled.py:30-97), AgentEncoder, get_comm (format_input.py:212-277),
 * get_self_info (compiled.py:100-125).  X[a][72] = [self_info 8 | fov_self 32 | comm 32].  Uses the ctx's
 * uploaded weights and instance table (agent offsets, goals, rads, speeds). */
int ctrm_encode_features(ctrm_ctx* ctx, const float* d_fov /*[A][2F^2]*/, const double* d_cur /*[A][2]*/,
                         const double* d_prev /*[A][2]*/, float* d_X /*[A][72]*/, void* stream);

/* replaces indicator+argmax+one_hot (learned_sampler.py:99-104) and CTRMNet.sample (cvae.py:128-138) for K copies
 * per agent: samples[a][k][3] = (dx,dy,mag).  d_uniforms [A][K][dim_latent] may be NULL -> Philox stream 8 keyed
 * by (seed; instance, k_traj, t, row k*N+i, column/4).
 * d_ind (optional) receives the argmax indicator per agent. */
int ctrm_cvae_sample(ctrm_ctx* ctx, const float* d_X /*[A][72]*/, const float* d_uniforms, int K, uint64_t seed,
                     int k_traj, int t, float* d_samples /*[A][K][3]*/, int32_t* d_ind /*[A] or NULL*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * the fused driver: replaces get_timed_roadmaps_multiple_paths_with_learned_indicator
 * (learned_sampler.py:19-204) incl. merge_samples / format_trms / append_goals (roadmap_learned/utils.py:21-174)
 * for the whole uploaded batch.  The step loop (FOV raster -> encoders -> CVAE -> candidate validation ->
 * merge/edge building -> advance) is captured once in a CUDA graph and replayed; every instance carries its
 * own (k_traj, t) so finished roll-outs never wait for the others.
 * ---------------------------------------------------------------------------------------------- */
int ctrm_build_roadmaps(ctrm_ctx* ctx, const ctrm_params* params, const ctrm_tables* tables /*or NULL*/, void* stream);

/* after ctrm_build_roadmaps: totals of the CSR result (format after append_goals). */
int ctrm_result_sizes(ctrm_ctx* ctx, int64_t* n_vertices, int64_t* n_edges, int64_t* n_layers_total);

/* CSR export.  Vertex order: instance, agent, layer t, index i (the reference's V[t][i]);
 *   layer_ptr [A*(L)+1]   first vertex of (agent a, layer t) at layer_ptr[a*L + t], L = max_T+2; layers beyond an
 *                         agent's length are empty
 *   pos       [nv][2] f64 ;  eptr [nv+1] ; eidx [ne]  = indices into the NEXT layer, ascending
 *   n_layers  [A] = len(trm.V) ;  cnt_collide [B] = StaticObjects.cnt_continuous_collide (static_objects.py:72-75)
 * dst_is_device selects whether the destination pointers are device or host memory. */
int ctrm_export_csr(ctrm_ctx* ctx, int32_t* layer_ptr, double* pos, int64_t* eptr, int32_t* eidx, int32_t* n_layers,
                    int64_t* cnt_collide, int dst_is_device, void* stream);

/* The same result in the COMPACT form meant for the trip to the host (0.75 MB instead of 1.46 MB per aamas22 instance): a
 * layer holds at most 32 * ceil((N_traj + 1) / 32) <= 128 vertices, so vertices per layer (layer_cnt[a][t], zero beyond the
 * agent's last layer), out-degree per vertex (deg[v]) and successor index (eidx[e]) are single bytes; positions stay fp64.
 * agent_v0 / agent_e0 [A + 1] = first vertex / first edge of every agent (last entry = totals), so any instance or agent can be
 * expanded on its own: layer_ptr = prefix sum of its layer_cnt row, eptr = prefix sum of its deg slice.  Same ownership and
 * stream rules as ctrm_export_csr; n_layers / cnt_collide / agent_v0 / agent_e0 may be NULL. */
int ctrm_export_csr_compact(ctrm_ctx* ctx, uint8_t* layer_cnt /*[A][L]*/, double* pos /*[nv][2]*/, uint8_t* deg /*[nv]*/,
                            uint8_t* eidx /*[ne]*/, int64_t* agent_v0 /*[A+1] or NULL*/, int64_t* agent_e0 /*[A+1] or NULL*/,
                            int32_t* n_layers /*[A] or NULL*/, int64_t* cnt_collide /*[B] or NULL*/, int dst_is_device,
                            void* stream);

/* per-step trace for parity tests (optional): when enabled before ctrm_build_roadmaps, the driver records for every
 * executed (instance, k_traj, t): samples computed by the CVAE [N_traj][max_T][A][K][3] f32, loc_pred and loc_next
 * [N_traj][max_T][A][2] f64 (NaN where the step was not executed). */
int ctrm_trace_enable(ctrm_ctx* ctx, int enable);
int ctrm_trace_read(ctrm_ctx* ctx, float* h_samples, double* h_loc_pred, double* h_loc_next);

/* device timing of the last ctrm_build_roadmaps, measured with CUDA events on its stream:
 * ms[0] = total, ms[1] = step loop, ms[2] = finalisation; steps = graph replays; launches = kernels launched. */
int ctrm_last_timing(ctrm_ctx* ctx, float ms[3], int64_t* steps, int64_t* launches);

/* profile mode: the next ctrm_build_roadmaps calls launch every step kernel eagerly (no CUDA graph) with a CUDA event
 * after each kernel on the driver's stream; ctrm_last_kernel_times returns the summed device time (ms) and the
 * launch count per kernel, in the order fov_raster, fov_encode, agent_comm, cvae_prior, cvae_decode, step (candidate
 * selection + merge + roll-out bookkeeping); the remaining slots are zero.  Used by bench.py for the roofline of the dominant kernel. */
/* executed roll-out steps per instance of the last build (sum over its N_traj roll-outs) */
int ctrm_get_instance_steps(ctrm_ctx* ctx, int32_t* h_steps /*[B]*/);
int ctrm_set_profile(ctrm_ctx* ctx, int enable);

/* which kernels run the roll-out loop (learned_sampler.py:65-196):
 *   0 = chosen by batch size (default): up to one instance per SM (148 on B200; CTRM_INSTANCE_MAX_BATCH overrides) takes the
 *       per-instance kernel (3) when the batch fits its limits (<= 64 agents per instance, K <= 8, temperature 2,
 *       <= 15 neighbours), larger batches the batch path with select + merge as (2) from ~20k agents and as (1) below;
 *   1 = batch path (one kernel sequence per step over all instances), select + merge one warp per agent;
 *   2 = batch path, select + merge one thread per agent (throughput, large batches);
 *   3 = per-instance kernel: one thread-block cluster per instance runs all N_traj x max_T steps inside ONE launch (single-
 *       instance latency); CTRM_ERR_ARG at build time if the batch does not fit its limits.
 * 1 and 2 produce identical roadmaps and counters.  3 evaluates the networks on the CUDA cores in plain fp32: samples agree
 * with 1/2 to ~1e-6, geometry and roadmap updates are the same code, results are bit-identical under teacher forcing. */
int ctrm_set_step_kernel(ctrm_ctx* ctx, int mode);
int ctrm_last_kernel_times(ctrm_ctx* ctx, double ms[8], int64_t launches[8]);

/* raw views of the context's device state for stage-level tests (valid until the next upload) */
typedef struct {
  int32_t B, A, O, M, F, L, S, W;
  const int32_t *d_agent_off, *d_obs_off, *d_agent_inst;
  const double *d_starts, *d_goals, *d_speeds, *d_speeds2, *d_rads, *d_obs_xyr;
  const uint8_t* d_occ;
  const uint16_t* d_ctg;
  int32_t ctg_layout; /* layout of d_ctg, to be passed to ctrm_fov_raster */
} ctrm_device_view;
int ctrm_get_device_view(ctrm_ctx* ctx, ctrm_device_view* out);

/* ------------------------------------------------------------------------------------------------
 * host-buffer convenience wrappers with the pybind11 modules' exact call shapes (B = 1); they copy in,
 * launch the kernels above on the GPU and copy out.
 * ---------------------------------------------------------------------------------------------- */
/* Training-time pairwise features: replaces numba get_arr_others_info (learning/compiled.py:30-97)
 * and get_self_info (compiled.py:100-125) as Dataset.map_instance_to_tensor calls them (learning/dataset.py:135-160) for all T
 * time steps of a demonstration: info[t][i*N + j][0:11] float32 (float64 arithmetic, one final cast, bit-exact),
 * self_info[t][i][0:8] (optional) = columns 3..10 of row (i, i).  The FOV half of that function is ctrm_fov_raster. */
int ctrm_others_info(const double* d_cur /*[T][N][2]*/, const double* d_goals /*[N][2]*/, const double* d_prev /*[T][N][2]*/,
                     const double* d_rads /*[N]*/, const double* d_speeds /*[N]*/, int T, int N,
                     float* d_info /*[T][N*N][11]*/, float* d_self_info /*[T][N][8] or NULL*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Baseline samplers on the collision / edge kernels: the device work behind
 * get_timed_roadmaps_random / _grid / _random_common / _grid_common / _fully_random (roadmap/random_sampler.py:22-78,
 * roadmap/grid_sampler.py:19-55, roadmap/utils.py:45-134, roadmap/timed_roadmap.py:169-210).
 * ---------------------------------------------------------------------------------------------- */

/* replaces StaticObjects.collide_sphere (environment/static_objects.py:44-57, :93-106) for n points at once:
 * verdict[i] = max(pos + rad) > 1 or min(pos - rad) < 0 or the circle touches any obstacle of instance pt_inst[i].  The
 * reference delegates the obstacle test to FCL 0.5.0 (absent here, Dockerfile:63); restated from FCL's sphereSphereIntersect:
 * collide <=> !(||c1 - c2|| > r1 + r2), fp64. */
int ctrm_collide_points(const double* d_pos /*[n][2]*/, const double* d_rad /*[n]*/, const int32_t* d_pt_inst /*[n] or NULL*/,
                        int64_t n, const double* d_obs_xyr /*[O][3]*/, const int32_t* d_obs_off /*[B+1]*/,
                        uint8_t* d_verdict /*[n]*/, void* stream);

/* number of 32-bit words ctrm_pair_adjacency writes: sum over sets of rows_s * ceil(cols_s / 32); -1 on bad offsets */
int64_t ctrm_pair_adjacency_words(const int32_t* h_row_off /*[n_sets+1]*/, const int32_t* h_col_off /*[n_sets+1]*/, int n_sets);

/* replaces the all-pairs valid_move loops (roadmap/utils.py:18-42 called from timed_roadmap.py:186-193, :116-131, :259-265 and
 * roadmap/utils.py:114-120): for every set s (one agent's samples) bit (i, j) of its row-major bit matrix is set iff
 * valid_move(rows[i], cols[j], speed_s, rad_s, obstacles of instance set_inst[s]).  symmetric != 0: rows == columns, the pair
 * is tested once as (min(i,j) -> max(i,j)) like the reference's `for j in range(i + 1, n)` and mirrored, the diagonal stays
 * clear.  d_cnt[s] (optional) = number of collide_continuous_sphere calls the reference makes for the set (pairs that pass the
 * L-inf and L2 pre-checks; symmetric: each pair once).  Offsets and per-set constants are small HOST arrays (speed2 = the
 * caller's own max_speed ** 2); positions, obstacles and outputs are DEVICE pointers. */
int ctrm_pair_adjacency(const double* d_rows /*[R][2]*/, const int32_t* h_row_off /*[n_sets+1]*/, const double* d_cols /*[C][2]*/,
                        const int32_t* h_col_off /*[n_sets+1]*/, int n_sets, const int32_t* h_set_inst /*[n_sets] or NULL*/,
                        const double* h_speed /*[n_sets]*/, const double* h_speed2 /*[n_sets]*/, const double* h_rad /*[n_sets]*/,
                        const double* d_obs_xyr /*[O][3]*/, const int32_t* d_obs_off /*[B+1]*/, int symmetric,
                        uint32_t* d_bits /*[ctrm_pair_adjacency_words]*/, int64_t* d_cnt /*[n_sets] or NULL*/, void* stream);

/* the adjacency lists of those bit matrices as CSR by popcount + exclusive prefix scan + ranked scatter.  Two passes:
 * d_eidx == NULL fills d_eptr[rows + 1] (d_eptr[rows] = number of entries); with d_eidx the entries are written: per row the
 * row's own local index first when self_first != 0 (the reference's `adjs = [[i] for i ...]`), then the set columns in
 * ascending order — exactly the order the reference's nested loops append them in. */
int ctrm_adjacency_csr(const uint32_t* d_bits, const int32_t* h_row_off /*[n_sets+1]*/, const int32_t* h_col_off /*[n_sets+1]*/,
                       int n_sets, int self_first, int64_t* d_eptr /*[rows+1]*/, int32_t* d_eidx /*[entries] or NULL*/,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Consumer of the roadmaps: replaces PrioritizedPlanning._solve
 * (planner/prioritized_planning.py:40-101) with its space-time A* (Planner.get_single_agent_path, planner/planner.py:311-384)
 * and its collision test against earlier agents (collide_dynamic_agents, planner.py:216-263 -> continuousCollideSpheres),
 * batched over B instances.  All pointers are DEVICE pointers; the roadmaps come in the CSR layout of ctrm_export_csr
 * (layer_ptr with a fixed stride of L slots per agent).  paths[a][t] = index within V[t] of the vertex agent a occupies at
 * time t (the goal vertex once arrived; -1 beyond the instance's horizon or when the instance is unsolved, as the reference
 * clears its solution); path_pos[a][t] the matching coordinates; solved[b] 0/1; failed_agent[b] = first agent without a
 * path, -1 if solved, -2 if a vertex has more than 128 successors (unsupported); counters[b] = {cnt_continuous_collide, lowlevel_expanded, lowlevel_explored} of the reference's Result.
 * max_agent_vertices = largest number of vertices of one agent's roadmap (sizes the per-CTA shared memory). */
int ctrm_plan_prioritized(int B, int A, int L, const int32_t* d_agent_off /*[B+1]*/, const int32_t* d_n_layers /*[A]*/,
                          const int32_t* d_layer_ptr /*[A*L+1]*/, const double* d_pos /*[nv][2]*/,
                          const int64_t* d_eptr /*[nv+1]*/, const int32_t* d_eidx /*[ne]*/, const double* d_goals /*[A][2]*/,
                          const double* d_speeds /*[A]*/, const double* d_rads /*[A]*/, int max_agent_vertices,
                          int32_t* d_paths /*[A][L]*/, double* d_path_pos /*[A][L][2]*/, int32_t* d_solved /*[B]*/,
                          int32_t* d_failed_agent /*[B]*/, int64_t* d_counters /*[B][3]*/, void* stream);

/* sphere_collision_check_wrapper.continuousCollideSpheres(from1,to1,rad1,from2,to2,rad2) -> bool
 * (collision_check.cpp:16-24, :60-67) */
int ctrm_host_continuous_collide_spheres(const double* from1, const double* to1, double rad1, const double* from2,
                                         const double* to2, double rad2, int dim, int* out_collide);
/* cost_to_go_wrapper.getCostToGo2D(goal, occupancy_map, occupancy_agent) -> int32[M][M] (cost_to_go.cpp:25-107);
 * occupancy_agent is the (2r+1)^2 stencil. */
int ctrm_host_cost_to_go_2d(const double* goal, const int32_t* occupancy_map, int M, const int32_t* occupancy_agent,
                            int agent_size, int32_t* out_cost /*[M][M]*/);
/* cost_to_go_wrapper.getFov2dOccupancyCost(current_pos, occupancy_map, cost_to_go, fov_size, flatten)
 * (cost_to_go.cpp:109-186); cost_to_go passed as int32 with negative = unreachable (the reference's inf -> INT_MIN). */
int ctrm_host_fov2d(const double* pos, const int32_t* occupancy_map, const int32_t* cost_to_go, int M, int F,
                    int32_t* out_fov /*[2][F][F]*/);

#ifdef __cplusplus
}
#endif
#endif /* CTRM_B200_H */
