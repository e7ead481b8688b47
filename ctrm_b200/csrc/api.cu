// C-ABI of libctrm_b200 (include/ctrm_b200.h): context, uploads, stage entry points, the graph-captured
// roll-out driver and the host-buffer wrappers with the pybind11 helpers' call shapes.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

int ctrm_fail(int code, const char* what) {
  g_last_error = what ? what : "";
  return code;
}
int ctrm_fail_cuda(cudaError_t e, const char* what) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what ? what : "?");
  g_last_error = buf;
  return CTRM_ERR_CUDA;
}

#define CHECK(expr)         \
  do {                      \
    int _r = (expr);        \
    if (_r != 0) return _r; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevArena {  // tracked device allocations with reuse: cudaMalloc / cudaFree of GB-sized buffers cost tens of
                   // milliseconds and serialise the device, and successive batches of the same shape ask for the same sizes
  struct Blk {
    void* p;
    size_t bytes;
    bool used;
  };
  std::vector<Blk> blks;
  template <typename T>
  int alloc(T** out, size_t n) {
    *out = nullptr;
    if (n == 0) n = 1;
    const size_t bytes = n * sizeof(T);
    for (Blk& b : blks)
      if (!b.used && b.bytes == bytes) {
        b.used = true;
        *out = reinterpret_cast<T*>(b.p);
        return 0;
      }
    // no exact fit: the smallest free block that is large enough and not wastefully larger.  Batches of a stream differ a
    // little in size (agents, vertices, edges); without this every upload / export of a new size pays cudaMalloc + cudaFree,
    // and cudaFree synchronises the whole device — it stalls the OTHER context of a PipelinedBuilder in mid roll-out
    Blk* best = nullptr;
    for (Blk& b : blks)
      if (!b.used && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + 4096 && (!best || b.bytes < best->bytes)) best = &b;
    if (best) {
      best->used = true;
      *out = reinterpret_cast<T*>(best->p);
      return 0;
    }
    void* p = nullptr;
    const size_t grown = bytes >= (1u << 20) ? bytes + bytes / 16 : bytes;  // head room for the next, slightly larger batch
    cudaError_t e = cudaMalloc(&p, grown);
    if (e != cudaSuccess && grown != bytes) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      if (e == cudaSuccess) { blks.push_back({p, bytes, true}); *out = reinterpret_cast<T*>(p); return 0; }
    }
    if (e != cudaSuccess) return ctrm_fail_cuda(e, "cudaMalloc");
    blks.push_back({p, grown, true});
    *out = reinterpret_cast<T*>(p);
    return 0;
  }
  void release() {  // keep the memory for the next round of allocations
    for (Blk& b : blks) b.used = false;
  }
  void trim() {  // free whatever the last round did not take again
    size_t w = 0;
    for (size_t i = 0; i < blks.size(); ++i) {
      if (blks[i].used) blks[w++] = blks[i];
      else cudaFree(blks[i].p);
    }
    blks.resize(w);
  }
  void destroy() {
    for (Blk& b : blks) cudaFree(b.p);
    blks.clear();
  }
};

struct ctrm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // internal non-blocking stream of the driver (graph capture is illegal on stream 0)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_t[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_block = nullptr;  // cudaEventBlockingSync: the host thread sleeps instead of spinning while a long chunk runs
  // weights
  DevArena w_arena;
  DevModel model{};
  bool have_weights = false;
  // batch (static part) + per-step tensors
  DevArena b_arena;
  Batch bt{};
  bool have_batch = false;
  std::vector<int32_t> h_agent_off, h_obs_off;
  std::vector<double> h_speeds;
  // roll-out state + roadmap store (sized by params)
  DevArena r_arena;
  int r_L = 0, r_S = 0, r_W = 0, r_K = 0, r_NT = 0, r_MT = 0, r_att = 0;
  bool r_trace = false;
  // tables
  DevArena t_arena;
  // scratch of ctrm_upload_instances (footprint stencils, passability masks): cached across batches
  DevArena s_arena;
  // results
  DevArena c_arena;
  int32_t* d_tfin = nullptr;
  int32_t *d_learn_agent = nullptr, *d_learn_inst = nullptr, *d_n_learn = nullptr;  // Batch::learn_* storage
  int32_t *d_retry_agent = nullptr, *d_retry_inst = nullptr, *d_n_retry = nullptr;  // Batch::retry_* storage
  int64_t *d_agent_nv = nullptr, *d_agent_ne = nullptr, *d_totals = nullptr;
  int64_t totals[3] = {0, 0, 0};
  bool have_result = false;
  bool trace = false;
  int32_t* h_ndone = nullptr;  // pinned
  float ms[3] = {0, 0, 0};
  int64_t steps = 0, launches = 0;
  bool profile = false;
  int step_mode = 0;
  double kernel_ms[16] = {0};
  int64_t kernel_n[16] = {0};
};

// wait for the context's stream.  Large batches (a chunk of steps runs for tens of milliseconds, an export for a hundred): the
// thread sleeps on a blocking event — several contexts per GPU and several ranks per box otherwise keep one host core each
// spinning in cudaStreamSynchronize and starve the threads that pack, copy and launch.  Small batches: spin (lowest latency).
static cudaError_t wait_stream(ctrm_ctx* c, bool blocking) {
  if (!blocking) return cudaStreamSynchronize(c->stream);
  cudaError_t e = cudaEventRecord(c->ev_block, c->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(c->ev_block);
}

static int sync_in(ctrm_ctx* c, cudaStream_t user) {
  CTRM_CUDA_OK(cudaEventRecord(c->ev_in, user));
  CTRM_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_in, 0));
  return 0;
}
static int sync_out(ctrm_ctx* c, cudaStream_t user) {
  CTRM_CUDA_OK(cudaEventRecord(c->ev_out, c->stream));
  CTRM_CUDA_OK(cudaStreamWaitEvent(user, c->ev_out, 0));
  return 0;
}

extern "C" int ctrm_abi_version(void) { return CTRM_B200_ABI_VERSION; }

extern "C" const char* ctrm_last_error(const ctrm_ctx*) { return g_last_error.c_str(); }

extern "C" int ctrm_ctx_create(int device, ctrm_ctx** out) {
  if (!out) return ctrm_fail(CTRM_ERR_ARG, "ctrm_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return ctrm_fail_cuda(e, "cudaGetDeviceCount (no CUDA device: ctrm_b200 has no CPU fallback)");
  if (device < 0 || device >= n) return ctrm_fail(CTRM_ERR_ARG, "ctrm_ctx_create: no such device");
  CTRM_CUDA_OK(cudaSetDevice(device));
  ctrm_ctx* c = new ctrm_ctx();
  c->device = device;
  CTRM_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CTRM_CUDA_OK(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
  CTRM_CUDA_OK(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
  for (int i = 0; i < 4; ++i) CTRM_CUDA_OK(cudaEventCreate(&c->ev_t[i]));
  CTRM_CUDA_OK(cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
  CTRM_CUDA_OK(cudaMallocHost(&c->h_ndone, sizeof(int32_t)));
  *out = c;
  return 0;
}

// Contexts that share a GPU (PipelinedBuilder) finish their batches in lock step when their kernels time-share the device
// evenly: both then export / pack / upload at the same moment and the GPU idles.  A higher stream priority for one of them
// staggers the two: its kernels go first, the other context fills the gaps its host phases leave.
extern "C" int ctrm_set_stream_priority(ctrm_ctx* c, int high) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  CTRM_CUDA_OK(cudaSetDevice(c->device));
  CTRM_CUDA_OK(cudaStreamSynchronize(c->stream));
  int least = 0, greatest = 0;
  CTRM_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  cudaStream_t ns = nullptr;
  CTRM_CUDA_OK(cudaStreamCreateWithPriority(&ns, cudaStreamNonBlocking, high ? greatest : least));
  cudaStreamDestroy(c->stream);
  c->stream = ns;
  return 0;
}

extern "C" void ctrm_ctx_destroy(ctrm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->w_arena.destroy();
  c->b_arena.destroy();
  c->r_arena.destroy();
  c->t_arena.destroy();
  c->s_arena.destroy();
  c->c_arena.destroy();
  if (c->h_ndone) cudaFreeHost(c->h_ndone);
  for (int i = 0; i < 4; ++i)
    if (c->ev_t[i]) cudaEventDestroy(c->ev_t[i]);
  if (c->ev_block) cudaEventDestroy(c->ev_block);
  if (c->ev_in) cudaEventDestroy(c->ev_in);
  if (c->ev_out) cudaEventDestroy(c->ev_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
static int check_desc(const ctrm_model_desc* d) {
  if (!d) return ctrm_fail(CTRM_ERR_ARG, "model desc is NULL");
  if (d->dim_hidden != 32 || d->dim_message != 32 || d->dim_attention != 10 || d->dim_latent != 64 || d->dim_ind != 3)
    return ctrm_fail(CTRM_ERR_ARG, "unsupported network dims (need hidden 32, message 32, attention 10, latent 64, ind 3)");
  if (d->fov_size < 1 || (d->fov_size & 1) == 0 || d->fov_size > 63) return ctrm_fail(CTRM_ERR_ARG, "fov_size must be odd, <= 63");
  if (d->map_size < 1 || d->map_size > 255) return ctrm_fail(CTRM_ERR_ARG, "map_size must be in 1..255 (reference limit)");
  return 0;
}

extern "C" int ctrm_weight_offsets(const ctrm_model_desc* d, ctrm_weight_offsets_t* o) {
  CHECK(check_desc(d));
  if (!o) return ctrm_fail(CTRM_ERR_ARG, "out is NULL");
  int p = 0;
  auto take = [&](int n) {
    int r = p;
    p += (n + 3) & ~3;  // keep every block 16-byte aligned
    return r;
  };
  const int KD = 2 * d->fov_size * d->fov_size;
  o->fov_w0 = take(KD * 64); o->fov_b0 = take(64);
  o->fovs_w1 = take(32 * 32); o->fovs_b1 = take(32); o->fovs_w2 = take(32 * 32); o->fovs_b2 = take(32);
  o->fovv_w1 = take(32 * 32); o->fovv_b1 = take(32); o->fovv_w2 = take(32 * 32); o->fovv_b2 = take(32);
  o->ag_w0i = take(11 * 32); o->ag_w0v = take(32 * 32); o->ag_b0 = take(32);
  o->ag_w1 = take(32 * 32); o->ag_b1 = take(32);
  o->ag_w2 = take(32 * 44); o->ag_b2 = take(44);
  o->ind_w0 = take(72 * 32); o->ind_b0 = take(32); o->ind_w1 = take(32 * 32); o->ind_b1 = take(32);
  o->ind_w2 = take(32 * 4); o->ind_b2 = take(4);
  o->enc_w0 = take(75 * 32); o->enc_b0 = take(32); o->enc_w1 = take(32 * 32); o->enc_b1 = take(32);
  o->enc_w2 = take(32 * 64); o->enc_b2 = take(64);
  o->dec_w0 = take(139 * 32); o->dec_b0 = take(32); o->dec_w1 = take(32 * 32); o->dec_b1 = take(32);
  o->dec_w2 = take(32 * 4); o->dec_b2 = take(4);
  o->total = p;
  return 0;
}

extern "C" int ctrm_load_weights(ctrm_ctx* c, const ctrm_model_desc* d, const float* h_blob, size_t n_floats) {
  if (!c || !h_blob) return ctrm_fail(CTRM_ERR_ARG, "ctrm_load_weights: NULL argument");
  ctrm_weight_offsets_t off;
  CHECK(ctrm_weight_offsets(d, &off));
  if (n_floats != (size_t)off.total) return ctrm_fail(CTRM_ERR_ARG, "ctrm_load_weights: blob size does not match ctrm_weight_offsets");
  CTRM_CUDA_OK(cudaSetDevice(c->device));
  CTRM_CUDA_OK(cudaStreamSynchronize(c->stream));
  c->w_arena.release();
  float* blob = nullptr;
  CHECK(c->w_arena.alloc(&blob, n_floats));
  CTRM_CUDA_OK(cudaMemcpy(blob, h_blob, n_floats * sizeof(float), cudaMemcpyHostToDevice));
  c->model.blob = blob;
  c->model.off = off;
  c->model.desc = *d;
  auto exp_bound = [&](int first, int count) {  // smallest e >= -20 with max|w| < 2^e
    float mxw = 0.f;
    for (int i = 0; i < count; ++i) mxw = std::fmax(mxw, std::fabs(h_blob[first + i]));
    int e = -20;
    while (std::ldexp(1.0f, e) <= mxw && e < 30) ++e;
    return e;
  };
  c->model.dec_w0z_exp = exp_bound(off.dec_w0, 64 * 32);
  {  // digit tiles of the FOV encoders' first layer for the tensor-core kernel
    const int F = d->fov_size;
    c->model.fov_w0_exp = exp_bound(off.fov_w0, 2 * F * F * 64);
    uint8_t* tiles = nullptr;
    CHECK(c->w_arena.alloc(&tiles, fov_w0_tiles_bytes(F)));
    CHECK(launch_fov_w0_prep(blob + off.fov_w0, F, c->model.fov_w0_exp, tiles, c->stream));
    CTRM_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->model.fov_w0_tiles = tiles;
  }
  {  // shared-memory weight images of the other tensor-core kernels
    uint8_t *pi = nullptr, *di = nullptr, *ai = nullptr, *fi = nullptr, *wi = nullptr;
    CHECK(c->w_arena.alloc(&pi, cvae_prior_image_bytes()));
    CHECK(c->w_arena.alloc(&di, cvae_decode_image_bytes()));
    CHECK(c->w_arena.alloc(&ai, agent_comm_image_bytes()));
    CHECK(c->w_arena.alloc(&fi, fov_encode_image_bytes()));
    CHECK(launch_fov_encode_image(blob, off, fi, c->stream));
    CHECK(launch_cvae_prior_image(blob, off, pi, c->stream));
    CHECK(launch_cvae_decode_image(blob, off, c->model.dec_w0z_exp, di, c->stream));
    CHECK(launch_agent_comm_image(blob, off, ai, c->stream));
    CHECK(c->w_arena.alloc(&wi, instance_w0_image_bytes()));
    CHECK(launch_instance_w0_image(blob, off, d->fov_size, wi, c->stream));
    CTRM_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->model.prior_img = pi; c->model.decode_img = di; c->model.agent_img = ai; c->model.fov_img = fi;
    c->model.inst_w0_img = wi;
  }
  c->have_weights = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// instances
// ------------------------------------------------------------------------------------------------
template <typename T>
static int upload(DevArena& ar, T** dptr, const T* h, size_t n, cudaStream_t st) {
  CHECK(ar.alloc(dptr, n));
  if (n) CTRM_CUDA_OK(cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

// disk((r,r), r) on a (2r+1)^2 grid: k^2 + l^2 < r^2  (learning/fov.py:48-56; identical to skimage's float test)
static void disk_stencil(int r, int ld, uint8_t* out) {
  for (int k = -r; k <= r; ++k)
    for (int l = -r; l <= r; ++l) out[(k + r) * ld + (l + r)] = (k * k + l * l < r * r) ? 1 : 0;
}

static int cost_to_go_device(const uint8_t* d_occ, int M, const std::vector<uint8_t>& stencils, const std::vector<int32_t>& st_r,
                             int st_ld, const std::vector<int32_t>& pm_inst, const std::vector<int32_t>& pm_st,
                             const std::vector<int32_t>& agent_pm, const double* d_goals, int A, uint16_t* d_ctg,
                             cudaStream_t st, DevArena* cache = nullptr, int layout = 0) {
  DevArena local;
  DevArena& tmp = cache ? *cache : local;  // a context passes its scratch arena: no cudaMalloc / cudaFree per batch
  uint8_t* d_st = nullptr;
  int32_t *d_str = nullptr, *d_pmi = nullptr, *d_pms = nullptr, *d_apm = nullptr;
  uint32_t* d_pmask = nullptr;
  int rc = 0;
  const int P = (int)pm_inst.size(), MW = (M + 31) / 32;
  do {
    if ((rc = upload(tmp, &d_st, stencils.data(), stencils.size(), st))) break;
    if ((rc = upload(tmp, &d_str, st_r.data(), st_r.size(), st))) break;
    if ((rc = upload(tmp, &d_pmi, pm_inst.data(), pm_inst.size(), st))) break;
    if ((rc = upload(tmp, &d_pms, pm_st.data(), pm_st.size(), st))) break;
    if ((rc = upload(tmp, &d_apm, agent_pm.data(), agent_pm.size(), st))) break;
    if ((rc = tmp.alloc(&d_pmask, (size_t)P * M * MW))) break;
    int nB = 0;  // maps referenced by the masks
    for (int32_t b : pm_inst) nB = b + 1 > nB ? b + 1 : nB;
    uint32_t* d_blocked = nullptr;
    if ((rc = tmp.alloc(&d_blocked, (size_t)nB * M * MW))) break;
    if ((rc = launch_passable_masks(d_occ, nB, M, d_st, d_str, st_ld, d_pmi, d_pms, P, d_blocked, d_pmask, st))) break;
    if ((rc = launch_wavefront(d_pmask, d_apm, d_goals, A, M, layout, d_ctg, st))) break;
    cudaError_t e = cudaStreamSynchronize(st);  // the host vectors and temporaries die here
    if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "cost_to_go sync");
  } while (0);
  if (cache) cache->release();
  else local.destroy();
  return rc;
}

extern "C" int ctrm_upload_instances(ctrm_ctx* c, int B, const int32_t* n_agents, const double* starts, const double* goals,
                                     const double* speeds, const double* speeds2, const double* rads, const int32_t* n_obs,
                                     const double* obs_xyr, void* user_stream) {
  if (!c || B <= 0 || !n_agents || !starts || !goals || !speeds || !speeds2 || !rads || !n_obs)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_upload_instances: NULL / empty argument");
  if (!c->have_weights) return ctrm_fail(CTRM_ERR_STATE, "ctrm_upload_instances: load weights first (map_size / fov_size come from the model)");
  CTRM_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  CHECK(sync_in(c, (cudaStream_t)user_stream));
  CTRM_CUDA_OK(cudaStreamSynchronize(st));
  c->b_arena.release();
  const int prev_A = c->have_batch ? c->bt.A : -1;
  c->have_batch = false;
  c->have_result = false;
  const int M = c->model.desc.map_size, F = c->model.desc.fov_size;
  std::vector<int32_t>& aoff = c->h_agent_off;
  std::vector<int32_t>& ooff = c->h_obs_off;
  aoff.assign(B + 1, 0);
  ooff.assign(B + 1, 0);
  int n_max = 0;
  for (int b = 0; b < B; ++b) {
    if (n_agents[b] < 1 || n_obs[b] < 0) return ctrm_fail(CTRM_ERR_ARG, "ctrm_upload_instances: bad agent / obstacle count");
    aoff[b + 1] = aoff[b] + n_agents[b];
    ooff[b + 1] = ooff[b] + n_obs[b];
    if (n_agents[b] > n_max) n_max = n_agents[b];
  }
  const int A = aoff[B], O = ooff[B];
  if (O > 0 && !obs_xyr) return ctrm_fail(CTRM_ERR_ARG, "ctrm_upload_instances: obstacles missing");
  std::vector<int32_t> ainst(A);
  for (int b = 0; b < B; ++b)
    for (int a = aoff[b]; a < aoff[b + 1]; ++a) ainst[a] = b;
  c->h_speeds.assign(speeds, speeds + A);
  // footprint stencils r = ceil(rad * M) (learning/fov.py:48) and the (instance, r) pairs that need a passability mask
  std::vector<int32_t> st_r;
  std::map<int, int> r_to_st;
  std::map<std::pair<int, int>, int> pm_of;
  std::vector<int32_t> pm_inst, pm_st, agent_pm(A);
  int rmax = 0;
  for (int a = 0; a < A; ++a) {
    int r = (int)std::ceil(rads[a] * (double)M);
    if (r < 0 || r > 127) return ctrm_fail(CTRM_ERR_ARG, "ctrm_upload_instances: agent radius out of range");
    if (!r_to_st.count(r)) {
      r_to_st[r] = (int)st_r.size();
      st_r.push_back(r);
      if (r > rmax) rmax = r;
    }
    auto key = std::make_pair(ainst[a], r_to_st[r]);
    if (!pm_of.count(key)) {
      pm_of[key] = (int)pm_inst.size();
      pm_inst.push_back(key.first);
      pm_st.push_back(key.second);
    }
    agent_pm[a] = pm_of[key];
  }
  const int st_ld = 2 * rmax + 1;
  std::vector<uint8_t> stencils(st_r.size() * (size_t)st_ld * st_ld, 0);
  for (size_t i = 0; i < st_r.size(); ++i) disk_stencil(st_r[i], st_ld, stencils.data() + i * (size_t)st_ld * st_ld);

  Batch& bt = c->bt;
  const Batch old_bt = bt;
  bt = Batch{};
  bt.B = B; bt.A = A; bt.O = O; bt.M = M; bt.F = F; bt.n_max = n_max;
  DevArena& ar = c->b_arena;
  int32_t *d_aoff, *d_ooff, *d_ainst;
  double *d_starts, *d_goals, *d_speeds, *d_speeds2, *d_rads, *d_obs, *d_merge2;
  CHECK(upload(ar, &d_aoff, aoff.data(), (size_t)B + 1, st));
  CHECK(upload(ar, &d_ooff, ooff.data(), (size_t)B + 1, st));
  CHECK(upload(ar, &d_ainst, ainst.data(), (size_t)A, st));
  CHECK(upload(ar, &d_starts, starts, (size_t)2 * A, st));
  CHECK(upload(ar, &d_goals, goals, (size_t)2 * A, st));
  CHECK(upload(ar, &d_speeds, speeds, (size_t)A, st));
  CHECK(upload(ar, &d_speeds2, speeds2, (size_t)A, st));
  CHECK(upload(ar, &d_rads, rads, (size_t)A, st));
  CHECK(upload(ar, &d_obs, obs_xyr, (size_t)3 * O, st));
  std::vector<float4> obs_f4((size_t)(O > 0 ? O : 1), make_float4(0.f, 0.f, 0.f, 0.f));
  for (int o = 0; o < O; ++o) obs_f4[o] = make_float4((float)obs_xyr[3 * o], (float)obs_xyr[3 * o + 1], (float)obs_xyr[3 * o + 2], 0.f);
  float4* d_obs_f4 = nullptr;
  CHECK(upload(ar, &d_obs_f4, obs_f4.data(), obs_f4.size(), st));
  CHECK(ar.alloc(&d_merge2, (size_t)A));
  bt.agent_off = d_aoff; bt.obs_off = d_ooff; bt.agent_inst = d_ainst;
  bt.starts = d_starts; bt.goals = d_goals; bt.speeds = d_speeds; bt.speeds2 = d_speeds2; bt.rads = d_rads;
  bt.obs_xyr = d_obs; bt.merge2 = d_merge2; bt.obs_f4 = d_obs_f4;
  uint8_t* d_occ;
  uint16_t* d_ctg;
  CHECK(ar.alloc(&d_occ, (size_t)B * M * M + 16));
  CHECK(ar.alloc(&d_ctg, (size_t)A * ctg_field_elems(M, 1)));
  bt.occ = d_occ; bt.ctg = d_ctg;
  // per-step tensors (A-sized)
  CHECK(ar.alloc(&bt.fov_tiles, fov_tiles_bytes(A, F)));
  CTRM_CUDA_OK(cudaMemsetAsync(bt.fov_tiles, 0, fov_tiles_bytes(A, F), st));
  CHECK(ar.alloc(&bt.e_self, (size_t)A * 32));
  CHECK(ar.alloc(&bt.e_vain, (size_t)A * 32));
  CHECK(ar.alloc(&bt.vain_proj, (size_t)A * 32));
  CHECK(ar.alloc(&bt.X, (size_t)A * CTRM_DIM_X));
  // one [A]-sized slot per indicator value: slot 0 alone in the usual run, all of them with randomize_indicator
  CHECK(ar.alloc(&bt.cv_sqrtp, (size_t)A * 64 * c->model.desc.dim_ind));
  CHECK(ar.alloc(&bt.cv_dec0, (size_t)A * 32 * c->model.desc.dim_ind));
  CHECK(ar.alloc(&bt.ind, (size_t)A));
  CHECK(ar.alloc(&bt.cur, (size_t)2 * A));
  CHECK(ar.alloc(&bt.prev, (size_t)2 * A));
  CHECK(ar.alloc(&bt.nxt, (size_t)2 * A));
  CHECK(ar.alloc(&bt.loc_pred, (size_t)2 * A));
  CHECK(ar.alloc(&bt.arrived, (size_t)A));
  CHECK(ar.alloc(&bt.k_traj, (size_t)B));
  CHECK(ar.alloc(&bt.t, (size_t)B));
  CHECK(ar.alloc(&bt.makespan, (size_t)B));
  CHECK(ar.alloc(&bt.done, (size_t)B));
  CHECK(ar.alloc(&bt.cnt_collide, (size_t)B));
  CHECK(ar.alloc(&bt.n_done, (size_t)1));
  CHECK(ar.alloc(&bt.steps_done, (size_t)B));
  CHECK(ar.alloc(&bt.arrive_cnt, (size_t)B));
  CHECK(ar.alloc(&bt.inst_pack, (size_t)8 * B));
  CHECK(ar.alloc(&bt.live_agent, (size_t)A));
  CHECK(ar.alloc(&bt.live_inst, (size_t)A));
  CHECK(ar.alloc(&bt.n_live, (size_t)1));
  CHECK(ar.alloc(&bt.live_off, (size_t)B));
  CHECK(ar.alloc(&c->d_learn_agent, (size_t)A));
  CHECK(ar.alloc(&c->d_learn_inst, (size_t)A));
  CHECK(ar.alloc(&c->d_n_learn, (size_t)1));
  CHECK(ar.alloc(&c->d_retry_agent, (size_t)A));
  CHECK(ar.alloc(&c->d_retry_inst, (size_t)A));
  CHECK(ar.alloc(&c->d_n_retry, (size_t)1));
  CHECK(ar.alloc(&bt.agent_pack, (size_t)6 * A));
  CHECK(ar.alloc(&bt.nlayers, (size_t)A));
  CHECK(ar.alloc(&c->d_tfin, (size_t)B));
  CHECK(ar.alloc(&c->d_agent_nv, (size_t)A));
  CHECK(ar.alloc(&c->d_agent_ne, (size_t)A));
  CHECK(ar.alloc(&c->d_totals, (size_t)4));
  ar.trim();
  if (A != prev_A) {  // the roadmap store is sized by the agent count: re-carve it from the arena (whose blocks are reused
    c->r_L = c->r_S = 0;  // when the new batch is about the same size) at the next build
  } else {
    bt.L = old_bt.L; bt.S = old_bt.S; bt.W = old_bt.W; bt.K = old_bt.K;
    bt.vpos = old_bt.vpos; bt.vcnt = old_bt.vcnt; bt.cmask = old_bt.cmask; bt.pmask = old_bt.pmask;
    bt.samples = old_bt.samples;
    bt.tr_samples = old_bt.tr_samples; bt.tr_loc_pred = old_bt.tr_loc_pred; bt.tr_loc_next = old_bt.tr_loc_next;
  }
  CTRM_CUDA_OK(cudaMemsetAsync(bt.done, 0, (size_t)B, st));
  // occupancy raster + cost-to-go for every goal (Format2D_CTRM_Input.set_instance, format_input.py:200-210)
  {
    int4* d_cells = nullptr;
    int rc = c->s_arena.alloc(&d_cells, (size_t)O);
    if (!rc) rc = launch_obs_to_cells(d_obs, O, M, d_cells, st);
    if (!rc) rc = launch_raster_occupancy(d_cells, d_ooff, B, O, M, d_occ, st);
    if (!rc) {
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "raster sync");
    }
    c->s_arena.release();
    CHECK(rc);
  }
  CHECK(cost_to_go_device(d_occ, M, stencils, st_r, st_ld, pm_inst, pm_st, agent_pm, d_goals, A, d_ctg, st, &c->s_arena, 1));
  c->have_batch = true;
  CHECK(sync_out(c, (cudaStream_t)user_stream));
  return 0;
}

extern "C" int ctrm_get_device_view(ctrm_ctx* c, ctrm_device_view* v) {
  if (!c || !v) return ctrm_fail(CTRM_ERR_ARG, "NULL argument");
  if (!c->have_batch) return ctrm_fail(CTRM_ERR_STATE, "no batch uploaded");
  const Batch& bt = c->bt;
  v->B = bt.B; v->A = bt.A; v->O = bt.O; v->M = bt.M; v->F = bt.F; v->L = bt.L; v->S = bt.S; v->W = bt.W;
  v->d_agent_off = bt.agent_off; v->d_obs_off = bt.obs_off; v->d_agent_inst = bt.agent_inst;
  v->d_starts = bt.starts; v->d_goals = bt.goals; v->d_speeds = bt.speeds; v->d_speeds2 = bt.speeds2; v->d_rads = bt.rads;
  v->d_obs_xyr = bt.obs_xyr; v->d_occ = bt.occ; v->d_ctg = bt.ctg;
  v->ctg_layout = 1;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// stage entry points
// ------------------------------------------------------------------------------------------------
// The context-free entry points take bare device pointers: make the device that owns them current, so that a caller whose
// current device is another GPU does not launch on a foreign stream (and the per-device caches key on the right device).
static int use_device_of(const void* p) {
  if (!p) return 0;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }  // not a CUDA pointer: the kernels will report it
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) CTRM_CUDA_OK(cudaSetDevice(at.device));
  return 0;
}
extern "C" int ctrm_collide_segments(const double* d_from, const double* d_to, const double* d_rad, const int32_t* d_seg_inst,
                                     int64_t n_seg, const double* d_obs_xyr, const int32_t* d_obs_off, uint8_t* d_verdict,
                                     int32_t* d_survivors, int32_t* d_n_survivors, void* stream) {
  CHECK(use_device_of(d_verdict));
  if (n_seg < 0 || (n_seg > 0 && (!d_from || !d_to || !d_rad || !d_obs_off || !d_verdict)))
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_collide_segments: NULL argument");
  if ((d_survivors == nullptr) != (d_n_survivors == nullptr))
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_collide_segments: survivors and n_survivors go together");
  return launch_collide_segments(d_from, d_to, d_rad, d_seg_inst, n_seg, d_obs_xyr, d_obs_off, d_verdict, d_survivors,
                                 d_n_survivors, (cudaStream_t)stream);
}

extern "C" int ctrm_collide_sphere_pairs(const double* f1, const double* t1, const double* r1, const double* f2, const double* t2,
                                         const double* r2, int dim, int64_t n, uint8_t* verdict, void* stream) {
  CHECK(use_device_of(verdict));
  if (n < 0 || (n > 0 && (!f1 || !t1 || !r1 || !f2 || !t2 || !r2 || !verdict)))
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_collide_sphere_pairs: NULL argument");
  return launch_collide_pairs(f1, t1, r1, f2, t2, r2, dim, n, verdict, (cudaStream_t)stream);
}

int keep_async_pool_warm() {
  static bool done[64] = {false};
  int dev = 0;
  CTRM_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || done[dev]) return 0;
  cudaMemPool_t pool;
  CTRM_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t keep = 256ull << 20;  // scratch of these entry points is a few MB; cap what the pool may hold back
  CTRM_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  done[dev] = true;
  return 0;
}

extern "C" int ctrm_raster_occupancy(const double* d_obs_xyr, const int32_t* d_obs_off, int B, int M, uint8_t* d_occ, void* stream) {
  CHECK(use_device_of(d_occ));
  if (B <= 0 || M < 1 || M > 255 || !d_obs_off || !d_occ) return ctrm_fail(CTRM_ERR_ARG, "ctrm_raster_occupancy: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t O = 0;
  CTRM_CUDA_OK(cudaMemcpyAsync(&O, d_obs_off + B, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CTRM_CUDA_OK(cudaStreamSynchronize(st));
  int4* d_cells = nullptr;
  CHECK(keep_async_pool_warm());
  CTRM_CUDA_OK(cudaMallocAsync(&d_cells, sizeof(int4) * (size_t)(O > 0 ? O : 1), st));
  int rc = launch_obs_to_cells(d_obs_xyr, O, M, d_cells, st);
  if (!rc) rc = launch_raster_occupancy(d_cells, d_obs_off, B, O, M, d_occ, st);
  cudaFreeAsync(d_cells, st);
  return rc;
}

extern "C" int ctrm_cost_to_go(const uint8_t* d_occ, const int32_t* d_agent_inst, const double* d_goals,
                               const int32_t* d_stencil_r, int A, int M, uint16_t* d_ctg, void* stream) {
  CHECK(use_device_of(d_ctg));
  if (A <= 0 || M < 1 || M > 255 || !d_occ || !d_agent_inst || !d_goals || !d_stencil_r || !d_ctg)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_cost_to_go: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int32_t> inst(A), rr(A);
  CTRM_CUDA_OK(cudaMemcpyAsync(inst.data(), d_agent_inst, sizeof(int32_t) * A, cudaMemcpyDeviceToHost, st));
  CTRM_CUDA_OK(cudaMemcpyAsync(rr.data(), d_stencil_r, sizeof(int32_t) * A, cudaMemcpyDeviceToHost, st));
  CTRM_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<int32_t> st_r, pm_inst, pm_st, agent_pm(A);
  std::map<int, int> r_to_st;
  std::map<std::pair<int, int>, int> pm_of;
  int rmax = 0;
  for (int a = 0; a < A; ++a) {
    int r = rr[a];
    if (r < 0 || r > 127) return ctrm_fail(CTRM_ERR_ARG, "ctrm_cost_to_go: stencil radius out of range");
    if (!r_to_st.count(r)) { r_to_st[r] = (int)st_r.size(); st_r.push_back(r); if (r > rmax) rmax = r; }
    auto key = std::make_pair(inst[a], r_to_st[r]);
    if (!pm_of.count(key)) { pm_of[key] = (int)pm_inst.size(); pm_inst.push_back(key.first); pm_st.push_back(key.second); }
    agent_pm[a] = pm_of[key];
  }
  const int st_ld = 2 * rmax + 1;
  std::vector<uint8_t> stencils(st_r.size() * (size_t)st_ld * st_ld, 0);
  for (size_t i = 0; i < st_r.size(); ++i) disk_stencil(st_r[i], st_ld, stencils.data() + i * (size_t)st_ld * st_ld);
  return cost_to_go_device(d_occ, M, stencils, st_r, st_ld, pm_inst, pm_st, agent_pm, d_goals, A, d_ctg, st);
}

extern "C" int ctrm_fov_raster(const uint8_t* d_occ, const uint16_t* d_ctg, int ctg_layout, const int32_t* d_agent_inst,
                               const double* d_pos, int A, int M, int F, float* d_fov, void* stream) {
  CHECK(use_device_of(d_fov));
  if (A <= 0 || M < 1 || M > 255 || F < 1 || (F & 1) == 0 || F > 63 || !d_occ || !d_ctg || !d_agent_inst || !d_pos || !d_fov)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_fov_raster: bad argument");
  if ((reinterpret_cast<uintptr_t>(d_fov) & 15) != 0) return ctrm_fail(CTRM_ERR_ARG, "ctrm_fov_raster: d_fov must be 16-byte aligned");
  if (ctg_layout != 0 && ctg_layout != 1) return ctrm_fail(CTRM_ERR_ARG, "ctrm_fov_raster: ctg_layout must be 0 or 1");
  return launch_fov_raster(d_occ, d_ctg, d_agent_inst, d_pos, nullptr, A, M, F, ctg_layout, d_fov, (cudaStream_t)stream);
}

extern "C" int ctrm_encode_features(ctrm_ctx* c, const float* d_fov, const double* d_cur, const double* d_prev, float* d_X,
                                    void* stream) {
  if (!c || !d_fov || !d_cur || !d_prev || !d_X) return ctrm_fail(CTRM_ERR_ARG, "ctrm_encode_features: NULL argument");
  if (!c->have_weights || !c->have_batch) return ctrm_fail(CTRM_ERR_STATE, "ctrm_encode_features: weights and instances first");
  cudaStream_t st = (cudaStream_t)stream;
  Batch bt = c->bt;
  bt.done = nullptr;
  bt.live_agent = nullptr;
  bt.learn_agent = nullptr;
  CHECK(launch_fov_f32_to_tiles(d_fov, bt.A, bt.F, bt.fov_tiles, st));
  CHECK(launch_fov_encode(c->model, bt.fov_tiles, bt.agent_inst, nullptr, bt.A, bt.e_self, bt.e_vain, bt.vain_proj, nullptr, nullptr,
                          st));
  CHECK(launch_agent_comm(c->model, bt, d_cur, d_prev, bt.e_self, bt.vain_proj, d_X, st));
  return 0;
}

extern "C" int ctrm_cvae_sample(ctrm_ctx* c, const float* d_X, const float* d_uniforms, int K, uint64_t seed, int k_traj, int t,
                                float* d_samples, int32_t* d_ind, void* stream) {
  if (!c || !d_X || !d_samples || K < 1) return ctrm_fail(CTRM_ERR_ARG, "ctrm_cvae_sample: bad argument");
  if (!c->have_weights || !c->have_batch) return ctrm_fail(CTRM_ERR_STATE, "ctrm_cvae_sample: weights and instances first");
  Batch bt = c->bt;
  bt.done = nullptr;
  bt.live_agent = nullptr;
  bt.learn_agent = nullptr;
  bt.K = K;
  bt.dim_ind = c->model.desc.dim_ind;
  bt.seed_lo = (uint32_t)(seed & 0xFFFFFFFFu);
  bt.seed_hi = (uint32_t)(seed >> 32);
  bt.inst_base = 0;
  CHECK(launch_cvae_prior(c->model, bt, d_X, bt.cv_sqrtp, bt.cv_dec0, d_ind, (cudaStream_t)stream));
  return launch_cvae_decode(c->model, bt, bt.cv_sqrtp, bt.cv_dec0, d_uniforms, 0, k_traj, t, d_samples, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// the driver
// ------------------------------------------------------------------------------------------------
extern "C" int ctrm_trace_enable(ctrm_ctx* c, int enable) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  c->trace = enable != 0;
  return 0;
}

static int ensure_rollout_store(ctrm_ctx* c, const ctrm_params* p, int K) {
  Batch& bt = c->bt;
  const int L = p->max_T + 2;
  const int W = (p->N_traj + 1 + 31) / 32;
  const int S = 32 * W;
  const bool same = c->r_L == L && c->r_S == S && c->r_K == K && c->r_NT == p->N_traj && c->r_MT == p->max_T &&
                    c->r_att == p->max_attempt && c->r_trace == c->trace;
  bt.L = L; bt.S = S; bt.W = W; bt.K = K;
  if (same) return 0;
  c->r_arena.release();
  DevArena& ar = c->r_arena;
  const size_t slots = (size_t)bt.A * L * S;
  CHECK(ar.alloc(&bt.vpos, slots * 2));
  CHECK(ar.alloc(&bt.vcnt, (size_t)bt.A * L));
  CHECK(ar.alloc(&bt.cmask, slots * W));
  CHECK(ar.alloc(&bt.pmask, slots * W));
  CHECK(ar.alloc(&bt.samples, (size_t)bt.A * K * 3));
  bt.tr_samples = nullptr; bt.tr_loc_pred = nullptr; bt.tr_loc_next = nullptr;
  if (c->trace) {
    const size_t steps = (size_t)p->N_traj * p->max_T;
    CHECK(ar.alloc(&bt.tr_samples, steps * bt.A * K * 3));
    CHECK(ar.alloc(&bt.tr_loc_pred, steps * bt.A * 2));
    CHECK(ar.alloc(&bt.tr_loc_next, steps * bt.A * 2));
  }
  ar.trim();
  c->r_L = L; c->r_S = S; c->r_W = W; c->r_K = K; c->r_NT = p->N_traj; c->r_MT = p->max_T; c->r_att = p->max_attempt;
  c->r_trace = c->trace;
  return 0;
}

// kernel ids of the per-kernel timing table (ctrm_last_kernel_times)
enum { KID_FOV_RASTER = 0, KID_FOV_ENCODE, KID_AGENT_COMM, KID_CVAE_PRIOR, KID_CVAE_DECODE, KID_STEP, KID_STEP_COUNT };

// profile mode: the step runs eagerly with a CUDA event after every kernel
struct StepProfiler {
  std::vector<cudaEvent_t> ev;  // [chunk][KID_STEP_COUNT + 1]
  int chunk = 0, cur = 0;
  double ms[KID_STEP_COUNT] = {0};
  int64_t n[KID_STEP_COUNT] = {0};
  bool seen[KID_STEP_COUNT] = {false};
  int init(int chunk_) {
    chunk = chunk_;
    ev.resize((size_t)chunk * (KID_STEP_COUNT + 1));
    for (auto& e : ev) CTRM_CUDA_OK(cudaEventCreate(&e));
    return 0;
  }
  cudaEvent_t at(int step, int k) { return ev[(size_t)step * (KID_STEP_COUNT + 1) + k]; }
  int flush(int steps_in_chunk, const bool* ran) {  // after a stream sync
    for (int s = 0; s < steps_in_chunk; ++s) {
      int prev = 0;
      for (int k = 0; k < KID_STEP_COUNT; ++k) {
        if (!ran[k]) continue;
        float t = 0.f;
        CTRM_CUDA_OK(cudaEventElapsedTime(&t, at(s, prev), at(s, k + 1)));
        ms[k] += t;
        n[k] += 1;
        prev = k + 1;
      }
    }
    return 0;
  }
  void destroy() {
    for (auto& e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

// batches up to this many instances take the per-instance kernel (CTRM_INSTANCE_MAX_BATCH moves the threshold, 0 = never).
// Default: the number of SMs — one CTA per instance still has every instance resident at once and measured 114 ms against the
// batch path's 148 ms at 148 instances (profiles/r2_instance_kernel.md); a second wave of clusters doubles its time.
static int instance_kernel_max_batch(int device) {
  const char* e = getenv("CTRM_INSTANCE_MAX_BATCH");
  if (e) return atoi(e) > 0 ? atoi(e) : 0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 64;
  return sms;
}
// CTAs per instance: 0 = chosen by the launcher (the largest of 8, 4, 2, 1 that keeps every cluster resident at once);
// CTRM_INSTANCE_CLUSTER pins 1, 2, 4 or 8
static int instance_cluster_size() {
  const char* e = getenv("CTRM_INSTANCE_CLUSTER");
  if (e) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) return v;
  }
  return 0;
}

// one roll-out step for every unfinished instance; `nn` = run the networks (skipped for teacher-forced runs
// that do not trace the CVAE output)
static int enqueue_step(ctrm_ctx* c, const Batch& bt, bool nn, cudaStream_t st, int* n_kernels, StepProfiler* prof = nullptr,
                        int pstep = 0, bool* ran = nullptr) {
  int k = 0;
  auto mark = [&](int kid) -> int {
    ++k;
    if (prof) {
      CTRM_CUDA_OK(cudaEventRecord(prof->at(pstep, kid + 1), st));
      ran[kid] = true;
    }
    return 0;
  };
  if (prof) CTRM_CUDA_OK(cudaEventRecord(prof->at(pstep, 0), st));
  if (nn) {
    CHECK(launch_fov_raster_tiles(bt.occ, bt.ctg, bt.agent_inst, bt.cur, bt.done, bt.A, bt.M, bt.F, fov_n_chunks(bt.F), 1,
                                  bt.fov_tiles, bt.live_agent, bt.n_live, st));
    CHECK(mark(KID_FOV_RASTER));
    CHECK(launch_fov_encode(c->model, bt.fov_tiles, bt.agent_inst, bt.done, bt.A, bt.e_self, nullptr, bt.vain_proj,
                            bt.live_agent, bt.n_live, st));
    CHECK(mark(KID_FOV_ENCODE));
    // the three kernels that only serve the learned branch run over the learned rows of this step (Batch::learn_agent)
    Batch bl = bt;
    if (bt.learn_agent != nullptr) {
      bl.live_agent = bt.learn_agent;
      bl.live_inst = bt.learn_inst;
      bl.n_live = bt.n_learn;
    }
    CHECK(launch_agent_comm(c->model, bl, bt.cur, bt.prev, bt.e_self, bt.vain_proj, bt.X, st));
    CHECK(mark(KID_AGENT_COMM));
    CHECK(launch_cvae_prior(c->model, bl, bt.X, bt.cv_sqrtp, bt.cv_dec0, bt.ind, st));
    if (bt.randomize_indicator) {  // learned_sampler.py:93-97: the encoder / decoder-x paths for every indicator value; the
      for (int v = 0; v < bt.dim_ind; ++v) {  // decoder rows pick theirs (random per row, or the argmax written above)
        CHECK(launch_cvae_prior(c->model, bl, bt.X, bt.cv_sqrtp + (size_t)v * bt.A * 64, bt.cv_dec0 + (size_t)v * bt.A * 32, nullptr,
                                st, v));
        ++k;
      }
    }
    CHECK(mark(KID_CVAE_PRIOR));
    if (bt.n_retry != nullptr) {
      CHECK(launch_cvae_decode(c->model, bl, bt.cv_sqrtp, bt.cv_dec0, nullptr, 1, 0, 0, bt.samples, st, bt.randomize_indicator, 0, 1, 1));
      Batch br = bt;  // the agents whose first candidate failed
      br.live_agent = bt.retry_agent;
      br.live_inst = bt.retry_inst;
      br.n_live = bt.n_retry;
      CHECK(launch_cvae_decode(c->model, br, bt.cv_sqrtp, bt.cv_dec0, nullptr, 1, 0, 0, bt.samples, st, bt.randomize_indicator, 1,
                               bt.K - 1, 0));
      ++k;
    } else {
      CHECK(launch_cvae_decode(c->model, bl, bt.cv_sqrtp, bt.cv_dec0, nullptr, 1, 0, 0, bt.samples, st, bt.randomize_indicator));
    }
    CHECK(mark(KID_CVAE_DECODE));
  }
  CHECK(launch_step(bt, st));
  ++k;  // launch_step is two kernels: step, then advance
  CHECK(mark(KID_STEP));
  *n_kernels = k;
  return 0;
}

extern "C" int ctrm_build_roadmaps(ctrm_ctx* c, const ctrm_params* p, const ctrm_tables* tb, void* user_stream) {
  if (!c || !p) return ctrm_fail(CTRM_ERR_ARG, "ctrm_build_roadmaps: NULL argument");
  if (!c->have_weights || !c->have_batch) return ctrm_fail(CTRM_ERR_STATE, "ctrm_build_roadmaps: weights and instances first");
  if (p->N_traj < 0 || p->N_traj > 32 * CTRM_MAX_W - 1) return ctrm_fail(CTRM_ERR_ARG, "N_traj must be in 0..127");
  if (p->max_T < 1 || p->max_T > 4096) return ctrm_fail(CTRM_ERR_ARG, "max_T must be in 1..4096");
  if (p->max_attempt < 0 || p->max_attempt > 255) return ctrm_fail(CTRM_ERR_ARG, "max_attempt must be in 0..255");
  if (p->randomize_indicator && p->rng_mode == 1)
    return ctrm_fail(CTRM_ERR_ARG, "randomize_indicator=True needs counter-based randomness (rng_mode 0): the recorded-stream "
                                   "tables hold no torch.randint draws");
  const int K = p->K > 0 ? p->K : c->model.desc.dim_ind;
  if (K > 4096) return ctrm_fail(CTRM_ERR_ARG, "K too large");
  if (p->rng_mode == 1 && (!tb || !tb->h_gate || (p->max_attempt > 0 && !tb->h_rw)))
    return ctrm_fail(CTRM_ERR_ARG, "rng_mode 1 needs the gate and rw tables");
  CTRM_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  CHECK(sync_in(c, (cudaStream_t)user_stream));
  CHECK(ensure_rollout_store(c, p, K));
  Batch& bt = c->bt;
  bt.max_T = p->max_T; bt.N_traj = p->N_traj; bt.max_attempt = p->max_attempt;
  bt.dim_ind = c->model.desc.dim_ind; bt.randomize_indicator = p->randomize_indicator ? 1 : 0;
  bt.rng_mode = p->rng_mode;
  bt.seed_lo = (uint32_t)(p->seed & 0xFFFFFFFFu); bt.seed_hi = (uint32_t)(p->seed >> 32);
  bt.inst_base = p->instance_id_base;
  bt.prob_after_goal = p->prob_uniform_sampling_after_goal;
  bt.step_mode = c->step_mode;
  // learned-row list: on, except when the run records the CVAE output of EVERY agent-step (trace) — the random-walking
  // agents' samples are never read by the path, but the parity tests compare them
  bt.learn_agent = c->trace ? nullptr : c->d_learn_agent;
  bt.learn_inst = c->trace ? nullptr : c->d_learn_inst;
  bt.n_learn = c->trace ? nullptr : c->d_n_learn;
  // lazy candidates: the first CVAE copy for every learned agent, the others only where it fails valid_edge.  Off when the run
  // records every sample (trace) and when the step takes its samples from a teacher table (validity must be tested on the
  // samples the step will use)
  const bool lazy = !c->trace && !(tb && tb->h_teacher) && K > 1;
  bt.retry_agent = lazy ? c->d_retry_agent : nullptr;
  bt.retry_inst = lazy ? c->d_retry_inst : nullptr;
  bt.n_retry = lazy ? c->d_n_retry : nullptr;
  c->have_result = false;
  // tables
  CTRM_CUDA_OK(cudaStreamSynchronize(st));
  c->t_arena.release();
  bt.tab_gate = bt.tab_rw = bt.tab_step = nullptr;
  bt.tab_teacher = bt.tab_uniforms = nullptr;
  const size_t steps_max = (size_t)p->N_traj * p->max_T;
  {
    const int MT1 = p->max_T + 1;
    std::vector<double> prw((size_t)MT1 * MT1, 0.0);
    if (tb && tb->h_p_rw) {
      memcpy(prw.data(), tb->h_p_rw, prw.size() * sizeof(double));
    } else {
      for (int t = 1; t <= p->max_T; ++t)
        for (int D = 1; D <= p->max_T; ++D)
          prw[(size_t)t * MT1 + D] =
              (1.0 - std::exp(-p->prob_uniform_gamma * t / D)) * (1.0 - p->prob_uniform_bias) + p->prob_uniform_bias;
    }
    double* d = nullptr;
    CHECK(upload(c->t_arena, &d, prw.data(), prw.size(), st));
    bt.p_rw = d;
    std::vector<double> m2(bt.A);
    if (tb && tb->h_merge2) {
      memcpy(m2.data(), tb->h_merge2, m2.size() * sizeof(double));
    } else {
      for (int a = 0; a < bt.A; ++a) {
        double md = p->merge_distance_rate * c->h_speeds[a];
        m2[a] = md * md;
      }
    }
    CTRM_CUDA_OK(cudaMemcpyAsync(const_cast<double*>(bt.merge2), m2.data(), m2.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CTRM_CUDA_OK(cudaStreamSynchronize(st));
  }
  if (tb) {
    double* d = nullptr;
    float* f = nullptr;
    if (tb->h_gate) { CHECK(upload(c->t_arena, &d, tb->h_gate, steps_max * bt.A, st)); bt.tab_gate = d; }
    if (tb->h_rw && p->max_attempt > 0) { CHECK(upload(c->t_arena, &d, tb->h_rw, steps_max * bt.A * p->max_attempt * 3, st)); bt.tab_rw = d; }
    if (tb->h_step) { CHECK(upload(c->t_arena, &d, tb->h_step, steps_max * bt.B, st)); bt.tab_step = d; }
    if (tb->h_teacher) { CHECK(upload(c->t_arena, &f, tb->h_teacher, steps_max * bt.A * K * 3, st)); bt.tab_teacher = f; }
    if (tb->h_uniforms) { CHECK(upload(c->t_arena, &f, tb->h_uniforms, steps_max * bt.A * K * 64, st)); bt.tab_uniforms = f; }
    CTRM_CUDA_OK(cudaStreamSynchronize(st));
  }
  if (bt.tab_uniforms != nullptr)
    return ctrm_fail(CTRM_ERR_ARG, "per-step uniform tables are only supported through ctrm_cvae_sample (stage entry point)");
  const bool inst_candidate = c->step_mode == 3 || (c->step_mode == 0 && !c->profile && bt.B <= instance_kernel_max_batch(c->device));
  c->t_arena.trim();
  const bool nn = (bt.tab_teacher == nullptr) || c->trace;
  // reset the roadmap store and the roll-out state
  CTRM_CUDA_OK(cudaEventRecord(c->ev_t[0], st));
  const size_t slots = (size_t)bt.A * bt.L * bt.S;
  CTRM_CUDA_OK(cudaMemsetAsync(bt.vcnt, 0, (size_t)bt.A * bt.L * sizeof(int32_t), st));
  CTRM_CUDA_OK(cudaMemsetAsync(bt.cmask, 0, slots * bt.W * sizeof(uint32_t), st));
  CTRM_CUDA_OK(cudaMemsetAsync(bt.pmask, 0, slots * bt.W * sizeof(uint32_t), st));
  if (c->trace) {
    CTRM_CUDA_OK(cudaMemsetAsync(bt.tr_samples, 0xFF, steps_max * bt.A * K * 3 * sizeof(float), st));
    CTRM_CUDA_OK(cudaMemsetAsync(bt.tr_loc_pred, 0xFF, steps_max * bt.A * 2 * sizeof(double), st));
    CTRM_CUDA_OK(cudaMemsetAsync(bt.tr_loc_next, 0xFF, steps_max * bt.A * 2 * sizeof(double), st));
  }
  if (bt.n_learn != nullptr) CTRM_CUDA_OK(cudaMemsetAsync(bt.n_learn, 0, sizeof(int32_t), st));
  if (bt.n_retry != nullptr) CTRM_CUDA_OK(cudaMemsetAsync(bt.n_retry, 0, sizeof(int32_t), st));
  // per-instance kernel (instance.cu): the whole roll-out loop in one launch, one cluster per instance.  Chosen for small
  // batches, where the batch path is one mostly empty tile per kernel and a step costs seven kernel latencies.
  bool use_inst = false;
  int inst_C = 0;
  if (inst_candidate) {
    use_inst = instance_kernel_supports(c->model, bt) != 0;
    if (!use_inst && c->step_mode == 3)
      return ctrm_fail(CTRM_ERR_ARG, "per-instance kernel: batch outside its limits (<= 64 agents per instance, K <= 8, temp 2, "
                                     "<= 15 neighbours, no randomize_indicator)");
    if (use_inst) {
      inst_C = instance_cluster_size();
      bt.learn_agent = nullptr; bt.learn_inst = nullptr; bt.n_learn = nullptr;
      bt.retry_agent = nullptr; bt.retry_inst = nullptr; bt.n_retry = nullptr;
    }
  }
  CHECK(launch_rollout_init(bt, st));
  int64_t launches = 1;
  // capture `chunk` steps as one CUDA graph, then replay it; n_done is polled after every replay.
  // profile mode runs the same kernels eagerly with an event after each one (per-kernel durations).
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(32, steps_max));
  int last_done = 0;
  int per_step = 0;
  int64_t steps = 0;
  int rc = 0;
  cudaError_t ce = cudaSuccess;
  if (use_inst) {
    CTRM_CUDA_OK(cudaEventRecord(c->ev_t[1], st));
    CHECK(launch_instance_rollout(c->model, bt, nn ? 1 : 0, inst_C, st));
    steps = (int64_t)steps_max;
    per_step = 0;
    launches += 1;
  } else if (!c->profile) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    // ONE graph holds `chunk` consecutive steps: a replay per step costs a graph launch (~3 us of host and front-end time),
    // which is most of a step at small batches (one instance: 7 kernels of ~3 us each).  Steps past an instance's last
    // roll-out are no-ops (every row is skipped), so the last chunk may overshoot steps_max.
    CTRM_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < chunk && !rc; ++i) rc = enqueue_step(c, bt, nn, st, &per_step);
    ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return ctrm_fail_cuda(ce, "cudaStreamEndCapture");
    ce = cudaGraphInstantiate(&gexec, graph, 0);
    if (ce != cudaSuccess) { cudaGraphDestroy(graph); return ctrm_fail_cuda(ce, "cudaGraphInstantiate"); }
    CTRM_CUDA_OK(cudaEventRecord(c->ev_t[1], st));
    while ((size_t)steps < steps_max) {
      ce = cudaGraphLaunch(gexec, st);
      if (ce != cudaSuccess) { rc = ctrm_fail_cuda(ce, "cudaGraphLaunch"); break; }
      steps += chunk;
      ce = cudaMemcpyAsync(c->h_ndone, bt.n_done, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = wait_stream(c, bt.A >= 20000);
      if (ce != cudaSuccess) { rc = ctrm_fail_cuda(ce, "roll-out step loop"); break; }
      if (*c->h_ndone >= bt.B) break;
      if (*c->h_ndone > last_done) {  // instances finished during this chunk: drop their rows
        last_done = *c->h_ndone;
        if ((rc = launch_compact_live(bt, st))) break;
        launches += 2;
      }
    }
    cudaGraphExecDestroy(gexec);
    cudaGraphDestroy(graph);
  } else {
    StepProfiler prof;
    CHECK(prof.init(chunk));
    bool ran[KID_STEP_COUNT] = {false};
    CTRM_CUDA_OK(cudaEventRecord(c->ev_t[1], st));
    while ((size_t)steps < steps_max && !rc) {
      int n = (int)std::min<size_t>(chunk, steps_max - steps);
      for (int i = 0; i < n && !rc; ++i) rc = enqueue_step(c, bt, nn, st, &per_step, &prof, i, ran);
      if (rc) break;
      steps += n;
      ce = cudaMemcpyAsync(c->h_ndone, bt.n_done, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
      if (ce != cudaSuccess) { rc = ctrm_fail_cuda(ce, "roll-out step loop (profile)"); break; }
      rc = prof.flush(n, ran);
      if (*c->h_ndone >= bt.B) break;
      if (!rc && *c->h_ndone > last_done) {
        last_done = *c->h_ndone;
        rc = launch_compact_live(bt, st);
        launches += 2;
      }
    }
    for (int k = 0; k < KID_STEP_COUNT; ++k) { c->kernel_ms[k] = prof.ms[k]; c->kernel_n[k] = prof.n[k]; }
    prof.destroy();
  }
  CHECK(rc);
  launches += steps * per_step;
  CTRM_CUDA_OK(cudaEventRecord(c->ev_t[2], st));
  // format_trms + append_goals, then CSR sizes
  CHECK(launch_finalize(bt, c->d_tfin, st));
  CHECK(launch_csr_count(bt, c->d_agent_nv, c->d_agent_ne, c->d_totals, st));
  launches += 6;
  CTRM_CUDA_OK(cudaEventRecord(c->ev_t[3], st));
  CTRM_CUDA_OK(cudaMemcpyAsync(c->totals, c->d_totals, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CTRM_CUDA_OK(wait_stream(c, bt.A >= 20000));
  CTRM_CUDA_OK(cudaEventElapsedTime(&c->ms[0], c->ev_t[0], c->ev_t[3]));
  CTRM_CUDA_OK(cudaEventElapsedTime(&c->ms[1], c->ev_t[1], c->ev_t[2]));
  CTRM_CUDA_OK(cudaEventElapsedTime(&c->ms[2], c->ev_t[2], c->ev_t[3]));
  c->steps = steps;
  c->launches = launches;
  c->have_result = true;
  CHECK(sync_out(c, (cudaStream_t)user_stream));
  return 0;
}

extern "C" int ctrm_result_sizes(ctrm_ctx* c, int64_t* nv, int64_t* ne, int64_t* nl) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  if (!c->have_result) return ctrm_fail(CTRM_ERR_STATE, "no result: call ctrm_build_roadmaps first");
  if (nv) *nv = c->totals[0];
  if (ne) *ne = c->totals[1];
  if (nl) *nl = c->totals[2];
  return 0;
}

extern "C" int ctrm_export_csr(ctrm_ctx* c, int32_t* layer_ptr, double* pos, int64_t* eptr, int32_t* eidx, int32_t* n_layers,
                               int64_t* cnt_collide, int dst_is_device, void* user_stream) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  if (!c->have_result) return ctrm_fail(CTRM_ERR_STATE, "no result: call ctrm_build_roadmaps first");
  CTRM_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  CHECK(sync_in(c, (cudaStream_t)user_stream));
  const Batch& bt = c->bt;
  const int64_t nv = c->totals[0], ne = c->totals[1];
  const size_t nlp = (size_t)bt.A * bt.L + 1;
  const cudaMemcpyKind kind = dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  int32_t *d_lp = layer_ptr, *d_ei = eidx;
  double* d_pos = pos;
  int64_t* d_ep = eptr;
  c->c_arena.release();
  if (!dst_is_device || !layer_ptr) CHECK(c->c_arena.alloc(&d_lp, nlp));
  if (!dst_is_device || !pos) CHECK(c->c_arena.alloc(&d_pos, (size_t)2 * nv));
  if (!dst_is_device || !eptr) CHECK(c->c_arena.alloc(&d_ep, (size_t)nv + 1));
  if (!dst_is_device || !eidx) CHECK(c->c_arena.alloc(&d_ei, (size_t)ne));
  c->c_arena.trim();
  CHECK(launch_csr_fill(bt, c->d_agent_nv, c->d_agent_ne, c->d_totals, d_lp, d_pos, d_ep, d_ei, st));
  if (!dst_is_device) {
    if (layer_ptr) CTRM_CUDA_OK(cudaMemcpyAsync(layer_ptr, d_lp, nlp * sizeof(int32_t), kind, st));
    if (pos) CTRM_CUDA_OK(cudaMemcpyAsync(pos, d_pos, (size_t)2 * nv * sizeof(double), kind, st));
    if (eptr) CTRM_CUDA_OK(cudaMemcpyAsync(eptr, d_ep, ((size_t)nv + 1) * sizeof(int64_t), kind, st));
    if (eidx) CTRM_CUDA_OK(cudaMemcpyAsync(eidx, d_ei, (size_t)ne * sizeof(int32_t), kind, st));
  }
  if (n_layers) CTRM_CUDA_OK(cudaMemcpyAsync(n_layers, bt.nlayers, (size_t)bt.A * sizeof(int32_t), kind, st));
  if (cnt_collide) CTRM_CUDA_OK(cudaMemcpyAsync(cnt_collide, bt.cnt_collide, (size_t)bt.B * sizeof(int64_t), kind, st));
  if (!dst_is_device) CTRM_CUDA_OK(wait_stream(c, bt.A >= 20000));
  CHECK(sync_out(c, (cudaStream_t)user_stream));
  return 0;
}

extern "C" int ctrm_export_csr_compact(ctrm_ctx* c, uint8_t* layer_cnt, double* pos, uint8_t* deg, uint8_t* eidx, int64_t* agent_v0,
                                       int64_t* agent_e0, int32_t* n_layers, int64_t* cnt_collide, int dst_is_device,
                                       void* user_stream) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  if (!c->have_result) return ctrm_fail(CTRM_ERR_STATE, "no result: call ctrm_build_roadmaps first");
  if (!layer_cnt || !pos || !deg || !eidx) return ctrm_fail(CTRM_ERR_ARG, "ctrm_export_csr_compact: NULL output");
  CTRM_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  CHECK(sync_in(c, (cudaStream_t)user_stream));
  const Batch& bt = c->bt;
  const int64_t nv = c->totals[0], ne = c->totals[1];
  const size_t nlc = (size_t)bt.A * bt.L;
  const cudaMemcpyKind kind = dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  uint8_t *d_lc = layer_cnt, *d_dg = deg, *d_ei = eidx;
  double* d_pos = pos;
  if (!dst_is_device) {
    c->c_arena.release();
    CHECK(c->c_arena.alloc(&d_lc, nlc));
    CHECK(c->c_arena.alloc(&d_pos, (size_t)2 * nv));
    CHECK(c->c_arena.alloc(&d_dg, (size_t)nv));
    CHECK(c->c_arena.alloc(&d_ei, (size_t)ne));
    c->c_arena.trim();
  }
  CHECK(launch_csr_fill_compact(bt, c->d_agent_nv, c->d_agent_ne, c->d_totals, d_lc, d_pos, d_dg, d_ei, st));
  if (!dst_is_device) {
    CTRM_CUDA_OK(cudaMemcpyAsync(pos, d_pos, (size_t)2 * nv * sizeof(double), kind, st));
    CTRM_CUDA_OK(cudaMemcpyAsync(eidx, d_ei, (size_t)ne, kind, st));
    CTRM_CUDA_OK(cudaMemcpyAsync(deg, d_dg, (size_t)nv, kind, st));
    CTRM_CUDA_OK(cudaMemcpyAsync(layer_cnt, d_lc, nlc, kind, st));
  }
  if (agent_v0) CTRM_CUDA_OK(cudaMemcpyAsync(agent_v0, c->d_agent_nv, (size_t)bt.A * sizeof(int64_t), kind, st));
  if (agent_e0) CTRM_CUDA_OK(cudaMemcpyAsync(agent_e0, c->d_agent_ne, (size_t)bt.A * sizeof(int64_t), kind, st));
  if (agent_v0) CTRM_CUDA_OK(cudaMemcpyAsync(agent_v0 + bt.A, c->d_totals, sizeof(int64_t), kind, st));
  if (agent_e0) CTRM_CUDA_OK(cudaMemcpyAsync(agent_e0 + bt.A, c->d_totals + 1, sizeof(int64_t), kind, st));
  if (n_layers) CTRM_CUDA_OK(cudaMemcpyAsync(n_layers, bt.nlayers, (size_t)bt.A * sizeof(int32_t), kind, st));
  if (cnt_collide) CTRM_CUDA_OK(cudaMemcpyAsync(cnt_collide, bt.cnt_collide, (size_t)bt.B * sizeof(int64_t), kind, st));
  if (!dst_is_device) CTRM_CUDA_OK(wait_stream(c, bt.A >= 20000));
  CHECK(sync_out(c, (cudaStream_t)user_stream));
  return 0;
}

extern "C" int ctrm_trace_read(ctrm_ctx* c, float* h_samples, double* h_loc_pred, double* h_loc_next) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  if (!c->have_result || !c->r_trace || !c->bt.tr_samples) return ctrm_fail(CTRM_ERR_STATE, "no trace recorded");
  const Batch& bt = c->bt;
  const size_t steps = (size_t)bt.N_traj * bt.max_T;
  CTRM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if (h_samples) CTRM_CUDA_OK(cudaMemcpy(h_samples, bt.tr_samples, steps * bt.A * bt.K * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  if (h_loc_pred) CTRM_CUDA_OK(cudaMemcpy(h_loc_pred, bt.tr_loc_pred, steps * bt.A * 2 * sizeof(double), cudaMemcpyDeviceToHost));
  if (h_loc_next) CTRM_CUDA_OK(cudaMemcpy(h_loc_next, bt.tr_loc_next, steps * bt.A * 2 * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ctrm_get_instance_steps(ctrm_ctx* c, int32_t* h_steps) {
  if (!c || !h_steps) return ctrm_fail(CTRM_ERR_ARG, "NULL argument");
  if (!c->have_result) return ctrm_fail(CTRM_ERR_STATE, "no result");
  CTRM_CUDA_OK(cudaStreamSynchronize(c->stream));
  CTRM_CUDA_OK(cudaMemcpy(h_steps, c->bt.steps_done, (size_t)c->bt.B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ctrm_set_step_kernel(ctrm_ctx* c, int mode) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  if (mode < 0 || mode > 3)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_set_step_kernel: mode must be 0 (auto), 1 (warp per agent), 2 (thread per agent) or 3 (per-instance kernel)");
  c->step_mode = mode;
  return 0;
}

extern "C" int ctrm_set_profile(ctrm_ctx* c, int enable) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  c->profile = enable != 0;
  return 0;
}

extern "C" int ctrm_last_kernel_times(ctrm_ctx* c, double ms[8], int64_t launches[8]) {
  if (!c || !ms || !launches) return ctrm_fail(CTRM_ERR_ARG, "NULL argument");
  for (int k = 0; k < 8; ++k) { ms[k] = k < KID_STEP_COUNT ? c->kernel_ms[k] : 0.0; launches[k] = k < KID_STEP_COUNT ? c->kernel_n[k] : 0; }
  return 0;
}

extern "C" int ctrm_last_timing(ctrm_ctx* c, float ms[3], int64_t* steps, int64_t* launches) {
  if (!c) return ctrm_fail(CTRM_ERR_ARG, "NULL ctx");
  if (ms) { ms[0] = c->ms[0]; ms[1] = c->ms[1]; ms[2] = c->ms[2]; }
  if (steps) *steps = c->steps;
  if (launches) *launches = c->launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// host-buffer wrappers with the pybind11 helpers' call shapes (B = 1); GPU only
// ------------------------------------------------------------------------------------------------
// training-time features
// ------------------------------------------------------------------------------------------------
extern "C" int ctrm_others_info(const double* d_cur, const double* d_goals, const double* d_prev, const double* d_rads,
                                const double* d_speeds, int T, int N, float* d_info, float* d_self_info, void* stream) {
  CHECK(use_device_of(d_info));
  if (T <= 0 || N <= 0 || !d_cur || !d_goals || !d_prev || !d_rads || !d_speeds || !d_info)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_others_info: bad argument");
  return launch_others_info(d_cur, d_goals, d_prev, d_rads, d_speeds, T, N, d_info, d_self_info, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// baseline samplers
// ------------------------------------------------------------------------------------------------
extern "C" int ctrm_collide_points(const double* d_pos, const double* d_rad, const int32_t* d_pt_inst, int64_t n,
                                   const double* d_obs_xyr, const int32_t* d_obs_off, uint8_t* d_verdict, void* stream) {
  CHECK(use_device_of(d_verdict));
  if (n < 0 || (n > 0 && (!d_pos || !d_rad || !d_obs_off || !d_verdict)))
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_collide_points: NULL argument");
  return launch_collide_points(d_pos, d_rad, d_pt_inst, n, d_obs_xyr, d_obs_off, d_verdict, (cudaStream_t)stream);
}

// layout of the bit rows / work items of n_sets (rows x columns) blocks; returns the number of 32-bit words
static int64_t adjacency_layout(const int32_t* h_row_off, const int32_t* h_col_off, int n_sets, std::vector<int64_t>* bits_off,
                                std::vector<int64_t>* item_off) {
  int64_t words = 0, items = 0;
  if (bits_off) bits_off->assign((size_t)n_sets + 1, 0);
  if (item_off) item_off->assign((size_t)n_sets + 1, 0);
  for (int s = 0; s < n_sets; ++s) {
    const int64_t nr = h_row_off[s + 1] - h_row_off[s], nc = h_col_off[s + 1] - h_col_off[s];
    if (nr < 0 || nc < 0) return -1;
    const int64_t W = (nc + 31) / 32;
    words += nr * W;
    items += nr * ((W + 7) / 8);  // PA_WORDS = 8 words per warp item (sampler.cu)
    if (bits_off) (*bits_off)[s + 1] = words;
    if (item_off) (*item_off)[s + 1] = items;
  }
  return words;
}

extern "C" int64_t ctrm_pair_adjacency_words(const int32_t* h_row_off, const int32_t* h_col_off, int n_sets) {
  if (!h_row_off || !h_col_off || n_sets < 0) return -1;
  return adjacency_layout(h_row_off, h_col_off, n_sets, nullptr, nullptr);
}

// small per-set host arrays -> stream-ordered device scratch
template <typename T>
static int scratch_upload(T** d, const T* h, size_t n, cudaStream_t st) {
  *d = nullptr;
  CTRM_CUDA_OK(cudaMallocAsync(d, sizeof(T) * (n ? n : 1), st));
  if (n) CTRM_CUDA_OK(cudaMemcpyAsync(*d, h, sizeof(T) * n, cudaMemcpyHostToDevice, st));
  return 0;
}

extern "C" int ctrm_pair_adjacency(const double* d_rows, const int32_t* h_row_off, const double* d_cols, const int32_t* h_col_off,
                                   int n_sets, const int32_t* h_set_inst, const double* h_speed, const double* h_speed2,
                                   const double* h_rad, const double* d_obs_xyr, const int32_t* d_obs_off, int symmetric,
                                   uint32_t* d_bits, int64_t* d_cnt, void* stream) {
  CHECK(use_device_of(d_bits));
  if (n_sets <= 0 || !d_rows || !h_row_off || !d_cols || !h_col_off || !h_speed || !h_speed2 || !h_rad || !d_obs_off || !d_bits)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_pair_adjacency: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int64_t> bits_off, item_off;
  if (adjacency_layout(h_row_off, h_col_off, n_sets, &bits_off, &item_off) < 0)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_pair_adjacency: offsets must be non-decreasing");
  if (symmetric)
    for (int s = 0; s < n_sets; ++s)
      if (h_row_off[s + 1] - h_row_off[s] != h_col_off[s + 1] - h_col_off[s])
        return ctrm_fail(CTRM_ERR_ARG, "ctrm_pair_adjacency: symmetric needs rows == columns");
  CHECK(keep_async_pool_warm());
  int32_t *d_ro = nullptr, *d_co = nullptr, *d_si = nullptr;
  double *d_sp = nullptr, *d_sp2 = nullptr, *d_ra = nullptr;
  int64_t *d_bo = nullptr, *d_io = nullptr;
  int rc = scratch_upload(&d_ro, h_row_off, (size_t)n_sets + 1, st);
  if (!rc) rc = scratch_upload(&d_co, h_col_off, (size_t)n_sets + 1, st);
  if (!rc && h_set_inst) rc = scratch_upload(&d_si, h_set_inst, (size_t)n_sets, st);
  if (!rc) rc = scratch_upload(&d_sp, h_speed, (size_t)n_sets, st);
  if (!rc) rc = scratch_upload(&d_sp2, h_speed2, (size_t)n_sets, st);
  if (!rc) rc = scratch_upload(&d_ra, h_rad, (size_t)n_sets, st);
  if (!rc) rc = scratch_upload(&d_bo, bits_off.data(), bits_off.size(), st);
  if (!rc) rc = scratch_upload(&d_io, item_off.data(), item_off.size(), st);
  if (!rc && d_cnt) {
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, sizeof(int64_t) * (size_t)n_sets, st);
    if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "memset");
  }
  if (!rc)
    rc = launch_pair_adjacency(d_rows, d_ro, d_cols, d_co, n_sets, d_si, d_sp, d_sp2, d_ra, d_obs_xyr, d_obs_off, symmetric, d_bo, d_io,
                               item_off[n_sets], d_bits, d_cnt, st);
  for (void* p : {(void*)d_ro, (void*)d_co, (void*)d_si, (void*)d_sp, (void*)d_sp2, (void*)d_ra, (void*)d_bo, (void*)d_io})
    if (p) cudaFreeAsync(p, st);
  return rc;
}

extern "C" int ctrm_adjacency_csr(const uint32_t* d_bits, const int32_t* h_row_off, const int32_t* h_col_off, int n_sets,
                                  int self_first, int64_t* d_eptr, int32_t* d_eidx, void* stream) {
  CHECK(use_device_of(d_eptr));
  if (n_sets <= 0 || !d_bits || !h_row_off || !h_col_off || !d_eptr) return ctrm_fail(CTRM_ERR_ARG, "ctrm_adjacency_csr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int64_t> bits_off;
  if (adjacency_layout(h_row_off, h_col_off, n_sets, &bits_off, nullptr) < 0)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_adjacency_csr: offsets must be non-decreasing");
  CHECK(keep_async_pool_warm());
  int32_t *d_ro = nullptr, *d_co = nullptr;
  int64_t* d_bo = nullptr;
  int rc = scratch_upload(&d_ro, h_row_off, (size_t)n_sets + 1, st);
  if (!rc) rc = scratch_upload(&d_co, h_col_off, (size_t)n_sets + 1, st);
  if (!rc) rc = scratch_upload(&d_bo, bits_off.data(), bits_off.size(), st);
  if (!rc) rc = launch_adjacency_csr(d_bits, d_bo, d_ro, d_co, n_sets, h_row_off[n_sets] - h_row_off[0], self_first, d_eptr, d_eidx, st);
  for (void* p : {(void*)d_ro, (void*)d_co, (void*)d_bo})
    if (p) cudaFreeAsync(p, st);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// planner
// ------------------------------------------------------------------------------------------------
extern "C" int ctrm_plan_prioritized(int B, int A, int L, const int32_t* d_agent_off, const int32_t* d_n_layers,
                                     const int32_t* d_layer_ptr, const double* d_pos, const int64_t* d_eptr,
                                     const int32_t* d_eidx, const double* d_goals, const double* d_speeds, const double* d_rads,
                                     int max_agent_vertices, int32_t* d_paths, double* d_path_pos, int32_t* d_solved,
                                     int32_t* d_failed_agent, int64_t* d_counters, void* stream) {
  CHECK(use_device_of(d_paths));
  if (B <= 0 || A <= 0 || L < 2 || !d_agent_off || !d_n_layers || !d_layer_ptr || !d_pos || !d_eptr || !d_eidx || !d_goals ||
      !d_speeds || !d_rads || max_agent_vertices < 1 || !d_paths || !d_path_pos || !d_solved || !d_failed_agent || !d_counters)
    return ctrm_fail(CTRM_ERR_ARG, "ctrm_plan_prioritized: bad argument");
  return launch_plan_prioritized(B, L, d_agent_off, d_n_layers, d_layer_ptr, d_pos, d_eptr, d_eidx, d_goals, d_speeds, d_rads,
                                 max_agent_vertices, A, d_paths, d_solved, d_failed_agent,
                                 reinterpret_cast<long long*>(d_counters), d_path_pos, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
extern "C" int ctrm_host_continuous_collide_spheres(const double* from1, const double* to1, double rad1, const double* from2,
                                                    const double* to2, double rad2, int dim, int* out) {
  if (!from1 || !to1 || !from2 || !to2 || !out || (dim != 2 && dim != 3)) return ctrm_fail(CTRM_ERR_ARG, "bad argument");
  double h[14];
  for (int i = 0; i < dim; ++i) { h[i] = from1[i]; h[3 + i] = to1[i]; h[6 + i] = from2[i]; h[9 + i] = to2[i]; }
  h[12] = rad1; h[13] = rad2;
  double* d = nullptr;
  uint8_t* dv = nullptr;
  CTRM_CUDA_OK(cudaMalloc(&d, sizeof(h)));
  CTRM_CUDA_OK(cudaMalloc(&dv, 16));
  int rc = 0;
  cudaError_t e = cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "memcpy");
  // the [n][dim] layout with n = 1: from1 at d, to1 at d+3, from2 at d+6, to2 at d+9
  if (!rc) rc = launch_collide_pairs(d, d + 3, d + 12, d + 6, d + 9, d + 13, dim, 1, dv, nullptr);
  uint8_t v = 0;
  if (!rc) {
    e = cudaMemcpy(&v, dv, 1, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "memcpy");
  }
  cudaFree(d);
  cudaFree(dv);
  *out = v ? 1 : 0;
  return rc;
}

extern "C" int ctrm_host_cost_to_go_2d(const double* goal, const int32_t* occ, int M, const int32_t* agent, int agent_size,
                                       int32_t* out) {
  if (!goal || !occ || !agent || !out || M < 1 || M > 255 || agent_size < 1 || (agent_size & 1) == 0 || agent_size > 255)
    return ctrm_fail(CTRM_ERR_ARG, "bad argument");
  const int r = (agent_size - 1) / 2;
  std::vector<uint8_t> occ8((size_t)M * M), st8((size_t)agent_size * agent_size);
  for (size_t i = 0; i < occ8.size(); ++i) occ8[i] = occ[i] == 1 ? 1 : 0;  // the reference tests `== 1` (cost_to_go.cpp:68)
  for (size_t i = 0; i < st8.size(); ++i) st8[i] = agent[i] == 1 ? 1 : 0;
  DevArena ar;
  uint8_t* d_occ = nullptr;
  double* d_goal = nullptr;
  uint16_t* d_ctg = nullptr;
  int rc = upload(ar, &d_occ, occ8.data(), occ8.size(), (cudaStream_t) nullptr);
  if (!rc) rc = upload(ar, &d_goal, goal, 2, (cudaStream_t) nullptr);
  if (!rc) rc = ar.alloc(&d_ctg, (size_t)M * M);
  std::vector<int32_t> st_r{r}, pm_inst{0}, pm_st{0}, agent_pm{0};
  if (!rc) rc = cost_to_go_device(d_occ, M, st8, st_r, agent_size, pm_inst, pm_st, agent_pm, d_goal, 1, d_ctg, nullptr);
  std::vector<uint16_t> h((size_t)M * M);
  if (!rc) {
    cudaError_t e = cudaMemcpy(h.data(), d_ctg, h.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "memcpy");
  }
  ar.destroy();
  if (!rc)
    for (size_t i = 0; i < h.size(); ++i) out[i] = (int32_t)h[i];
  return rc;
}

extern "C" int ctrm_host_fov2d(const double* pos, const int32_t* occ, const int32_t* ctg, int M, int F, int32_t* out) {
  if (!pos || !occ || !ctg || !out || M < 1 || M > 255 || F < 1 || (F & 1) == 0 || F > 63) return ctrm_fail(CTRM_ERR_ARG, "bad argument");
  std::vector<uint8_t> occ8((size_t)M * M);
  std::vector<uint16_t> ctg16((size_t)M * M);
  for (size_t i = 0; i < occ8.size(); ++i) {
    occ8[i] = (uint8_t)(occ[i] & 1);
    // the reference sees inf as INT_MIN: `_c >= 0` is its reachability test (cost_to_go.cpp:176-179)
    ctg16[i] = (ctg[i] < 0 || ctg[i] >= 65535) ? (uint16_t)65535 : (uint16_t)ctg[i];
  }
  DevArena ar;
  uint8_t* d_occ = nullptr;
  uint16_t* d_ctg = nullptr;
  double* d_pos = nullptr;
  int32_t* d_inst = nullptr;
  float* d_fov = nullptr;
  const int32_t zero = 0;
  const size_t n = (size_t)2 * F * F;
  int rc = upload(ar, &d_occ, occ8.data(), occ8.size(), (cudaStream_t) nullptr);
  if (!rc) rc = upload(ar, &d_ctg, ctg16.data(), ctg16.size(), (cudaStream_t) nullptr);
  if (!rc) rc = upload(ar, &d_pos, pos, 2, (cudaStream_t) nullptr);
  if (!rc) rc = upload(ar, &d_inst, &zero, 1, (cudaStream_t) nullptr);
  if (!rc) rc = ar.alloc(&d_fov, n + 8);
  if (!rc) rc = launch_fov_raster(d_occ, d_ctg, d_inst, d_pos, nullptr, 1, M, F, 0, d_fov, nullptr);
  std::vector<float> h(n);
  if (!rc) {
    cudaError_t e = cudaMemcpy(h.data(), d_fov, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = ctrm_fail_cuda(e, "memcpy");
  }
  ar.destroy();
  if (!rc)
    for (size_t i = 0; i < n; ++i) out[i] = (int32_t)h[i];
  return rc;
}
