"""Timed-roadmap construction with the learned sampler — the public API of the path.

  get_timed_roadmaps_multiple_paths_with_learned_indicator(ins, pred_basename, ...) -> list[TimedRoadmap]
      drop-in for ctrm.roadmap_learned (roadmap_learned/learned_sampler.py:19-204), same keyword arguments and
      defaults.  Behind it: RoadmapBuilder -> libctrm_b200 (CUDA).  No CPU fallback.
  RoadmapBuilder.build(instances, ...) -> CSRRoadmaps
      the batched form: many independent instances per call (the reference parallelises over instances with joblib,
      scripts/eval.py:108-111; here they share every kernel launch).

Randomness: the reference consumes numpy's / torch's global streams in a data-dependent order, which a GPU cannot
reproduce; every draw is keyed instead (Philox4x32-10 over (seed; instance, k_traj, t, agent, attempt) — see
csrc/common.cuh).  `seed=None` takes one draw from numpy's global generator so that `set_global_seeds` still makes
runs repeatable.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .environment import Instance
from .timed_roadmap import CompactCSRRoadmaps, CSRRoadmaps
from .weights import load_bundle, pack_blob

_BUILDERS: dict = {}


def prob_random_walk_table(max_T: int, gamma: float, bias: float) -> np.ndarray:
    """p_rw[t][D] = (1 - exp(-gamma * t / D)) * (1 - bias) + bias  (learned_sampler.py:84-91), D = makespan or max_T.
    Evaluated scalar by scalar with numpy exactly like the reference so the doubles are identical."""
    tab = np.zeros((max_T + 1, max_T + 1), dtype=np.float64)
    for t in range(1, max_T + 1):
        for D in range(1, max_T + 1):
            tab[t, D] = (1 - np.exp(-gamma * t / D)) * (1 - bias) + bias
    return tab


class RoadmapBuilder:
    """owns one libctrm_b200 context (one GPU) with the networks resident"""

    def __init__(self, pred_basename: str, device: int = 0, map_size: Optional[int] = None):
        self.lib = _lib.load()
        self.bundle = load_bundle(pred_basename)
        if map_size is not None:  # the networks only see F x F crops, so the grid resolution M is free
            self.bundle.map_size = int(map_size)
        self.desc, self.off, self.blob = pack_blob(self.bundle)
        self.ctx = C.c_void_p()
        _lib.check(self.lib.ctrm_ctx_create(int(device), C.byref(self.ctx)), "ctrm_ctx_create")
        _lib.check(self.lib.ctrm_load_weights(self.ctx, C.byref(self.desc), _lib.ptr(self.blob), self.blob.size),
                   "ctrm_load_weights")
        self.device = int(device)
        self._uploaded = None
        self._prw_cache = {}
        self._pinned = {}  # cached pinned host buffers of export(); a CSRRoadmaps returned with pinned=True is a view into
        #                    them and stays valid until the next export() of this builder

    def close(self):
        if self.ctx:
            self.lib.ctrm_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- instances -------------------------------------------------------------------------------------------
    @staticmethod
    def pack_instances(instances: Sequence[Instance]):
        """SoA host arrays (pinned when CUDA is available) in the layout of ctrm_upload_instances"""
        parts = [ins.arrays() for ins in instances]
        n_agents = np.array([p[0].shape[0] for p in parts], dtype=np.int32)
        n_obs = np.array([p[4].shape[0] for p in parts], dtype=np.int32)
        starts = np.ascontiguousarray(np.concatenate([p[0] for p in parts]))
        goals = np.ascontiguousarray(np.concatenate([p[1] for p in parts]))
        speeds = np.ascontiguousarray(np.concatenate([p[2] for p in parts]))
        rads = np.ascontiguousarray(np.concatenate([p[3] for p in parts]))
        obs = np.ascontiguousarray(np.concatenate([p[4] for p in parts]).reshape(-1, 3))
        # max_speed ** 2 in Python-float arithmetic, as roadmap/utils.py:40 and roadmap_learned/utils.py:60 do: CPython's
        # float ** 2 goes through libm pow, which is NOT always the correctly rounded square numpy's s * s gives
        speeds2 = np.concatenate([ins.speeds2() for ins in instances]) if all(hasattr(i, "speeds2") for i in instances) else \
            np.array([float(s) ** 2 for s in speeds], dtype=np.float64)
        return dict(n_agents=n_agents, n_obs=n_obs, starts=starts, goals=goals, speeds=speeds, speeds2=speeds2, rads=rads,
                    obs=obs)

    def upload(self, instances: Sequence[Instance], packed: Optional[dict] = None, stream: int = 0):
        """Format2D_CTRM_Input.set_instance for the whole batch: H2D copy, occupancy raster, cost-to-go wavefront"""
        p = packed if packed is not None else self.pack_instances(instances)
        B = len(p["n_agents"])
        _lib.check(self.lib.ctrm_upload_instances(self.ctx, B, _lib.ptr(p["n_agents"]), _lib.ptr(p["starts"]),
                                                  _lib.ptr(p["goals"]), _lib.ptr(p["speeds"]), _lib.ptr(p["speeds2"]),
                                                  _lib.ptr(p["rads"]), _lib.ptr(p["n_obs"]),
                                                  _lib.ptr(p["obs"]) if p["obs"].size else None, stream),
                   "ctrm_upload_instances")
        self._uploaded = p
        return p

    def device_view(self) -> _lib.DeviceView:
        v = _lib.DeviceView()
        _lib.check(self.lib.ctrm_get_device_view(self.ctx, C.byref(v)), "ctrm_get_device_view")
        return v

    # ---- roll-outs -------------------------------------------------------------------------------------------
    def run(self, prob_uniform_sampling_after_goal: float = 0.9, prob_uniform_bias: float = 0.0,
            prob_uniform_gamma: float = 5.0, max_T: int = 64, N_traj: int = 25, max_attempt: int = 3,
            randomize_indicator: bool = False, merge_distance_rate: float = 0.25, K: int = 0, seed: Optional[int] = None,
            instance_id_base: int = 0, tables: Optional[dict] = None, trace: bool = False, stream: int = 0,
            export: bool = True, pinned: bool = True, compact: bool = True) -> Optional[CSRRoadmaps]:
        """roll-outs + finalisation + CSR export for the uploaded batch"""
        if self._uploaded is None:
            raise _lib.CtrmError("upload instances first")
        p = self._uploaded
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        tables = tables or {}
        prm = _lib.Params(float(prob_uniform_sampling_after_goal), float(prob_uniform_bias), float(prob_uniform_gamma),
                          int(max_T), int(N_traj), int(max_attempt), int(bool(randomize_indicator)),
                          float(merge_distance_rate), int(K), 1 if "gate" in tables else 0, int(seed) & (2 ** 64 - 1),
                          int(instance_id_base))
        key = (int(max_T), float(prob_uniform_gamma), float(prob_uniform_bias))
        if key not in self._prw_cache:
            self._prw_cache[key] = prob_random_walk_table(*key)
        prw = self._prw_cache[key]
        # merge_distance = rate * max_speed ; merge_distance ** 2  (learned_sampler.py:177, utils.py:87)
        merge2 = np.array([(merge_distance_rate * float(s)) ** 2 for s in p["speeds"]], dtype=np.float64)
        keep = []

        def tab(name, dtype):
            if name not in tables or tables[name] is None:
                return None
            arr = np.ascontiguousarray(tables[name], dtype=dtype)
            keep.append(arr)
            return _lib.ptr(arr)

        tb = _lib.Tables(tab("gate", np.float64), tab("rw", np.float64), tab("step", np.float64), tab("teacher", np.float32),
                         tab("uniforms", np.float32), _lib.ptr(prw), _lib.ptr(merge2))
        _lib.check(self.lib.ctrm_trace_enable(self.ctx, int(trace)), "ctrm_trace_enable")
        _lib.check(self.lib.ctrm_build_roadmaps(self.ctx, C.byref(prm), C.byref(tb), stream), "ctrm_build_roadmaps")
        self._last = dict(max_T=int(max_T), N_traj=int(N_traj), K=int(K) if K else int(self.bundle.dim_ind))
        if not export:
            return None
        return self.export(stream=stream, pinned=pinned, compact=compact)

    def result_sizes(self):
        nv, ne, nl = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.ctrm_result_sizes(self.ctx, C.byref(nv), C.byref(ne), C.byref(nl)), "ctrm_result_sizes")
        return nv.value, ne.value, nl.value

    def timing(self):
        ms = (C.c_float * 3)()
        steps, launches = C.c_int64(), C.c_int64()
        _lib.check(self.lib.ctrm_last_timing(self.ctx, C.cast(ms, C.c_void_p), C.byref(steps), C.byref(launches)),
                   "ctrm_last_timing")
        return dict(total_ms=ms[0], rollout_ms=ms[1], finalize_ms=ms[2], steps=steps.value, launches=launches.value)

    def export(self, stream: int = 0, pinned: bool = True, compact: bool = True) -> CSRRoadmaps:
        """D2H copy of the CSR result.  compact (default): byte-sized layer counts / degrees / successor indices
        (ctrm_export_csr_compact, half the bytes over PCIe) wrapped in a CompactCSRRoadmaps, which expands instances on
        demand; compact=False: the wide arrays of ctrm_export_csr"""
        p = self._uploaded
        nv, ne, _ = self.result_sizes()
        A = int(p["n_agents"].sum())
        B = len(p["n_agents"])
        L = self._last["max_T"] + 2

        def host(name, shape, dtype):
            """pinned staging buffers are cached and grown geometrically: cudaHostAlloc costs ~0.3 s per GB"""
            if not pinned:
                return np.empty(shape, dtype=dtype)
            import torch

            n = int(np.prod(shape))
            tdt = {np.int32: torch.int32, np.int64: torch.int64, np.float64: torch.float64, np.uint8: torch.uint8}[dtype]
            buf = self._pinned.get(name)
            if buf is None or buf.numel() < n or buf.dtype != tdt:
                buf = torch.empty(max(n + n // 8, 1), dtype=tdt, pin_memory=True)
                self._pinned[name] = buf
            return buf[:n].numpy().reshape(shape)

        agent_off = np.concatenate([[0], np.cumsum(p["n_agents"])]).astype(np.int32)
        if compact:
            layer_cnt = host("layer_cnt", A * L, np.uint8)
            pos = host("pos", (nv, 2), np.float64)
            deg = host("deg", max(nv, 1), np.uint8)[:nv]
            eidx8 = host("eidx8", max(ne, 1), np.uint8)[:ne]
            av0, ae0 = host("agent_v0", A + 1, np.int64), host("agent_e0", A + 1, np.int64)
            n_layers = host("n_layers", A, np.int32)
            cnt = host("cnt", B, np.int64)
            _lib.check(self.lib.ctrm_export_csr_compact(self.ctx, _lib.ptr(layer_cnt), _lib.ptr(pos),
                                                        deg.ctypes.data_as(C.c_void_p), eidx8.ctypes.data_as(C.c_void_p),
                                                        _lib.ptr(av0), _lib.ptr(ae0), _lib.ptr(n_layers), _lib.ptr(cnt), 0, stream),
                       "ctrm_export_csr_compact")
            return CompactCSRRoadmaps(agent_off, L, layer_cnt, pos, deg, eidx8, av0, ae0, n_layers, cnt, self.timing())
        layer_ptr = host("layer_ptr", A * L + 1, np.int32)
        pos = host("pos", (nv, 2), np.float64)
        eptr = host("eptr", nv + 1, np.int64)
        eidx = host("eidx", max(ne, 1), np.int32)[:ne]
        n_layers = host("n_layers", A, np.int32)
        cnt = host("cnt", B, np.int64)
        _lib.check(self.lib.ctrm_export_csr(self.ctx, _lib.ptr(layer_ptr), _lib.ptr(pos), _lib.ptr(eptr),
                                            eidx.ctypes.data_as(C.c_void_p), _lib.ptr(n_layers), _lib.ptr(cnt), 0, stream),
                   "ctrm_export_csr")
        return CSRRoadmaps(agent_off, L, layer_ptr, pos, eptr, eidx, n_layers, cnt, self.timing())

    KERNELS = ("fov_raster", "fov_encode", "agent_comm", "cvae_prior", "cvae_decode", "step")

    def set_step_kernel(self, mode: int):
        """0 = by batch size, 1 = one warp per agent, 2 = one thread per agent (ctrm_set_step_kernel); same results"""
        _lib.check(self.lib.ctrm_set_step_kernel(self.ctx, int(mode)), "ctrm_set_step_kernel")

    def set_profile(self, enable: bool):
        """eager launches with a CUDA event after every kernel (per-kernel device times, see kernel_times)"""
        _lib.check(self.lib.ctrm_set_profile(self.ctx, int(enable)), "ctrm_set_profile")

    def kernel_times(self) -> dict:
        ms = (C.c_double * 8)()
        n = (C.c_int64 * 8)()
        _lib.check(self.lib.ctrm_last_kernel_times(self.ctx, C.cast(ms, C.c_void_p), C.cast(n, C.c_void_p)),
                   "ctrm_last_kernel_times")
        return {k: dict(ms=ms[i], launches=n[i]) for i, k in enumerate(self.KERNELS)}

    def instance_steps(self) -> np.ndarray:
        out = np.zeros(len(self._uploaded["n_agents"]), dtype=np.int32)
        _lib.check(self.lib.ctrm_get_instance_steps(self.ctx, _lib.ptr(out)), "ctrm_get_instance_steps")
        return out

    def export_device(self, stream: int = 0, compact: bool = False) -> dict:
        """CSR result as torch CUDA tensors (no D2H): the wide arrays the planner kernel reads, or (compact=True) the
        byte-sized form of ctrm_export_csr_compact that the multi-GPU gather ships over NVLink"""
        import torch

        p = self._uploaded
        nv, ne, _ = self.result_sizes()
        A = int(p["n_agents"].sum())
        B = len(p["n_agents"])
        L = self._last["max_T"] + 2
        dev = torch.device("cuda", self.device)
        if compact:
            out = dict(layer_cnt=torch.empty(A * L, dtype=torch.uint8, device=dev),
                       pos=torch.empty((nv, 2), dtype=torch.float64, device=dev),
                       deg=torch.empty(max(nv, 1), dtype=torch.uint8, device=dev),
                       eidx8=torch.empty(max(ne, 1), dtype=torch.uint8, device=dev),
                       agent_v0=torch.empty(A + 1, dtype=torch.int64, device=dev),
                       agent_e0=torch.empty(A + 1, dtype=torch.int64, device=dev),
                       n_layers=torch.empty(A, dtype=torch.int32, device=dev),
                       cnt_collide=torch.empty(B, dtype=torch.int64, device=dev))
            _lib.check(self.lib.ctrm_export_csr_compact(self.ctx, out["layer_cnt"].data_ptr(), out["pos"].data_ptr(),
                                                        out["deg"].data_ptr(), out["eidx8"].data_ptr(), out["agent_v0"].data_ptr(),
                                                        out["agent_e0"].data_ptr(), out["n_layers"].data_ptr(),
                                                        out["cnt_collide"].data_ptr(), 1, stream), "ctrm_export_csr_compact")
            out["deg"], out["eidx8"] = out["deg"][:nv], out["eidx8"][:ne]
            return out
        out = dict(layer_ptr=torch.empty(A * L + 1, dtype=torch.int32, device=dev),
                   pos=torch.empty((nv, 2), dtype=torch.float64, device=dev),
                   eptr=torch.empty(nv + 1, dtype=torch.int64, device=dev),
                   eidx=torch.empty(max(ne, 1), dtype=torch.int32, device=dev),
                   n_layers=torch.empty(A, dtype=torch.int32, device=dev),
                   cnt_collide=torch.empty(B, dtype=torch.int64, device=dev))
        _lib.check(self.lib.ctrm_export_csr(self.ctx, out["layer_ptr"].data_ptr(), out["pos"].data_ptr(), out["eptr"].data_ptr(),
                                            out["eidx"].data_ptr(), out["n_layers"].data_ptr(), out["cnt_collide"].data_ptr(), 1,
                                            stream), "ctrm_export_csr")
        out["eidx"] = out["eidx"][:ne]
        return out

    def plan(self, csr_device: Optional[dict] = None) -> dict:
        """prioritized planning (planner/prioritized_planning.py:40-101) for every instance of the batch just built, on the
        device-resident CSR roadmaps: no roadmap ever visits the host.  Returns CUDA tensors: paths [A][L] (vertex index per
        time step), path_pos [A][L][2], solved [B], failed_agent [B], counters [B][3], and elapsed_ms."""
        import torch

        from .planner import plan_prioritized_device

        p = self._uploaded
        d = csr_device if csr_device is not None else self.export_device()
        dev = d["pos"].device
        A, B, L = int(p["n_agents"].sum()), len(p["n_agents"]), self._last["max_T"] + 2
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
        agent_off = up(np.concatenate([[0], np.cumsum(p["n_agents"])]), np.int32)
        return plan_prioritized_device(B, A, L, agent_off, d["n_layers"], d["layer_ptr"], d["pos"], d["eptr"], d["eidx"],
                                       up(p["goals"], np.float64), up(p["speeds"], np.float64), up(p["rads"], np.float64))

    def read_trace(self):
        """(samples [N_traj][max_T][A][K][3] f32, loc_pred, loc_next [N_traj][max_T][A][2] f64); NaN = not executed"""
        p = self._uploaded
        A = int(p["n_agents"].sum())
        NT, MT, K = self._last["N_traj"], self._last["max_T"], self._last["K"]
        s = np.empty((NT, MT, A, K, 3), dtype=np.float32)
        lp = np.empty((NT, MT, A, 2), dtype=np.float64)
        ln = np.empty((NT, MT, A, 2), dtype=np.float64)
        _lib.check(self.lib.ctrm_trace_read(self.ctx, _lib.ptr(s), _lib.ptr(lp), _lib.ptr(ln)), "ctrm_trace_read")
        return s, lp, ln

    def build(self, instances: Sequence[Instance], **kw) -> CSRRoadmaps:
        import time

        t0 = time.perf_counter()
        packed = self.pack_instances(instances)
        t1 = time.perf_counter()
        self.upload(None, packed=packed)
        t2 = time.perf_counter()
        export = kw.pop("export", True)
        self.run(export=False, **{k: v for k, v in kw.items() if k not in ("pinned", "compact")})
        t3 = time.perf_counter()
        out = self.export(stream=kw.get("stream", 0), pinned=kw.get("pinned", True), compact=kw.get("compact", True)) if export else None
        # host wall time of the four phases of the last build (ms): pack, upload (H2D + occupancy + wavefront), roll-outs, export
        self.last_phases_ms = tuple(1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, time.perf_counter())))
        return out


class PipelinedBuilder:
    """Roadmaps for a STREAM of instance batches (what scripts/eval.py's joblib loop over instances is, `eval.py:108-111`).

    Two RoadmapBuilder contexts on one GPU take the batches alternately, each from its own host thread (the C-ABI calls
    release the GIL), so the host packing, the host->device upload and the device->host export of one batch overlap the
    roll-out kernels of the other; the kernels of the two contexts share the GPU without gain or loss.  Results come back
    in order.  A yielded CSRRoadmaps is a view of its context's pinned staging buffers and is valid ONLY UNTIL THE NEXT
    ITEM IS REQUESTED from the generator: asking for the next result hands the context that produced this one its next
    batch, whose export overwrites the same buffers.  `.copy()` a result to keep it longer."""

    def __init__(self, pred_basename: str, device: int = 0, map_size: Optional[int] = None, depth: int = 2):
        self.builders = [RoadmapBuilder(pred_basename, device, map_size=map_size) for _ in range(max(1, int(depth)))]
        # stagger the contexts: with equal priorities their kernels time-share the GPU evenly, both batches finish together
        # and the GPU idles while both export / pack / upload; the first context's kernels go first instead
        for i, b in enumerate(self.builders):
            _lib.check(b.lib.ctrm_set_stream_priority(b.ctx, 1 if i == 0 else 0), "ctrm_set_stream_priority")

    def close(self):
        for b in self.builders:
            b.close()

    def map(self, batches, **kw):
        """yield one CSRRoadmaps per batch of `batches` (an iterable of sequences of Instance), in order; `kw` as
        RoadmapBuilder.run.  `instance_id_base` advances by the batch sizes so every instance keeps its own RNG key."""
        from concurrent.futures import ThreadPoolExecutor

        base = int(kw.pop("instance_id_base", 0))
        n = len(self.builders)
        with ThreadPoolExecutor(max_workers=n) as pool:
            pending = []
            for i, batch in enumerate(batches):
                if len(pending) == n:  # the context about to be reused must have delivered its previous result
                    yield pending.pop(0).result()
                pending.append(pool.submit(self.builders[i % n].build, batch, instance_id_base=base, **kw))
                base += len(batch)
            for f in pending:
                yield f.result()


def get_builder(pred_basename: str, device: int = 0) -> RoadmapBuilder:
    """one cached builder per (weights, device): the reference re-reads the model files on every call
    (learned_sampler.py:53, inside the timed region of scripts/eval.py:57-61); here they stay resident"""
    key = (pred_basename, int(device))
    if key not in _BUILDERS:
        _BUILDERS[key] = RoadmapBuilder(pred_basename, device)
    return _BUILDERS[key]


def get_timed_roadmaps_multiple_paths_with_learned_indicator(
    ins: Instance,
    pred_basename: str,
    prob_uniform_sampling_after_goal: float = 0.9,
    prob_uniform_bias: float = 0.0,
    prob_uniform_gamma: float = 5.0,
    max_T: int = 64,
    N_traj: int = 25,
    max_attempt: int = 3,
    randomize_indicator: bool = False,
    merge_distance_rate: float = 0.25,
    verbose: int = 0,
    seed: Optional[int] = None,
    lazy: bool = True,
):
    """drop-in for the reference function of the same name (learned_sampler.py:19-204); returns one TimedRoadmap per
    agent (a lazy view over the GPU's CSR result; `lazy=False` materialises plain lists eagerly) and updates
    ins.objs.cnt_continuous_collide like the reference's collision calls do (static_objects.py:72-75)."""
    builder = get_builder(pred_basename)
    csr = builder.build([ins], prob_uniform_sampling_after_goal=prob_uniform_sampling_after_goal,
                        prob_uniform_bias=prob_uniform_bias, prob_uniform_gamma=prob_uniform_gamma, max_T=max_T,
                        N_traj=N_traj, max_attempt=max_attempt, randomize_indicator=randomize_indicator,
                        merge_distance_rate=merge_distance_rate, seed=seed).copy()
    if hasattr(ins, "objs") and hasattr(ins.objs, "cnt_continuous_collide"):
        ins.objs.cnt_continuous_collide += int(csr.cnt_collide[0])
        ins.objs.time_continuous_collide += csr.timing.get("total_ms", 0.0) * 1e-3
    if verbose:
        print(f"roadmap_gen: {csr.n_vertices} vertices, {csr.n_edges} edges, {csr.timing}")
    return csr.timed_roadmaps(0, lazy=lazy)
