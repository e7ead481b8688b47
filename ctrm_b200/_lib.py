"""ctypes binding of libctrm_b200.so (include/ctrm_b200.h).

The product path fails loudly when the CUDA library is missing or no GPU is present: there is no CPU
fallback.  `load()` only dlopen()s the library and checks the ABI version (that works on a CPU-only box and
is what the `-m "not gpu"` tests use); every compute call needs a device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libctrm_b200.so")
ABI_VERSION = 1

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)


class ModelDesc(C.Structure):
    _fields_ = [("fov_size", C.c_int32), ("map_size", C.c_int32), ("dim_hidden", C.c_int32),
                ("dim_message", C.c_int32), ("dim_attention", C.c_int32), ("dim_ind", C.c_int32),
                ("dim_latent", C.c_int32), ("num_neighbors", C.c_int32), ("use_k_neighbor", C.c_int32),
                ("include_self_attention", C.c_int32), ("temp", C.c_float)]


class Params(C.Structure):
    _fields_ = [("prob_uniform_sampling_after_goal", C.c_double), ("prob_uniform_bias", C.c_double),
                ("prob_uniform_gamma", C.c_double), ("max_T", C.c_int32), ("N_traj", C.c_int32),
                ("max_attempt", C.c_int32), ("randomize_indicator", C.c_int32), ("merge_distance_rate", C.c_double),
                ("K", C.c_int32), ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("instance_id_base", C.c_uint32)]


class Tables(C.Structure):
    _fields_ = [("h_gate", C.c_void_p), ("h_rw", C.c_void_p), ("h_step", C.c_void_p), ("h_teacher", C.c_void_p),
                ("h_uniforms", C.c_void_p), ("h_p_rw", C.c_void_p), ("h_merge2", C.c_void_p)]


_OFF_NAMES = ["fov_w0", "fov_b0", "fovs_w1", "fovs_b1", "fovs_w2", "fovs_b2", "fovv_w1", "fovv_b1", "fovv_w2", "fovv_b2",
              "ag_w0i", "ag_w0v", "ag_b0", "ag_w1", "ag_b1", "ag_w2", "ag_b2",
              "ind_w0", "ind_b0", "ind_w1", "ind_b1", "ind_w2", "ind_b2",
              "enc_w0", "enc_b0", "enc_w1", "enc_b1", "enc_w2", "enc_b2",
              "dec_w0", "dec_b0", "dec_w1", "dec_b1", "dec_w2", "dec_b2", "total"]


class WeightOffsets(C.Structure):
    _fields_ = [(n, C.c_int32) for n in _OFF_NAMES]


class DeviceView(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "A", "O", "M", "F", "L", "S", "W")] + [
        (n, C.c_void_p) for n in ("d_agent_off", "d_obs_off", "d_agent_inst", "d_starts", "d_goals", "d_speeds",
                                  "d_speeds2", "d_rads", "d_obs_xyr", "d_occ", "d_ctg")] + [("ctg_layout", C.c_int32)]


# every symbol include/ctrm_b200.h declares: name -> (restype, argtypes)
V = C.c_void_p
SIGNATURES = {
    "ctrm_abi_version": (C.c_int, []),
    "ctrm_ctx_create": (C.c_int, [C.c_int, C.POINTER(V)]),
    "ctrm_ctx_destroy": (None, [V]),
    "ctrm_last_error": (C.c_char_p, [V]),
    "ctrm_weight_offsets": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(WeightOffsets)]),
    "ctrm_load_weights": (C.c_int, [V, C.POINTER(ModelDesc), V, C.c_size_t]),
    "ctrm_upload_instances": (C.c_int, [V, C.c_int, V, V, V, V, V, V, V, V, V]),
    "ctrm_collide_segments": (C.c_int, [V, V, V, V, C.c_int64, V, V, V, V, V, V]),
    "ctrm_collide_sphere_pairs": (C.c_int, [V, V, V, V, V, V, C.c_int, C.c_int64, V, V]),
    "ctrm_raster_occupancy": (C.c_int, [V, V, C.c_int, C.c_int, V, V]),
    "ctrm_cost_to_go": (C.c_int, [V, V, V, V, C.c_int, C.c_int, V, V]),
    "ctrm_others_info": (C.c_int, [V] * 5 + [C.c_int, C.c_int, V, V, V]),
    "ctrm_collide_points": (C.c_int, [V, V, V, C.c_int64, V, V, V, V]),
    "ctrm_pair_adjacency_words": (C.c_int64, [V, V, C.c_int]),
    "ctrm_pair_adjacency": (C.c_int, [V, V, V, V, C.c_int, V, V, V, V, V, V, C.c_int, V, V, V]),
    "ctrm_adjacency_csr": (C.c_int, [V, V, V, C.c_int, C.c_int, V, V, V]),
    "ctrm_plan_prioritized": (C.c_int, [C.c_int, C.c_int, C.c_int] + [V] * 9 + [C.c_int] + [V] * 6),
    "ctrm_fov_raster": (C.c_int, [V, V, C.c_int, V, V, C.c_int, C.c_int, C.c_int, V, V]),
    "ctrm_encode_features": (C.c_int, [V, V, V, V, V, V]),
    "ctrm_cvae_sample": (C.c_int, [V, V, V, C.c_int, C.c_uint64, C.c_int, C.c_int, V, V, V]),
    "ctrm_build_roadmaps": (C.c_int, [V, C.POINTER(Params), C.POINTER(Tables), V]),
    "ctrm_result_sizes": (C.c_int, [V, c_i64p, c_i64p, c_i64p]),
    "ctrm_export_csr": (C.c_int, [V, V, V, V, V, V, V, C.c_int, V]),
    "ctrm_export_csr_compact": (C.c_int, [V] * 9 + [C.c_int, V]),
    "ctrm_trace_enable": (C.c_int, [V, C.c_int]),
    "ctrm_trace_read": (C.c_int, [V, V, V, V]),
    "ctrm_last_timing": (C.c_int, [V, V, c_i64p, c_i64p]),
    "ctrm_get_instance_steps": (C.c_int, [V, V]),
    "ctrm_set_step_kernel": (C.c_int, [V, C.c_int]),
    "ctrm_set_stream_priority": (C.c_int, [V, C.c_int]),
    "ctrm_set_profile": (C.c_int, [V, C.c_int]),
    "ctrm_last_kernel_times": (C.c_int, [V, V, V]),
    "ctrm_get_device_view": (C.c_int, [V, C.POINTER(DeviceView)]),
    "ctrm_host_continuous_collide_spheres": (C.c_int, [V, V, C.c_double, V, V, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "ctrm_host_cost_to_go_2d": (C.c_int, [V, V, C.c_int, V, C.c_int, V]),
    "ctrm_host_fov2d": (C.c_int, [V, V, V, C.c_int, C.c_int, V]),
}

_lib = None


class CtrmError(RuntimeError):
    pass


def load():
    """dlopen libctrm_b200.so and bind every symbol of the header.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CtrmError(f"{LIB_PATH} is missing: build it with `python -m ctrm_b200.build` (needs nvcc). "
                        "ctrm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ctrm_abi_version() != ABI_VERSION:
        raise CtrmError("libctrm_b200.so ABI version mismatch: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ctrm_last_error(None)
        raise CtrmError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")


def ptr(arr):
    """host pointer of a C-contiguous numpy array (or None)"""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(C.c_void_p)
