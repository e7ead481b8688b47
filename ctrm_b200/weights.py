"""Load the trained CTRM networks and pack them into the fp32 blob libctrm_b200 consumes.

Mirrors the load half of the reference's `reconstruct` (learning/utils.py:32-53): reconstruct_net
(learning/model/__init__.py:19-59: pickled hyper-parameters + state_dict, `.eval()`), reconstruct_format_input
(learning/formats/format_input.py:70-82, :190-198) and reconstruct_format_output (format_output.py:40-42).
File suffixes are the reference's (model/model.py:12-13, agent_encoder.py:28-29, format_input.py:144-147,
format_output.py:20).  A `.npz` produced by oracle/gen_golden.py (same state_dict keys) is accepted too.

BatchNorm1d runs in eval mode on this path (model/__init__.py:53), so it is folded into the preceding Linear here,
in float64, and rounded once to float32.
"""
from __future__ import annotations

import ctypes as C
import io
import os
import pickle
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class ModelBundle:
    arrays: dict          # state_dict entries, keys '<net>.<key>' with net in model/agent/fov_self/fov_vain
    fov_size: int = 19
    map_size: int = 160
    num_neighbors: int = 15
    use_k_neighbor: bool = True
    include_self_attention: bool = False
    dim_ind: int = 3
    dim_latent: int = 64
    dim_message: int = 32
    dim_attention: int = 10
    dim_hidden: int = 32
    temp: float = 2.0

    def desc(self) -> _lib.ModelDesc:
        return _lib.ModelDesc(self.fov_size, self.map_size, self.dim_hidden, self.dim_message, self.dim_attention,
                              self.dim_ind, self.dim_latent, self.num_neighbors, int(self.use_k_neighbor),
                              int(self.include_self_attention), float(self.temp))


class _Stub:
    """placeholder for reference classes inside the pickles (Format2D_CTRM_Input / _Output instances)"""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _RefUnpickler(pickle.Unpickler):
    """the format pickles reference classes of the `ctrm` package; only their attribute dicts are needed"""

    def find_class(self, module, name):
        if module.split(".")[0] == "ctrm":
            return type(name, (_Stub,), {"__module__": module})
        return super().find_class(module, name)


def _unpickle(path):
    with open(path, "rb") as f:
        return _RefUnpickler(io.BytesIO(f.read())).load()


def resolve_basename(basename: str) -> str:
    """learning/utils.py:44-45: a 19-character basename is a date label under /workspace/log"""
    if len(basename) == 19:
        return f"/workspace/log/{basename}/best"
    return basename


def load_bundle(basename_or_npz: str) -> ModelBundle:
    if basename_or_npz.endswith(".npz"):
        z = np.load(basename_or_npz, allow_pickle=False)
        arrays = {k: np.asarray(z[k]) for k in z.files if not k.startswith("hp.")}
        hp = {k[3:]: z[k].item() for k in z.files if k.startswith("hp.")}
        return ModelBundle(arrays=arrays, fov_size=int(hp["fov_size"]), map_size=int(hp["map_size"]),
                           num_neighbors=int(hp["num_neighbors"]), use_k_neighbor=bool(hp["use_k_neighbor"]),
                           include_self_attention=bool(hp["include_self_attention"]), dim_ind=int(hp["dim_ind"]),
                           dim_latent=int(hp["dim_latent"]), dim_message=int(hp["dim_message"]),
                           dim_attention=int(hp["dim_attention"]), temp=float(hp["temp"]))
    import torch

    base = resolve_basename(basename_or_npz)
    arrays = {}
    for prefix, suffix in (("model", "_model.pt"), ("agent", "_agent_encoder.pt"),
                           ("fov_self", "_fov_encoder_self.pt"), ("fov_vain", "_fov_encoder_vain.pt")):
        sd = torch.load(base + suffix, map_location="cpu")
        for k, v in sd.items():
            if k.startswith("encoder_output") or k.endswith("num_batches_tracked"):
                continue  # train-only branch (cvae.py:66-68, :140-172)
            arrays[f"{prefix}.{k}"] = v.detach().cpu().numpy()
    hyp = _unpickle(base + "_model_hypra.pkl")
    fov = _unpickle(base + "_fov_encoder_self.pkl")
    age = _unpickle(base + "_agent_encoder.pkl")
    fi = _unpickle(base + "_format_input.pkl")
    fo = _unpickle(base + "_format_output.Pl")
    fid, fod = fi.__dict__, fo.__dict__
    # the kernels implement ONE architecture (the reference's defaults, trained_models/aamas22-main): anything else would
    # load without error and sample wrongly (a sigmoid or a dropped comm vector silently ignored), so refuse it here
    fov_v = _unpickle(base + "_fov_encoder_vain.pkl")
    for name, hp_ in (("fov_encoder_self", fov), ("fov_encoder_vain", fov_v), ("agent_encoder", age)):
        for key, want in (("use_sigmoid", False), ("use_batch_norm", False), ("num_mid_layers", 1), ("dim_hidden", 32)):
            if key in hp_ and hp_[key] != want:
                raise _lib.CtrmError(f"{name}: {key}={hp_[key]!r} is not supported (need {want!r})")
    for key, want in (("num_mid_layers_encoder", 1), ("num_mid_layers_decoder", 1), ("dim_hidden", 32), ("dim_input", 72)):
        if key in hyp and hyp[key] != want:
            raise _lib.CtrmError(f"model: {key}={hyp[key]!r} is not supported (need {want!r})")
    if fid.get("without_comm", 0):
        raise _lib.CtrmError("format_input: without_comm=True is not supported")
    if (int(fov["fov_size"]), int(fov["map_size"])) != (int(fov_v["fov_size"]), int(fov_v["map_size"])):
        raise _lib.CtrmError("the two FOV encoders disagree on fov_size / map_size")
    return ModelBundle(arrays=arrays, fov_size=int(fov["fov_size"]), map_size=int(fov["map_size"]),
                       num_neighbors=int(fid.get("num_neighbors", 15)), use_k_neighbor=bool(fid.get("use_k_neighbor", True)),
                       include_self_attention=bool(fid.get("include_self_attention", False)),
                       dim_ind=int(fod.get("num_divide", fod.get("dim_indicators", 3))),
                       dim_latent=int(hyp["dim_latent"]), dim_message=int(age["dim_message"]),
                       dim_attention=int(age["dim_attention"]), temp=float(hyp.get("temp", 2.0)))


def _lin_T(w, ld=None):
    """torch Linear weight [out][in] -> [in][ld] row-major"""
    wt = np.asarray(w, dtype=np.float64).T
    if ld is not None and ld != wt.shape[1]:
        pad = np.zeros((wt.shape[0], ld), dtype=np.float64)
        pad[:, :wt.shape[1]] = wt
        wt = pad
    return wt


def _pad(b, ld):
    out = np.zeros(ld, dtype=np.float64)
    out[:len(b)] = np.asarray(b, dtype=np.float64)
    return out


def _fold_bn(arr, prefix, lin, bn, eps=1e-5):
    """Linear followed by eval-mode BatchNorm1d -> one Linear (float64)"""
    W = np.asarray(arr[f"{prefix}.{lin}.weight"], dtype=np.float64)
    b = np.asarray(arr[f"{prefix}.{lin}.bias"], dtype=np.float64)
    g = np.asarray(arr[f"{prefix}.{bn}.weight"], dtype=np.float64)
    be = np.asarray(arr[f"{prefix}.{bn}.bias"], dtype=np.float64)
    mu = np.asarray(arr[f"{prefix}.{bn}.running_mean"], dtype=np.float64)
    var = np.asarray(arr[f"{prefix}.{bn}.running_var"], dtype=np.float64)
    s = g / np.sqrt(var + eps)
    return W * s[:, None], (b - mu) * s + be


def pack_blob(bundle: ModelBundle):
    """-> (ModelDesc, WeightOffsets, float32 blob) in the layout of ctrm_weight_offsets()"""
    lib = _lib.load()
    desc = bundle.desc()
    off = _lib.WeightOffsets()
    _lib.check(lib.ctrm_weight_offsets(C.byref(desc), C.byref(off)), "ctrm_weight_offsets")
    a = bundle.arrays
    blob = np.zeros(off.total, dtype=np.float64)

    def put(name, x):
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        o = getattr(off, name)
        nxt = min([v for f, _ in off._fields_ if (v := getattr(off, f)) > o] or [off.total])
        if o + x.size > nxt:  # a checkpoint of another width must not spill into the neighbouring slot
            raise _lib.CtrmError(f"weight block {name}: {x.size} values do not fit its slot of {nxt - o}")
        blob[o:o + x.size] = x

    put("fov_w0", np.concatenate([_lin_T(a["fov_self.mlp.0.weight"]), _lin_T(a["fov_vain.mlp.0.weight"])], axis=1))
    put("fov_b0", np.concatenate([a["fov_self.mlp.0.bias"], a["fov_vain.mlp.0.bias"]]))
    for tag, net in (("fovs", "fov_self"), ("fovv", "fov_vain")):
        put(f"{tag}_w1", _lin_T(a[f"{net}.mlp.2.weight"]))
        put(f"{tag}_b1", a[f"{net}.mlp.2.bias"])
        put(f"{tag}_w2", _lin_T(a[f"{net}.mlp.4.weight"]))
        put(f"{tag}_b2", a[f"{net}.mlp.4.bias"])
    w0 = _lin_T(a["agent.mlp.0.weight"])  # [43][32]: rows 0..10 others_info, 11..42 fov_vain (format_input.py:328-331)
    put("ag_w0i", w0[:11])
    put("ag_w0v", w0[11:])
    put("ag_b0", a["agent.mlp.0.bias"])
    put("ag_w1", _lin_T(a["agent.mlp.2.weight"]))
    put("ag_b1", a["agent.mlp.2.bias"])
    put("ag_w2", _lin_T(a["agent.mlp.4.weight"], 44))
    put("ag_b2", _pad(a["agent.mlp.4.bias"], 44))
    for tag, net, ld_out in (("ind", "model.indicator", 4), ("enc", "model.encoder_input", 64), ("dec", "model.decoder", 4)):
        W0, b0 = _fold_bn(a, net, 0, 1)
        W1, b1 = _fold_bn(a, net, 3, 4)
        put(f"{tag}_w0", _lin_T(W0))
        put(f"{tag}_b0", b0)
        put(f"{tag}_w1", _lin_T(W1))
        put(f"{tag}_b1", b1)
        put(f"{tag}_w2", _lin_T(a[f"{net}.6.weight"], ld_out))
        put(f"{tag}_b2", _pad(a[f"{net}.6.bias"], ld_out))
    return desc, off, np.ascontiguousarray(blob.astype(np.float32))
