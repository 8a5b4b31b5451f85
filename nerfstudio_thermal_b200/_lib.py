"""ctypes binding of libtn_b200.so (C ABI declared in include/tn_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing, every product
entry point raises `TnLibraryError`.  Build it with `python -c "import __graft_entry__ as g; g.build()"`
(or `nerfstudio_thermal_b200/csrc/build.sh`).
"""
import ctypes
import os
from ctypes import c_double, POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtn_b200.so")


class TnLibraryError(RuntimeError):
    pass


class TnKernelError(RuntimeError):
    pass


_P = c_void_p  # device pointers travel as integers
_FPP = POINTER(c_void_p)  # host array of device pointers

# name -> (argtypes); restype is int unless listed in _RESTYPES
_SIGNATURES = {
    "tn_version": [],
    "tn_last_error_string": [],
    "tn_build_arch": [],
    "tn_hash_encode_fwd": [_P, _P, c_int, POINTER(c_float), c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tn_hash_encode_bwd": [_P, _P, c_int, POINTER(c_float), _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "tn_sample_positions_fwd": [_P, _P, _P, c_int64, c_int, _P, _P, _P],
    "tn_sample_positions_bwd": [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P],
    "tn_contract_points_fwd": [_P, c_int64, _P, _P, _P],
    "tn_contract_points_bwd": [_P, _P, c_int64, _P, _P],
    "tn_mlp_fwd": [_P, c_int64, c_int, c_int, c_int, c_int, _FPP, _FPP, c_int, _P, _P],
    "tn_mlp_bwd": [_P, _P, c_int64, c_int, c_int, c_int, c_int, _FPP, _FPP, c_int, _P, _FPP, _FPP, _P],
    "tn_mlp_tc_fwd": [_P, c_int64, c_int, c_int, c_int, c_int, c_int, _FPP, _FPP, c_int, _P, c_float, _P, _P, _P],
    "tn_mlp_tc_bwd": [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _FPP, _FPP, c_int, _P, c_float, _P, _FPP,
                      _FPP, _P],
    "tn_sh4": [_P, c_int64, _P, _P],
    "tn_ray_features": [_P, _P, _P, c_int64, c_int, _P, _P, _P],
    "tn_piecewise_bins": [_P, _P, _P, _P, c_int, c_int64, c_int, _P, _P, _P],
    "tn_pdf_sample": [_P, _P, _P, _P, _P, _P, c_int, _P, c_int64, c_int, c_int, c_float, c_float, _P, _P, _P],
    "tn_weights_fwd": [_P, _P, c_int64, c_int, _P, _P],
    "tn_weights_bwd": [_P, _P, _P, c_int64, c_int, _P, _P],
    "tn_render_fwd": [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, POINTER(c_float), c_int, _P, _P, _P, _P, _P, _P],
    "tn_render_bwd": [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, POINTER(c_float), _P, _P, _P],
    "tn_field_split_fwd": [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P],
    "tn_field_split_bwd": [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P],
    "tn_density_act_fwd": [_P, c_int, _P, c_int64, c_float, _P, _P],
    "tn_density_act_bwd": [_P, c_int, _P, _P, c_int64, c_float, _P, c_int, _P],
    "tn_prop_density_fwd": [_P, _P, _P, _P, POINTER(c_float), c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P,
                            c_float, _P, _P],
    "tn_prop_density_bwd": [_P, _P, _P, _P, POINTER(c_float), c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P,
                            c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "tn_camera_opt_fwd": [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P],
    "tn_camera_opt_bwd": [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P, _P],
    "tn_pixel_losses": [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P],
    "tn_density_l1": [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, _P, _P, c_int, _P, _P, _P, _P, _P, _P],
    "tn_distortion_loss": [_P, _P, c_int64, c_int, _P, _P, _P],
    "tn_interlevel_loss": [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P],
    "tn_field_head_fwd": [_P, _P, _P, _P, c_int64, c_int, c_int, c_float, _FPP, _FPP, c_int, _P, _P, _P, _P],
    "tn_field_head_bwd": [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _FPP, _FPP, c_int, _P, _P, _FPP,
                          _FPP, _P],
    "tn_camera_reg_fwd": [_P, c_int, c_float, c_float, c_float, _P, _P],
    "tn_camera_reg_bwd": [_P, _P, c_int, c_float, c_float, c_float, c_int, _P, _P],
    "tn_embed_bwd": [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "tn_loss_sum": [_FPP, POINTER(c_float), POINTER(c_int), c_int, c_int, _P, _P],
    "tn_patch_pixel_indices": [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P],
    "tn_gather_pixels": [_P, c_int, c_int64, c_int, c_int, c_int, _P, _P, _P, c_int64, _P, _P, _P],
    "tn_generate_rays": [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P],
    "tn_adam_step": [_P, _P, _P, _P, c_int64, POINTER(c_int64), POINTER(c_int64), POINTER(c_float), c_int, c_double,
                     c_double, _P, c_int, c_int, c_uint, _P, _P, c_int, _P],
    "tn_step_counters_tick": [_P, c_int, c_uint, _P],
    "tn_grad_unscale_check": [_P, c_int64, _P, _P, _P],
    "tn_counter_add": [_P, c_int, _P],
    "tn_level_resample": [_P, _P, _P, _P, _P, _P, _P, c_int, _P, c_int64, c_int, c_int, c_float, c_float, _P, _P, _P, _P,
                          _P],
    "tn_ray_heads_fwd": [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, POINTER(c_float), c_int, c_int, _FPP, _FPP,
                         POINTER(c_int), _FPP, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_float, c_float, _P, _P],
    "tn_ray_heads_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, POINTER(c_float), c_int,
                         _FPP, _FPP, _FPP, POINTER(c_int), _P, _P, _FPP, _P],
    "tn_shard_mean": [_P, _P, c_int, c_int64, c_int64, c_float, c_int, _P],
    "tn_l2_probe": [c_int, _P, c_int, c_int, c_int, _P, _P],
}
_RESTYPES = {"tn_last_error_string": c_char_p, "tn_build_arch": c_char_p}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once).  Raises TnLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TnLibraryError(
            f"{LIB_PATH} not found: the sm_100a CUDA library is required (there is no CPU fallback). "
            "Build it with nerfstudio_thermal_b200/csrc/build.sh or __graft_entry__.build().")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


class LaunchStats:
    """Launch counter and optional per-launch CUDA-event timing of the C-ABI calls (bench.py / profiling).

    `count` always counts launches.  With `timing=True` every call is bracketed by two events on the current
    stream; `summary()` (after a synchronize) returns {tag: (launches, total_ms, total_units)} where tag = entry
    point name plus the caller-supplied shape tag and units = what the caller says a launch processed (points,
    samples; 0 when not given) -- the multiplier of the per-unit algorithmic bytes in the rooflines."""

    def __init__(self):
        self.count = 0
        self.timing = False
        self.events = []
        self.tag = ""

    def reset(self, timing=False):
        self.count, self.timing, self.events = 0, timing, []

    def summary(self):
        out = {}
        for tag, e0, e1, units in self.events:
            n, ms, u = out.get(tag, (0, 0.0, 0))
            out[tag] = (n + 1, ms + e0.elapsed_time(e1), u + units)
        return out


STATS = LaunchStats()


def call(name, *args, tag="", units=0):
    """Invoke an int-returning entry point; map non-zero codes to TnKernelError."""
    lib = load()
    STATS.count += 1
    if STATS.timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        STATS.events.append((f"{name}{tag}", e0, e1, int(units)))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise TnKernelError(f"{name} failed ({rc}): {lib.tn_last_error_string().decode()}")


def ptr(t):
    """Device pointer of a CUDA float32/int32 contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TnKernelError("libtn_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise TnKernelError("libtn_b200 kernels take contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def float_array(values):
    arr = (c_float * len(values))(*[float(v) for v in values])
    return arr


def ptr_array(tensors):
    return (c_void_p * len(tensors))(*[ptr(t) for t in tensors])
