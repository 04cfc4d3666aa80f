"""ctypes binding of liblgm_b200.so (include/lgm_b200.h).  There is NO fallback: if the CUDA library is missing or
a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblgm_b200.so")
GRAD_ROW = 12


class RenderParams(ctypes.Structure):
    _fields_ = [
        ("n_scenes", ctypes.c_int32), ("n_gaussians", ctypes.c_int32), ("n_views", ctypes.c_int32),
        ("image_height", ctypes.c_int32), ("image_width", ctypes.c_int32), ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float), ("scale_modifier", ctypes.c_float),
    ]


class LgmError(RuntimeError):
    pass


_lib = None
ABI_VERSION = 4  # include/lgm_b200.h LGM_ABI_VERSION
_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
_pp = ctypes.POINTER(RenderParams)

_SIGNATURES = {
    "lgm_abi_version": (ctypes.c_int, []),
    "lgm_last_error_string": (ctypes.c_char_p, []),
    "lgm_tiles_per_view": (ctypes.c_int, [_i32, _i32]),
    "lgm_num_block_sums": (_i64, [_i32, _i32]),
    "lgm_bin_workspace_bytes": (ctypes.c_int, [_pp, _i64, _i64, ctypes.POINTER(_sz)]),
    "lgm_forward_geom": (ctypes.c_int, [_vp, _pp] + [_vp] * 12),
    "lgm_forward_geom_cov3d": (ctypes.c_int, [_vp, _pp] + [_vp] * 13),
    "lgm_forward_geom_rows": (ctypes.c_int, [_vp, _pp] + [_vp] * 14),
    "lgm_backward_geom_cov3d": (ctypes.c_int, [_vp, _pp] + [_vp] * 8 + [_i32, _vp, _vp]),
    "lgm_set_tuning": (ctypes.c_int, [ctypes.c_char_p, _i32]),
    "lgm_count_workspace_bytes": (ctypes.c_int, [_pp, ctypes.POINTER(_sz)]),
    "lgm_forward_count": (ctypes.c_int, [_vp, _pp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lgm_direct_bin_tile_cap": (ctypes.c_int, []),
    "lgm_forward_bin": (ctypes.c_int, [_vp, _pp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _i32]),
    "lgm_last_bin_coarse": (ctypes.c_int, []),
    "lgm_forward_composite": (ctypes.c_int, [_vp, _pp] + [_vp] * 8 + [_i32] + [_vp] * 4),
    "lgm_forward_bin_render": (ctypes.c_int, [_vp, _pp] + [_vp] * 7 + [_i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _i32] + [_vp] * 4),
    "lgm_backward": (ctypes.c_int, [_vp, _pp] + [_vp] * 19 + [_i32]),
    "lgm_backward_composite": (ctypes.c_int, [_vp, _pp] + [_vp] * 14),
    "lgm_backward_geom": (ctypes.c_int, [_vp, _pp] + [_vp] * 8 + [_i32]),
    "lgm_screen_gradients": (ctypes.c_int, [_vp, _pp, _vp, _vp, _vp]),
    "lgm_last_bin_mode": (ctypes.c_int, []),
    "lgm_mark_visible": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "lgm_activate_forward": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _i32, _vp]),
    "lgm_activate_backward": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp]),
    "lgm_mse_loss_grad": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, ctypes.c_float, _vp, _vp, _vp, _i64, ctypes.c_float, _vp, _vp]),
    "lgm_resize_bilinear_forward": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, ctypes.c_float, ctypes.c_float]),
    "lgm_resize_bilinear_backward": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, ctypes.c_float]),
    "lgm_mse_loss_grad_u8": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, ctypes.c_float, _vp, _vp, _vp, _i64, ctypes.c_float, _vp, _vp]),
    "lgm_sh_forward": (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lgm_sh_backward": (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lgm_sort_input_is_tmp": (ctypes.c_int, [_i32]),
    "lgm_sort_workspace_bytes": (ctypes.c_int, [_i64, _i32, ctypes.POINTER(_sz)]),
    "lgm_sort_pairs": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load the CUDA library (once).  Raises LgmError when it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LgmError(
                f"{LIB_PATH} not found: build it with `python -m lgm_b200.build` (nvcc, sm_100a). "
                "lgm_b200 has no CPU or PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.lgm_abi_version() != ABI_VERSION:  # a stale build: the argument lists below would not match
            raise LgmError(f"{LIB_PATH} has ABI version {l.lgm_abi_version()}, this package needs {ABI_VERSION}: "
                           "rebuild with `python -m lgm_b200.build --force`")
        _lib = l
    return _lib


# Tuning / test hooks of the library (lgm_set_tuning) and the environment variables that drive them from Python.  The
# C library itself never reads the environment; apply_env_tuning() is called by ops at the head of every render.
_TUNING_ENV = {"fwd_batch": "LGM_FWD_BATCH", "bwd_batch": "LGM_BWD_BATCH", "patch_lanes": "LGM_PATCH_LANES",
               "sort_variant": "LGM_SORT_VARIANT", "enum_global": "LGM_ENUM_GLOBAL", "coarse_ratio": "LGM_COARSE_RATIO", "c2_occ": "LGM_C2_OCC", "sort_bulk": "LGM_SORT_BULK", "sparse_lanes": "LGM_SPARSE_LANES", "fine_tile_major": "LGM_FINE_TILE_MAJOR"}
_tuning_applied = {}
BIN_MODE_IDS = {"auto": 0, "onesweep": 1, "hybrid": 2, "direct": 3}


def set_tuning(name, value):
    check(lib().lgm_set_tuning(name.encode(), int(value)), f"lgm_set_tuning({name})")


def apply_env_tuning():
    """Forward LGM_* tuning variables to the library when they changed; returns the binning mode id (LGM_BIN_MODE)."""
    env = os.environ
    for name, var in _TUNING_ENV.items():
        v = env.get(var)
        if _tuning_applied.get(name) != v:
            try:
                set_tuning(name, -1 if v is None else int(v))
            except ValueError:
                set_tuning(name, -1)
            _tuning_applied[name] = v
    return BIN_MODE_IDS.get(env.get("LGM_BIN_MODE", "auto"), 0)


def check(rc, what):
    if rc != 0:
        msg = lib().lgm_last_error_string().decode("utf-8", "replace")
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise LgmError(f"{what} failed ({kind} {rc}): {msg}")


def ptr(t):
    """Device (or host) pointer of a tensor as c_void_p; None -> NULL."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def make_params(n_scenes, n_gaussians, n_views, H, W, tanfovx, tanfovy, scale_modifier):
    p = RenderParams()
    p.n_scenes, p.n_gaussians, p.n_views = int(n_scenes), int(n_gaussians), int(n_views)
    p.image_height, p.image_width = int(H), int(W)
    p.tanfovx, p.tanfovy, p.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    return p
