"""ctypes binding of ``libbgflow_b200.so`` (the C ABI declared in ``include/bgflow_b200.h``).

The shared library is built in-tree by ``bgflow_b200._build.build()`` (nvcc, sm_100a).  There
is no fallback: if it cannot be loaded every compute entry point of the package raises.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbgflow_b200.so")

BGX_MAX_LAYERS = 8
BGX_MAX_SEGS = 8

BGX_OK = 0
ERRORS = {-1: "BGX_ERR_INVALID (bad argument / inconsistent shapes)",
          -2: "BGX_ERR_UNSUPPORTED (shape outside what the kernels support)",
          -3: "BGX_ERR_WORKSPACE (workspace too small)",
          -4: "BGX_ERR_CUDA"}

ACT_NONE, ACT_RELU, ACT_SILU, ACT_TANH = 0, 1, 2, 3
FLAG_INVERSE, FLAG_PRESERVE_VOLUME, FLAG_CIRCULAR, FLAG_BF16X6, FLAG_FORCE_SIMT = 1, 2, 4, 8, 16
FLAG_NO_PAIR, FLAG_FORCE_WIDE, FLAG_PREFER_PAIR = 32, 64, 128
KERNEL_IDS = {"spline_pair": 0, "spline_pair_wide": 1, "spline_tc2": 2, "spline_tc": 3, "spline_simt": 4,
              "affine_tc2": 5, "affine_tc": 6, "affine_simt": 7, "affine_pair": 8, "affine_pair_wide": 9}


class bgx_mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("act", C.c_int32),
                ("dims", C.c_int32 * (BGX_MAX_LAYERS + 1)),
                ("W", C.c_void_p * BGX_MAX_LAYERS), ("b", C.c_void_p * BGX_MAX_LAYERS),
                ("raw_width", C.c_int32), ("n_periodic", C.c_int32),
                ("periodic_idx", C.POINTER(C.c_int32)),
                ("periodic_left", C.c_float), ("periodic_right", C.c_float)]


class bgx_packed_mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("act", C.c_int32),
                ("K", C.c_int32 * BGX_MAX_LAYERS), ("N", C.c_int32 * BGX_MAX_LAYERS),
                ("Kp", C.c_int32 * BGX_MAX_LAYERS), ("Np", C.c_int32 * BGX_MAX_LAYERS),
                ("Wt", C.c_void_p * BGX_MAX_LAYERS), ("bias", C.c_void_p * BGX_MAX_LAYERS),
                ("in_map", C.c_void_p),
                ("periodic_scale", C.c_float), ("periodic_left", C.c_float),
                ("raw_width", C.c_int32),
                ("spline_dims_per_pass", C.c_int32), ("spline_stride", C.c_int32),
                ("total_floats", C.c_int64),
                ("Wb", (C.c_void_p * BGX_MAX_LAYERS) * 3),
                ("spline_bias", C.c_void_p), ("spline_bias_pad", C.c_int32), ("reserved_", C.c_int32)]


class bgx_train_mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (BGX_MAX_LAYERS + 1)), ("act", C.c_int32 * BGX_MAX_LAYERS),
                ("fwd", bgx_packed_mlp * BGX_MAX_LAYERS), ("bwd", bgx_packed_mlp * BGX_MAX_LAYERS),
                ("total_floats", C.c_int64)]


class bgx_train_buffers(C.Structure):
    _fields_ = [("z", C.c_void_p * BGX_MAX_LAYERS), ("h", C.c_void_p * BGX_MAX_LAYERS), ("g", C.c_void_p * BGX_MAX_LAYERS),
                ("part", C.c_void_p)]


class bgx_spline_layout(C.Structure):
    _fields_ = [("d_t", C.c_int32), ("n_bins", C.c_int32), ("is_circular", C.POINTER(C.c_uint8))]


class bgx_seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("width", C.c_int32), ("stride", C.c_int32)]


class bgx_coupling_io(C.Structure):
    _fields_ = [("batch", C.c_int64), ("n_cond", C.c_int32), ("cond", bgx_seg * BGX_MAX_SEGS),
                ("n_tr", C.c_int32), ("tr_in", bgx_seg * BGX_MAX_SEGS),
                ("tr_out", bgx_seg * BGX_MAX_SEGS),
                ("dlogp_in", C.c_void_p), ("dlogp_out", C.c_void_p)]


class bgx_spline_cfg(C.Structure):
    _fields_ = [("n_bins", C.c_int32), ("left", C.c_float), ("right", C.c_float),
                ("bottom", C.c_float), ("top", C.c_float), ("min_bin_width", C.c_float),
                ("min_bin_height", C.c_float), ("min_derivative", C.c_float),
                ("identity_init", C.c_int32), ("oob_counter", C.c_void_p), ("status", C.c_void_p)]


class bgx_zplan(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("seeds", C.c_int32 * 3), ("n_rel", C.c_int32),
                ("rel", C.c_void_p), ("order", C.c_void_p), ("slot_of_col", C.c_void_p),
                ("normalize_angles", C.c_int32),
                ("eps", C.c_float)]


class bgx_cdf_col(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p", C.c_float * 7)]


class bgx_relplan(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("n_fixed", C.c_int32), ("n_rel", C.c_int32),
                ("fixed", C.c_void_p), ("rel", C.c_void_p), ("order", C.c_void_p),
                ("normalize_angles", C.c_int32), ("eps", C.c_float), ("keepdims", C.c_int32),
                ("mean", C.c_void_p), ("blacken", C.c_void_p), ("whiten", C.c_void_p),
                ("log_det_whiten", C.c_float)]


DIST_NONE, DIST_NORMAL, DIST_TRUNCNORMAL, DIST_UNIFORM = 0, 1, 2, 3

# every symbol include/bgflow_b200.h declares: (name, restype, argtypes)
P = C.POINTER
SYMBOLS = {
    "bgx_pack_mlp": (C.c_int, [P(bgx_mlp), P(bgx_spline_layout), C.c_void_p, C.c_int64,
                               P(bgx_packed_mlp), C.c_void_p]),
    "bgx_affine_coupling": (C.c_int, [P(bgx_coupling_io), P(bgx_packed_mlp), P(bgx_packed_mlp),
                                      C.c_float, C.c_int, C.c_void_p]),
    "bgx_spline_coupling": (C.c_int, [P(bgx_coupling_io), P(bgx_packed_mlp), P(bgx_spline_cfg),
                                      C.c_int, C.c_void_p]),
    "bgx_spline_backward": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, P(bgx_spline_cfg), C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "bgx_ic_to_xyz": (C.c_int, [P(bgx_zplan), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_ic_from_xyz": (C.c_int, [P(bgx_zplan), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "bgx_cdf_col_init": (C.c_int, [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, P(bgx_cdf_col)]),
    "bgx_cdf_col_set_truncation": (C.c_int, [P(bgx_cdf_col), C.c_double, C.c_double]),
    "bgx_cdf_map": (C.c_int, [C.c_int64, C.c_int32, P(bgx_seg), P(bgx_seg), C.c_void_p, C.c_float, C.c_float,
                              C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_ic_to_xyz_mapped": (C.c_int, [P(bgx_zplan), C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_ic_from_xyz_mapped": (C.c_int, [P(bgx_zplan), C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_relic_to_xyz": (C.c_int, [P(bgx_relplan), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_relic_from_xyz": (C.c_int, [P(bgx_relplan), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_split_merge": (C.c_int, [C.c_int64, P(bgx_seg), C.c_int32, P(bgx_seg), C.c_int, C.c_void_p]),
    "bgx_linear": (C.c_int, [C.c_int64, C.c_void_p, P(bgx_packed_mlp), C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_train_pack": (C.c_int, [P(bgx_mlp), P(C.c_int32), C.c_void_p, C.c_int64, P(bgx_train_mlp), C.c_void_p]),
    "bgx_mlp_train_part_floats": (C.c_int64, [C.c_int64, P(bgx_train_mlp)]),
    "bgx_mlp_forward_train": (C.c_int, [C.c_int64, P(bgx_train_mlp), C.c_void_p, P(bgx_train_buffers), C.c_void_p, C.c_void_p]),
    "bgx_mlp_backward": (C.c_int, [C.c_int64, P(bgx_train_mlp), C.c_void_p, P(bgx_train_buffers), C.c_void_p, C.c_void_p,
                                   P(C.c_void_p), P(C.c_void_p), C.c_void_p, C.c_void_p]),
    "bgx_spline_coupling_backward": (C.c_int, [C.c_int64, P(bgx_train_mlp), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, P(bgx_spline_cfg), C.c_int, P(bgx_train_buffers),
                                               C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_void_p), P(C.c_void_p),
                                               C.c_void_p, C.c_void_p]),
    "bgx_affine_backward_scratch_floats": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "bgx_affine_coupling_backward": (C.c_int, [C.c_int64, P(bgx_train_mlp), P(bgx_train_mlp), C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, P(bgx_train_buffers),
                                               P(bgx_train_buffers), C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_void_p),
                                               P(C.c_void_p), P(C.c_void_p), P(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_gemm_tn_slices": (C.c_int, [C.c_int64, C.c_int]),
    "bgx_gemm_tn": (C.c_int, [C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_split_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgx_tc_selftest": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "bgx_set_status_buffer": (C.c_int, [C.c_void_p]),
    "bgx_debug_set_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "bgx_version": (C.c_char_p, []),
    "bgx_last_cuda_error": (C.c_char_p, []),
    "bgx_launch_count": (C.c_int64, []),
    "bgx_kernel_count": (C.c_int64, [C.c_int]),
}

_lib = None


class BgxError(RuntimeError):
    pass


def load():
    """Load the shared library (building it first if the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError here == the ABI is incomplete: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != BGX_OK:
        msg = ERRORS.get(rc, str(rc))
        if rc == -4:
            msg += ": " + load().bgx_last_cuda_error().decode()
        raise BgxError(f"{what} failed: {msg}")


def launch_count():
    return int(load().bgx_launch_count())


def kernel_counts():
    """Coupling calls served so far per kernel family (``KERNEL_IDS``)."""
    lib = load()
    return {name: int(lib.bgx_kernel_count(i)) for name, i in KERNEL_IDS.items()}
