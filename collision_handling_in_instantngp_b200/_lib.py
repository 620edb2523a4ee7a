"""ctypes binding of libgngf_sm100.so (the C ABI declared in include/gngf.h).

There is no fallback: if the library is missing or a call fails, this module raises.  The shared object is
built in-tree by ``__graft_entry__.build()`` / ``make -C collision_handling_in_instantngp_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

MAX_LEVELS = 32
MAX_FEATURES = 8
MAX_TOPK = 128

MIX_SOFTMAX, MIX_WEIGHTED_AVG, MIX_RAW = 1, 0, 2
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU, ACT_SIGMOID = 0, 1, 2, 3

LIB_NAME = "libgngf_sm100.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


class GngfError(RuntimeError):
    pass


class Lattice(Structure):
    """gngf_lattice (include/gngf.h)."""
    _fields_ = [
        ("num_levels", c_int32),
        ("n", c_int32 * MAX_LEVELS),
        ("ox", c_int32), ("oy", c_int32), ("wx", c_int32), ("wy", c_int32),
        ("lox", c_int32 * MAX_LEVELS),
        ("loy", c_int32 * MAX_LEVELS),
        ("lwx", c_int32 * MAX_LEVELS),
        ("lwy", c_int32 * MAX_LEVELS),
        ("loff", c_int64 * (MAX_LEVELS + 1)),
    ]

    @property
    def num_nodes(self) -> int:          # U
        return self.wx * self.wy

    @property
    def num_level_nodes(self) -> int:    # S
        return self.loff[self.num_levels]


class Tables(Structure):
    """gngf_tables (include/gngf.h)."""
    _fields_ = [("ptr", c_void_p * MAX_LEVELS)]


class AdamTensor(Structure):
    """gngf_adam_tensor (include/gngf.h)."""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("step", c_void_p), ("n", c_int64),
                ("lr", c_float), ("weight_decay", c_float)]


ADAM_MAX_TENSORS = 64

_P = c_void_p  # device pointers travel as integers (tensor.data_ptr())

# name -> (restype, argtypes); every symbol declared in include/gngf.h appears here
SIGNATURES = {
    "gngf_strerror": (c_char_p, [c_int]),
    "gngf_abi_version": (c_int, []),
    "gngf_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "gngf_launch_count": (c_int64, []),
    "gngf_corners_fwd": (c_int, [_P, c_int64, Lattice, _P, _P, _P]),
    "gngf_fast_hash_fwd": (c_int, [_P, c_int64, Lattice, c_int64, _P, _P]),
    "gngf_hpd_first_layer_fwd": (c_int, [Lattice, _P, _P, c_int32, c_int32, _P, _P]),
    "gngf_linear_fwd": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "gngf_linear_bwd": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "gngf_sigmoid_bwd": (c_int, [_P, _P, c_int64, _P, _P]),
    "gngf_hpd_first_layer_bwd": (c_int, [Lattice, _P, c_int32, _P, _P, _P]),
    "gngf_hpd_first_layer_fwd_nodes": (c_int, [Lattice, _P, c_int64, _P, _P, c_int32, c_int32, _P, _P]),
    "gngf_hpd_first_layer_bwd_nodes": (c_int, [Lattice, _P, c_int64, _P, c_int32, _P, _P, _P]),
    "gngf_active_nodes_bitmap_words": (c_int64, [c_int64]),
    "gngf_active_nodes_chunks": (c_int64, [c_int64]),
    "gngf_lattice_mark_nodes": (c_int, [_P, c_int64, Lattice, _P, _P]),
    "gngf_compact_nodes": (c_int, [_P, c_int64, _P, _P, c_int64, _P, _P]),
    "gngf_scatter_node_rows": (c_int, [_P, c_int64, _P, c_int64, _P, _P]),
    "gngf_tc_gemm_set_formats": (c_int, [c_int32, c_int32]),
    "gngf_bitmap_or": (c_int, [_P, c_int32, c_int64, _P, _P]),
    "gngf_gather_node_adjoints": (c_int, [Lattice, _P, c_int64, c_int32, _P, _P, _P, _P, _P]),
    "gngf_split_bf16x3": (c_int, [_P, c_int64, _P, _P]),
    "gngf_split_f16x2": (c_int, [_P, c_int64, _P, _P, _P]),
    "gngf_split_bf16x3_t": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P]),
    "gngf_tc_gemm_bf16x3": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "gngf_hpd_stream_workspace_floats": (c_int64, [c_int64, c_int64, c_int32]),
    "gngf_hpd_stream_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P]),
    "gngf_hpd_stream_refined_workspace_floats": (c_int64, [c_int64, c_int64]),
    "gngf_hpd_stream_fwd_refined": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int32, _P, _P, _P, _P,
                                            _P, _P]),
    "gngf_hpd_stream_bwd_workspace_floats": (c_int64, [c_int64, c_int32]),
    "gngf_hpd_stream_bwd_stats": (c_int, [c_void_p, c_int32]),
    "gngf_hpd_stream_bwd": (c_int, [Lattice, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int32, _P, _P, _P, _P,
                                    _P, _P, _P, c_int32, _P, _P, _P, _P, _P]),
    "gngf_hpd_stream_bwd_nodes": (c_int, [Lattice, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int32, _P,
                                          _P, _P, _P, _P, _P, _P, c_int32, _P, _P, _P, _P, _P]),
    "gngf_hpd_small_supported": (c_int, [c_int32, POINTER(c_int32), c_int32]),
    "gngf_hpd_small_fwd": (c_int, [Lattice, c_int32, POINTER(c_int32), POINTER(c_void_p), POINTER(c_void_p),
                                   POINTER(c_void_p), c_int32, _P, _P, _P, _P]),
    "gngf_hpd_small_bwd": (c_int, [Lattice, c_int32, POINTER(c_int32), POINTER(c_void_p), POINTER(c_void_p),
                                   POINTER(c_void_p), POINTER(c_void_p), _P, c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gngf_hpd_small_fwd_enc": (c_int, [Lattice, c_int32, POINTER(c_int32), POINTER(c_void_p), POINTER(c_void_p),
                                       POINTER(c_void_p), c_int32, _P, _P, _P, Tables, c_int32, c_int32, _P, _P]),
    "gngf_hpd_small_bwd_enc": (c_int, [Lattice, c_int32, POINTER(c_int32), POINTER(c_void_p), POINTER(c_void_p),
                                       POINTER(c_void_p), POINTER(c_void_p), _P, c_int32, _P, _P, _P, _P, _P, _P, _P,
                                       Tables, Tables, c_int32, c_int32, _P, _P]),
    "gngf_mlp3_supported": (c_int, [c_int32, c_int32, c_int32, c_int32]),
    "gngf_mlp3_fwd": (c_int, [_P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gngf_mlp3_bwd_workspace_floats": (c_int64, [c_int32, c_int32]),
    "gngf_mlp3_bwd": (c_int, [_P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P,
                              _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gngf_mlp3_tc_supported": (c_int, [c_int32, c_int32, c_int32, c_int32]),
    "gngf_mlp3_tc_fwd": (c_int, [_P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gngf_mlp3_tc_bwd": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P,
                                 _P, _P, _P, _P, _P, _P, _P, _P]),
    "gngf_peer_allreduce": (c_int, [_P, _P, c_int32, c_int32, _P, _P, c_int64, c_int64, c_int32, c_float, _P, _P]),
    "gngf_peer_allreduce_set_timeout_ms": (c_int, [c_int64]),
    "gngf_counts_per_level": (c_int, [_P, c_int64, Lattice, _P, c_int64, c_int64, _P, _P, _P, _P]),
    "gngf_adam_step": (c_int, [POINTER(AdamTensor), c_int32, c_float, c_float, c_float, _P, _P]),
    "gngf_loss_fwd_bwd": (c_int, [_P, _P, c_int64, _P, c_int32, c_int64, c_float, c_float, c_float, c_float, c_float,
                                  _P, _P, _P, _P, _P]),
    "gngf_loss_parts": (c_int, [_P, _P, c_int64, _P, c_int32, c_int64, c_float, c_float, c_float, c_float, c_float,
                                _P, _P, _P, _P, c_int32, _P]),
    "gngf_count_distinct_workspace_words": (c_int64, [c_int32, c_int32, c_int64]),
    "gngf_count_distinct_f32": (c_int, [_P, c_int64, c_int32, c_int32, c_int32, c_int64, _P, _P, _P, _P]),
    "gngf_count_distinct_i64": (c_int, [_P, c_int64, c_int32, c_int32, c_int32, c_int64, _P, _P, _P, _P]),
    "gngf_softmax_topk_fwd": (c_int, [_P, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P]),
    "gngf_topk_fwd": (c_int, [_P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "gngf_topk_bwd": (c_int, [_P, _P, c_int64, c_int64, c_int32, _P, _P]),
    "gngf_node_features_fwd": (c_int, [Lattice, Tables, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "gngf_encode_fwd": (c_int, [_P, c_int64, Lattice, c_int32, _P, _P, _P, _P, _P, _P]),
    "gngf_cell_to_node_counts": (c_int, [Lattice, _P, _P, _P]),
    "gngf_encode_hash_fwd": (c_int, [_P, c_int64, Lattice, Tables, c_int64, c_int32, _P, _P, _P]),
    "gngf_lattice_colsum": (c_int, [Lattice, _P, _P, c_int64, _P, _P]),
    "gngf_lattice_gather_rows": (c_int, [_P, c_int64, Lattice, _P, c_int64, _P, _P]),
    "gngf_lattice_gather_rows_i64": (c_int, [_P, c_int64, Lattice, _P, c_int64, _P, _P]),
    "gngf_lattice_scatter_rows": (c_int, [_P, c_int64, Lattice, _P, c_int64, _P, _P]),
    "gngf_encode_bwd": (c_int, [_P, c_int64, Lattice, c_int32, _P, _P, _P]),
    "gngf_node_features_bwd": (c_int, [Lattice, Tables, Tables, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P]),
    "gngf_encode_hash_bwd": (c_int, [_P, c_int64, Lattice, Tables, c_int64, c_int32, _P, _P]),
    "gngf_hpd_dlogits": (c_int, [Lattice, _P, c_int64, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Loads the shared library (once) and installs the prototypes.  Raises GngfError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GngfError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C collision_handling_in_instantngp_b200/csrc`). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError == ABI mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().gngf_strerror(status).decode()
        raise GngfError(f"{what or 'gngf call'} failed: {msg} ({status})")


PROFILER = None   # bench.py installs an object with .record(name, fn, args) to time every call with CUDA events


def call(name: str, *args) -> None:
    """Calls an int-returning entry point and raises on a non-zero status."""
    fn = getattr(load(), name)
    if PROFILER is None:
        check(fn(*args), name)
    else:
        check(PROFILER.record(name, fn, args), name)


def launch_count() -> int:
    return int(load().gngf_launch_count())


def ptr_array(tensors):
    """Host array of device pointers (for the `const float* const*` parameters)."""
    arr = (c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def int_array(values):
    return (c_int32 * len(values))(*[int(v) for v in values])


def make_tables(tensors) -> Tables:
    t = Tables()
    for i, x in enumerate(tensors):
        t.ptr[i] = x.data_ptr()
    return t
