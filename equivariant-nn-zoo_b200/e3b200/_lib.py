"""ctypes binding of libe3b200.so (the C ABI declared in include/e3b200.h).

The product path fails loudly: if the library is missing or a call returns a non-zero status a
RuntimeError is raised -- there is no CPU or eager-PyTorch fallback for the kernels."""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "lib", "libe3b200.so")

E3B_MAX_BLOCKS = 16
E3B_MAX_PATHS = 96

c_int, c_i32, c_i64, c_f32, c_f64, c_vp = (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                                           ctypes.c_double, ctypes.c_void_p)


class TpDesc(ctypes.Structure):
    _fields_ = [("mul", c_i32), ("n_in", c_i32), ("in_l", c_i32 * E3B_MAX_BLOCKS), ("in_p", c_i32 * E3B_MAX_BLOCKS),
                ("n_sh", c_i32), ("sh_l", c_i32 * E3B_MAX_BLOCKS), ("sh_p", c_i32 * E3B_MAX_BLOCKS), ("n_paths", c_i32),
                ("path_in", c_i32 * E3B_MAX_PATHS), ("path_sh", c_i32 * E3B_MAX_PATHS),
                ("path_lout", c_i32 * E3B_MAX_PATHS), ("path_slot", c_i32 * E3B_MAX_PATHS),
                ("w3j_sign_preset", c_i32)]


class GateDesc(ctypes.Structure):
    _fields_ = [("n_scalar_blocks", c_i32), ("scalar_mul", c_i32 * E3B_MAX_BLOCKS),
                ("scalar_act", c_i32 * E3B_MAX_BLOCKS), ("scalar_cst", c_f64 * E3B_MAX_BLOCKS),
                ("n_gated_blocks", c_i32), ("gated_mul", c_i32 * E3B_MAX_BLOCKS), ("gated_l", c_i32 * E3B_MAX_BLOCKS),
                ("gate_act", c_i32 * E3B_MAX_BLOCKS), ("gate_cst", c_f64 * E3B_MAX_BLOCKS)]


class GemmProblem(ctypes.Structure):
    _fields_ = [("A", c_vp), ("a_s1", c_i64), ("a_s2", c_i64), ("a_d", c_i32), ("B_packed", c_vp), ("C", c_vp),
                ("c_s1", c_i64), ("c_s2", c_i64), ("c_s3", c_i64), ("c_d", c_i32), ("aux", c_vp), ("aux_ld", c_i64),
                ("aux_d", c_i32), ("V", c_i32), ("aux_cols", c_i32), ("H", c_vp), ("h_ld", c_i64), ("M", c_i32), ("N", c_i32), ("K", c_i32),
                ("epilogue", c_i32), ("accumulate", c_i32), ("alpha", c_f32), ("act_cst", c_f32), ("row_map", c_vp),
                ("b_sel", c_vp), ("n_blocks", c_vp), ("b_set_stride", c_i64)]


class GemmPackDesc(ctypes.Structure):
    _fields_ = [("src", c_vp), ("s1", c_i64), ("s2", c_i64), ("sk", c_i64), ("d", c_i32), ("n2_valid", c_i32), ("N", c_i32),
                ("K", c_i32), ("dst", c_vp), ("n_sets", c_i32), ("set_stride", c_i64)]


class WgradProblem(ctypes.Structure):
    _fields_ = [("A", c_vp), ("a_s1", c_i64), ("a_s2", c_i64), ("a_d", c_i32), ("B", c_vp), ("b_s1", c_i64), ("b_s2", c_i64),
                ("b_d", c_i32), ("aux", c_vp), ("aux_ld", c_i64), ("aux_d", c_i32), ("V", c_i32), ("C", c_vp), ("c_s1", c_i64),
                ("c_s2", c_i64), ("c_s3", c_i64), ("c_d", c_i32), ("R", c_i64), ("K1", c_i32), ("K2", c_i32), ("alpha", c_f32),
                ("accumulate", c_i32)]


class PairCriteriaStruct(ctypes.Structure):
    _fields_ = [("segment", c_vp), ("max_separation", c_i64), ("p_random", c_f32), ("uniforms", c_vp), ("pair_ptr", c_vp),
                ("seed", ctypes.c_uint64)]


E3B_GEMM_MAX_GROUP = 16
E3B_WGRAD_MAX_GROUP = 8
CELL_GRID_BYTES = 48

# name -> (restype, argtypes); must list EVERY symbol include/e3b200.h declares
SIGNATURES = {
    "e3b_abi_version": (c_int, []),
    "e3b_last_error": (ctypes.c_char_p, []),
    "e3b_struct_size": (c_i64, [c_int]),
    "e3b_radius_graph_count": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, c_vp, c_vp]),
    "e3b_radius_graph_fill": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "e3b_radius_graph_fill_padded": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_i64, c_f32, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp,
                                             c_vp]),
    "e3b_pair_graph_count": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, ctypes.POINTER(PairCriteriaStruct), c_vp, c_vp]),
    "e3b_pair_graph_fill": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, ctypes.POINTER(PairCriteriaStruct), c_vp, c_i64, c_vp,
                                    c_vp]),
    "e3b_cell_graph_bin": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "e3b_cell_graph_sort": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "e3b_cell_graph_count": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "e3b_cell_graph_fill": (c_int, [c_vp, c_i64, c_vp, c_i32, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "e3b_csr_fill": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "e3b_edge_vectors_fwd": (c_int, [c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "e3b_edge_vectors_bwd": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "e3b_sh_fwd": (c_int, [c_int, c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "e3b_sh_bwd": (c_int, [c_int, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "e3b_radial_fwd": (c_int, [c_int, c_vp, c_i64, c_vp, c_int, c_f64, c_f64, c_int, c_int, c_f64, c_vp, c_vp]),
    "e3b_radial_bwd_blocks": (c_i64, [c_i64]),
    "e3b_radial_bwd": (c_int, [c_int, c_vp, c_vp, c_i64, c_vp, c_int, c_f64, c_f64, c_int, c_int, c_f64, c_vp, c_vp, c_vp]),
    "e3b_tp_plan_create": (c_int, [ctypes.POINTER(TpDesc), ctypes.POINTER(c_vp)]),
    "e3b_tp_plan_destroy": (None, [c_vp]),
    "e3b_tp_plan_is_specialized": (c_int, [c_vp]),
    "e3b_tp_plan_dims": (c_int, [c_vp] + [ctypes.POINTER(c_i32)] * 5),
    "e3b_tpconv_fwd": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "e3b_tpconv_bwd": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "e3b_tpconv_bwd_nodes": (c_int, [c_vp, c_i64, c_i64] + [c_vp] * 12),
    "e3b_tpconv_fwd_shared": (c_int, [c_vp, c_i64, c_i64] + [c_vp] * 9),
    "e3b_pair_sum_act": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_i32, c_vp, c_vp]),
    "e3b_segment_sum": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "e3b_gate_fwd": (c_int, [ctypes.POINTER(GateDesc), c_int, c_vp, c_i64, c_vp, c_vp]),
    "e3b_gate_bwd": (c_int, [ctypes.POINTER(GateDesc), c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "e3b_gate_bwd2": (c_int, [ctypes.POINTER(GateDesc), c_int, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "e3b_gate_imu_fwd": (c_int, [ctypes.POINTER(GateDesc), c_int, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "e3b_gate_imu_bwd": (c_int, [ctypes.POINTER(GateDesc), c_int, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "e3b_gemm_tile_n": (c_int, [c_i32, c_i32]),
    "e3b_gemm_packed_floats": (c_i64, [c_i32, c_i32]),
    "e3b_gemm_pack": (c_int, [ctypes.POINTER(GemmPackDesc), c_i32, c_vp]),
    "e3b_gemm_run": (c_int, [ctypes.POINTER(GemmProblem), c_i32, c_vp]),
    "e3b_wgrad_workspace_floats": (c_i64, [ctypes.POINTER(WgradProblem), c_i32]),
    "e3b_wgrad_run": (c_int, [ctypes.POINTER(WgradProblem), c_i32, c_vp, c_vp]),
    "e3b_layernorm_fwd": (c_int, [c_int, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_f64, c_vp, c_vp, c_vp]),
    "e3b_layernorm_bwd_blocks": (c_i64, [c_i64]),
    "e3b_layernorm_bwd": (c_int, [c_int, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "e3b_adam_ema_step": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f64, c_f64, c_f32, c_f32, c_i64, c_f32, c_vp,
                                  c_vp, c_vp, c_vp, c_vp]),
    "e3b_layout_convert": (c_int, [c_int, c_vp, c_i64, c_i32, c_vp, c_vp, c_int, c_vp, c_vp]),
}

_lib = None


def load():
    """Loads libe3b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The B200 path has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if lib.e3b_abi_version() != 1:
            raise RuntimeError("libe3b200.so ABI version mismatch")
        for which, st in enumerate((TpDesc, GateDesc, GemmProblem, GemmPackDesc, WgradProblem)):
            if lib.e3b_struct_size(which) != ctypes.sizeof(st):
                raise RuntimeError(f"libe3b200.so struct layout mismatch for {st.__name__}")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libe3b200 error {rc}: {load().e3b_last_error().decode()}")


def ptr(t):
    """device pointer of a tensor (None -> NULL); the tensor must be contiguous"""
    if t is None:
        return None
    assert t.is_contiguous(), "libe3b200 needs contiguous tensors"
    return t.data_ptr()


def stream():
    """raw cudaStream_t of torch's current stream (torch.cuda.current_stream() costs ~15 us per call)"""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def dtype_code(t):
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.float64:
        return 1
    raise TypeError(f"libe3b200 supports float32/float64, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libe3b200 kernels need CUDA tensors (the B200 path has no CPU fallback)")


# number of kernel launches issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0


def count_launch(n=1):
    global launch_count
    launch_count += n
