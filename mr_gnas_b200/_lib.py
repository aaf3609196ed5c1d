"""ctypes binding of libmrgnas.so (include/mrgnas.h).  No torch types cross the boundary:
tensors are passed as raw device pointers + sizes, the stream as a cudaStream_t handle.

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised (build it with ``python -m mr_gnas_b200.build``)."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmrgnas.so")

COMP = {"pre_sub": 0, "pre_mult": 1, "pre_add": 2, "sub": 0, "mult": 1, "add": 2}
RED = {"sum": 0, "mean": 1, "max": 2}
CHUNK_ROWS = 32


class MrgAct(Structure):
    _fields_ = [("data", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("relu", c_int32)]


class MrgGrad(Structure):
    _fields_ = [("ds", c_void_p), ("y", MrgAct), ("coef", c_void_p)]


class MrgGateParams(Structure):
    _fields_ = [("W", c_void_p * 3), ("b", c_void_p * 3), ("a", c_void_p * 3)]


class MrgGateGrads(Structure):
    _fields_ = [("dW", c_void_p * 3), ("db", c_void_p * 3), ("da", c_void_p * 3)]


class MrgActList(Structure):
    _fields_ = [("acts", MrgAct * 8), ("n", c_int32)]


P, I32, I64, F32, SZ = c_void_p, c_int32, c_int64, c_float, c_size_t

_SIGNATURES = {
    "mrg_abi_version": (c_int32, []),
    "mrg_last_error": (c_char_p, []),
    "mrg_graph_workspace_bytes": (SZ, [I64, I64, I64]),
    "mrg_graph_build": (I32, [P, P, P, I64, I64, I64, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "mrg_graph_build_part": (I32, [P, P, P, I64, I64, I64, I64, I64, P, P, P, P, P, P, P, SZ, P]),
    "mrg_edge_norm": (I32, [P, P, P, I64, P, P]),
    "mrg_chunk_capacity": (I64, [I64, I64]),
    "mrg_chunk_workspace_bytes": (SZ, [I64]),
    "mrg_chunk_build": (I32, [P, I64, P, P, P, SZ, P]),
    "mrg_stats_nparts": (I32, [I64]),
    "mrg_stats_max_parts": (I32, []),
    "mrg_colstats": (I32, [MrgAct, I64, I32, P, P]),
    "mrg_bn_finalize": (I32, [P, I32, I64, I32, P, P, F32, F32, P, P, P, P, P, P, P]),
    "mrg_affine_act": (I32, [MrgAct, I64, I32, P, P]),
    "mrg_bn_bwd_reduce": (I32, [P, MrgAct, I64, I32, P, P]),
    "mrg_bn_bwd_finalize": (I32, [P, I32, I64, I32, P, P, P, P, P, P, P]),
    "mrg_bn_bwd_apply": (I32, [P, MrgAct, P, I64, I32, P, I32, P]),
    "mrg_compose_fwd": (I32, [P, P, P, P, I64, I32, I32, P, P, P]),
    "mrg_compose_bwd_rows": (I32, [P, P, P, I64, I32, I32, P, P, P]),
    "mrg_sparse_gate_fwd": (I32, [MrgAct, MrgAct, I64, I32, P, P, P, P, F32, P, P, P, P]),
    "mrg_gate_collapse_fwd": (I32, [MrgGateParams, I32, I32, I32, I32, P, P, P, P]),
    "mrg_gate_collapse_bwd": (I32, [MrgGateParams, I32, I32, I32, I32, P, P, P, MrgGateGrads, P]),
    "mrg_gate_dparam_count": (I64, [I64, I32]),
    "mrg_sparse_gate_bwd": (I32, [P, MrgAct, MrgAct, P, I64, I32, P, P, P, F32, P, P, I32, P, P]),
    "mrg_sparse_gate_bwd_finalize": (I32, [P, I64, I32, P, P, P, P]),
    "mrg_sparse_gate_bwd_fused_supported": (I32, [I32]),
    "mrg_sparse_gate_bwd_fused": (I32, [MrgGrad, MrgAct, MrgAct, P, I64, I32, P, P, P, F32, P, P, I32, P, P, P]),
    "mrg_dense_gate_fwd": (I32, [P, MrgAct, I64, I32, I32, P, F32, P, P, P]),
    "mrg_dense_gate_bwd": (I32, [P, P, MrgAct, I64, I32, I32, P, F32, P, P, I32, P]),
    "mrg_mixed_sum_fwd": (I32, [MrgActList, P, I64, I32, P, P]),
    "mrg_mixed_bwd_scale": (I32, [P, P, P, P, P, P, P, P, I32, P, I32, I32, P]),
    "mrg_seg_reduce_workspace_bytes": (SZ, [I64, I32, I32]),
    "mrg_seg_reduce_fwd": (I32, [I32, MrgAct, P, P, P, P, I64, I64, I32, P, P, F32, MrgAct, I32, P, P, P, SZ, P]),
    "mrg_seg_reduce_bwd": (I32, [I32, P, P, P, MrgAct, P, P, I64, I64, I32, P, I32, P]),
    "mrg_amax_tc_supported": (I32, [I32]),
    "mrg_amax_tc_workspace_bytes": (SZ, [I64, I32]),
    "mrg_amax_tc_fwd": (I32, [MrgAct, P, P, P, P, I64, I64, I32, MrgAct, P, P, P, SZ, P]),
    "mrg_amax_tc_fwd_bf16": (I32, [MrgAct, P, P, P, P, I64, I64, I32, MrgAct, P, P, P, SZ, P]),
    "mrg_amax_bwd_workspace_bytes": (SZ, [I64, I64, I32]),
    "mrg_amax_bwd": (I32, [P, P, MrgAct, P, P, P, P, P, P, I64, I64, I64, I32, P, P, P, P, SZ, P]),
    "mrg_debug_set_dw_prof": (I32, [P]),
    "mrg_labels_from_csr": (I32, [P, P, I64, I64, I64, F32, F32, P, P]),
    "mrg_filtered_rank": (I32, [P, P, P, I64, I64, P, P]),
    "mrg_distmult_bce_supported": (I32, [I32]),
    "mrg_distmult_bce_nparts": (I32, [I64]),
    "mrg_distmult_bce_workspace_bytes": (SZ, [I32]),
    "mrg_distmult_bce_fwd": (I32, [P, P, P, I64, I64, I32, P, P, P, P, SZ, P]),
    "mrg_linear_tc_supported": (I32, [I32]),
    "mrg_linear_tc_workspace_bytes": (SZ, [I32]),
    "mrg_linear_tc_fwd": (I32, [P, P, P, I64, I32, I32, P, I64, P, SZ, P]),
    "mrg_mixed_pre_stats": (I32, [P, P, I64, I32, P, I32, P, P]),
    "mrg_mixed_pre_fwd": (I32, [P, P, I64, I32, P, I32, P, P, P, P, P]),
    "mrg_mixed_pre_bwd_stats": (I32, [P, P, P, I64, I32, P, I32, P, P, P, P]),
    "mrg_mixed_pre_bwd": (I32, [P, P, P, I64, I32, P, I32, P, P, P, P, P, P]),
    "mrg_gemm_red_workspace_bytes": (SZ, [I64, I32, I32]),
    "mrg_gemm_red": (I32, [P, I64, I32, P, I64, I64, I32, I32, P, I64, P, P, P, SZ, P]),
    "mrg_transe_fwd": (I32, [P, P, I64, I64, I32, F32, P, P]),
    "mrg_transe_bwd_workspace_bytes": (SZ, [I64, I64, I32]),
    "mrg_transe_bwd": (I32, [P, P, P, I64, I64, I32, P, P, P, SZ, P]),
    "mrg_bce_nparts": (I32, [I64]),
    "mrg_sigmoid_bce_fwd": (I32, [P, P, I64, P, P, P, P]),
    "mrg_sigmoid_bce_bwd": (I32, [P, P, I64, P, P, P]),
}
# optional tensor-core entry points (declared in mrgnas.h once built)
_OPTIONAL = {}

_lib = None
# kernels enqueued per C-ABI call (default 1); used for the gpu_launches count bench.py reports
KERNELS_PER_CALL = {"mrg_amax_tc_fwd": 3, "mrg_amax_tc_fwd_bf16": 3, "mrg_amax_bwd": 5, "mrg_distmult_bce_fwd": 3, "mrg_linear_tc_fwd": 2, "mrg_gemm_red": 1, "mrg_seg_reduce_fwd": 2, "mrg_sigmoid_bce_fwd": 2, "mrg_transe_bwd": 3, "mrg_graph_build": 12, "mrg_graph_build_part": 10, "mrg_chunk_build": 3}
launch_count = 0   # libmrgnas kernels launched so far
_profile = None    # when a list: (name, start_event, end_event) per call (bench.py per-kernel timing)


def declared_symbols():
    return sorted(list(_SIGNATURES) + list(_OPTIONAL))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"libmrgnas.so not found at {LIB_PATH}: run `python -m mr_gnas_b200.build` "
                           "(the product path has no CPU / eager fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in {**_SIGNATURES, **_OPTIONAL}.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mrg_abi_version() != 1:
        raise RuntimeError("libmrgnas ABI version mismatch")
    _lib = lib
    return lib


def ptr(t):
    """Raw device pointer of a tensor; refuses anything the kernels cannot address."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libmrgnas got a CPU tensor: the message-passing path has no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("libmrgnas expects contiguous tensors")
    return c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """cudaStream_t of torch's current stream.  (torch.cuda.current_stream() costs ~16 us of Python per call --
    measured: 11 ms of a 51 ms launch-bound supernet step -- the raw getter is a single C call.)"""
    if _raw_stream is not None and _cur_device is not None:
        return c_void_p(_raw_stream(_cur_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def check_f32(*ts):
    for t in ts:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"libmrgnas expects contiguous fp32 CUDA tensors, got {t.dtype} {t.device} "
                               f"contiguous={t.is_contiguous()}")


def check_i32(*ts):
    for t in ts:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
            raise RuntimeError("libmrgnas expects contiguous int32 CUDA index tensors")


def act(data, scale=None, shift=None, relu=False):
    """mrg_act view of `data` read through relu?(scale*x+shift)."""
    if data is None:
        return MrgAct(None, None, None, 0)
    check_f32(data, scale, shift)
    return MrgAct(data.data_ptr(), scale.data_ptr() if scale is not None else None,
                  shift.data_ptr() if shift is not None else None, 1 if relu else 0)


def grad(ds, y_act=None, coef=None):
    """mrg_grad view: `ds` read through the lazy BN(+ReLU) backward of its state (coef from mrg_bn_bwd_finalize)."""
    check_f32(ds, coef)
    return MrgGrad(ds.data_ptr(), y_act if y_act is not None else MrgAct(None, None, None, 0),
                   coef.data_ptr() if coef is not None else None)


def call(name, *args, nbytes=None):
    """Enqueue one C-ABI call.  `nbytes` (optional) = compulsory HBM bytes of this call (inputs read once +
    outputs written once); only used by the profiler for roofline accounting."""
    global launch_count
    lib = load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.mrg_last_error().decode()}")
    if _profile is not None:
        e1.record()
        ints = tuple(a for a in args if type(a) is int)[:3]
        _profile.append((name + str(ints), e0, e1, nbytes))
    launch_count += KERNELS_PER_CALL.get(name, 1)
    return rc


def start_profile():
    global _profile
    _profile = []


def stop_profile():
    """-> {call name: (count, total ms, total compulsory bytes or None)} from CUDA events around each call."""
    global _profile
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, nb in _profile or []:
        c, t, b = out.get(name, (0, 0.0, 0))
        out[name] = (c + 1, t + e0.elapsed_time(e1), (b + nb) if (nb is not None and b is not None) else None)
    _profile = None
    return out
