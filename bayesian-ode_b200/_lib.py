"""ctypes binding of libbode_b200.so (the C ABI declared in include/bode_b200.h).

PyTorch is used only for device memory and streams: tensors cross the boundary as raw
device pointers.  There is no CPU fallback -- if the library is missing or a call fails
this module raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbode_b200.so")

BODE_EULER, BODE_MIDPOINT, BODE_RK4 = 0, 1, 2
METHODS = {"euler": BODE_EULER, "midpoint": BODE_MIDPOINT, "rk4": BODE_RK4}
GRAD_DISCRETE, GRAD_ADJOINT = 0, 1


class BodeError(RuntimeError):
    pass


class NpdeFieldStruct(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("m", C.c_int32), ("grid_mx", C.c_int32), ("grid_my", C.c_int32),
        ("gx", C.c_double * 32), ("gy", C.c_double * 32), ("ell", C.c_double * 2),
        ("Z", C.c_void_p), ("A", C.c_void_p), ("Ksym", C.c_void_p), ("U", C.c_void_p), ("U_stride", C.c_int64),
        ("AT", C.c_void_p),
    ]


class MlpFieldStruct(C.Structure):
    _fields_ = [("P", C.c_int32), ("H", C.c_int32), ("theta", C.c_void_p), ("theta_stride", C.c_int64)]


class Dopri5Opts(C.Structure):
    _fields_ = [("t", C.c_void_p), ("rtol", C.c_double), ("atol", C.c_double), ("safety", C.c_double), ("ifactor", C.c_double),
                ("dfactor", C.c_double), ("max_num_steps", C.c_int32), ("user_first_step", C.c_int32), ("stats", C.c_void_p),
                ("controller", C.c_int32), ("n_groups", C.c_int32), ("group_end", C.c_int32 * 4)]


class GridStruct(C.Structure):
    _fields_ = [
        ("S", C.c_int32), ("T", C.c_int32), ("sign", C.c_float),
        ("dt", C.c_void_p), ("obs_ptr", C.c_void_p), ("adj_dt", C.c_void_p), ("adj_ptr", C.c_void_p),
    ]


_lib = None

# every symbol include/bode_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "bode_version": (C.c_int, []),
    "bode_last_error": (C.c_char_p, []),
    "bode_device_sm_count": (C.c_int, []),
    "bode_peak_kernel": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "bode_stamp": (C.c_int, [_P, _P]),
    "bode_npde_set_row_kernel": (C.c_int, [C.c_int32]),
    "bode_npde_scratch_floats": (C.c_size_t, [C.c_int32] * 6),
    "bode_npde_scratch_floats_m": (C.c_size_t, [C.c_int32] * 7),
    "bode_npde_odeint": (C.c_int, [C.POINTER(NpdeFieldStruct), C.POINTER(GridStruct), C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "bode_npde_odeint_backward": (C.c_int, [C.POINTER(NpdeFieldStruct), C.POINTER(GridStruct), C.c_int32, C.c_int32, C.c_int32,
                                             _P, C.c_int32, _P, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "bode_npde_nlp_grad": (C.c_int, [C.POINTER(NpdeFieldStruct), C.POINTER(GridStruct), C.c_int32, C.c_int32, C.c_int32,
                                      _P, C.c_int32, _P, _P, C.c_int64, C.c_float, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64,
                                      _P, C.c_size_t, _P]),
    "bode_mlp_odeint": (C.c_int, [C.POINTER(MlpFieldStruct), C.POINTER(GridStruct), C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "bode_npde_dopri5": (C.c_int, [C.POINTER(NpdeFieldStruct), C.POINTER(Dopri5Opts), C.c_int32, C.c_float, C.c_int32, _P, C.c_int32, _P, _P]),
    "bode_mlp_dopri5": (C.c_int, [C.POINTER(MlpFieldStruct), C.POINTER(Dopri5Opts), C.c_int32, C.c_float, C.c_int32, _P, C.c_int32, _P, _P]),
    "bode_dopri5_scratch_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bode_npde_dopri5_backward": (C.c_int, [C.POINTER(NpdeFieldStruct), C.POINTER(Dopri5Opts), C.c_int32, C.c_float, C.c_int32, _P, C.c_int32,
                                             _P, _P, C.c_int64, _P, _P, C.c_size_t, C.c_int32, _P]),
    "bode_npde_dopri5_nlp_grad": (C.c_int, [C.POINTER(NpdeFieldStruct), C.POINTER(Dopri5Opts), C.c_int32, C.c_float, C.c_int32, _P, C.c_int32,
                                             _P, _P, C.c_int64, C.c_float, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64, _P, C.c_size_t,
                                             C.c_int32, _P]),
    "bode_mlp_dopri5_backward": (C.c_int, [C.POINTER(MlpFieldStruct), C.POINTER(Dopri5Opts), C.c_int32, C.c_float, C.c_int32, _P, C.c_int32,
                                            _P, _P, C.c_int64, _P, _P, C.c_size_t, C.c_int32, _P]),
    "bode_mlp_dopri5_sse_grad": (C.c_int, [C.POINTER(MlpFieldStruct), C.POINTER(Dopri5Opts), C.c_int32, C.c_float, C.c_int32, _P, C.c_int32,
                                            _P, C.c_float, C.c_float, C.c_float, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_size_t, C.c_int32, _P]),
    "bode_mlp_odeint_backward": (C.c_int, [C.POINTER(MlpFieldStruct), C.POINTER(GridStruct), C.c_int32, C.c_int32, C.c_int32,
                                            _P, C.c_int32, _P, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "bode_mlp_sse_grad": (C.c_int, [C.POINTER(MlpFieldStruct), C.POINTER(GridStruct), C.c_int32, C.c_int32, C.c_int32,
                                     _P, C.c_int32, _P, C.c_float, C.c_float, C.c_float, C.c_int32, _P, _P, _P, C.c_int64,
                                     _P, C.c_size_t, _P]),
    "bode_sgld_step": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_int32, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "bode_psgld_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "bode_asghmc_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "bode_hamcmc_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bode_hamcmc_step": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P, C.c_int64, _P,
                                    C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, _P, _P]),
    "bode_hamcmc_contig_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bode_hamcmc_contig_step": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P,
                                          C.c_int64, _P, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_uint64, C.c_uint32, _P, _P]),
    "bode_axpy": (C.c_int, [_P, _P, C.c_float, C.c_int64, _P, _P, _P]),
    "bode_sampler_schedule": (C.c_int, [_P, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_uint32, _P]),
    "bode_fill_normal": (C.c_int, [_P, C.c_int64, C.c_uint64, C.c_uint32, _P]),
    "bode_npde_set_lanes_per_pair": (C.c_int, [C.c_int32]),
    "bode_npde_set_cta_limit": (C.c_int, [C.c_int32]),
    "bode_mala_accept": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_float,
                                   C.c_int32, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "bode_svgd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "bode_svgd_set_tensor_cores": (C.c_int, [C.c_int32]),
    "bode_svgd_sqdist": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _P, C.c_size_t,
                                    C.POINTER(C.c_void_p), _P]),
    "bode_svgd_workspace_init": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.c_size_t, _P]),
    "bode_svgd_window_table": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "bode_svgd_window_select": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "bode_mlp_set_tensor_cores": (C.c_int, [C.c_int32]),
    "bode_svgd_window_disarm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "bode_svgd_radix_fallback": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "bode_svgd_hist_pass": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "bode_svgd_select_digit": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "bode_svgd_gamma": (C.c_int, [C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "bode_svgd_phi": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int64, _P, C.c_int64, C.c_float, C.c_int32, C.c_int32, C.c_int32, _P, _P,
                                 _P, C.c_int64, _P, C.c_int64, C.c_float, _P]),
    "bode_svgd_staged_supported": (C.c_int, [C.c_int32, C.c_int32]),
    "bode_svgd_set_gram_split": (C.c_int, [C.c_int32]),
    "bode_svgd_d2_tiled": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "bode_svgd_peer_gather": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.POINTER(C.c_void_p), C.c_void_p]),
    "bode_svgd_peer_status": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "bode_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "bode_peer_free": (C.c_int, [_P]),
    "bode_peer_export": (C.c_int, [_P, _P]),
    "bode_peer_import": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "bode_peer_release": (C.c_int, [_P]),
    "bode_svgd_set_peers": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "bode_svgd_sqdist_staged": (C.c_int, [C.c_int32, _P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_uint64,
                                           _P, C.c_size_t, C.POINTER(C.c_void_p), _P]),
    "bode_svgd_phi_staged": (C.c_int, [C.c_int32, _P, C.c_int64, C.c_int32, _P, C.c_int64, _P, C.c_int64, C.c_float, C.c_int32, C.c_int32,
                                        C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64, C.c_float, _P]),
    "bode_svgd_set_select_ctas": (C.c_int, [C.c_int32]),
    "bode_svgd_arm_score_tiles": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.c_float]),
    "bode_svgd_disarm_score_tiles": (C.c_int, []),
}
SVGD_PREPARE, SVGD_COMPUTE, SVGD_PREPARE_POSITIONS = 1, 2, 4


def load():
    """Load the shared library (once).  Raises BodeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("BODE_LIB_PATH", LIB_PATH)          # developer aid: load an experimental build of the same ABI
    if not os.path.exists(path):
        raise BodeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C bayesian-ode_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().bode_last_error().decode("utf-8", "replace")
        raise BodeError(f"libbode_b200 call failed (status {status}): {msg}")


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-contiguous tensor required at the C ABI"
    return C.c_void_p(t.data_ptr())


def rows(t, inner):
    """(pointer, row stride in floats) of a [P, ...] fp32 CUDA tensor whose trailing dims are dense with ``inner``
    elements per particle -- e.g. a column block of a flat theta[P, d] buffer."""
    assert t.is_cuda and t.dtype == torch.float32
    if t.dim() >= 2 and t[0].is_contiguous() and t[0].numel() == inner:
        return C.c_void_p(t.data_ptr()), int(t.stride(0)) if t.shape[0] > 1 else max(int(t.stride(0)), inner)
    raise BodeError("tensor rows must be dense (got shape %s strides %s)" % (tuple(t.shape), t.stride()))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise BodeError("bayesian_ode_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BodeError("expected CUDA tensors; the product has no CPU path")
