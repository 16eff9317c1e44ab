"""ctypes binding of libdva_b200.so (the C ABI declared in include/dva_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  PyTorch is used above this layer only for device memory and streams.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVA_B200_LIB: developer knob to load a tuning variant of the same library (bench sweeps)
LIB_PATH = os.environ.get("DVA_B200_LIB") or os.path.join(_HERE, "libdva_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

DVA_OK, DVA_EINVAL, DVA_EALIGN, DVA_EUNSUPPORTED = 0, -1, -2, -3
DVA_F32, DVA_BF16, DVA_F16 = 0, 1, 2
REDUCE_CODES = {"sum": 0, "add": 0, "mean": 1, "max": 2, "min": 3}
DTYPE_CODES = {torch.float32: DVA_F32, torch.bfloat16: DVA_BF16, torch.float16: DVA_F16}

_vp, _i64, _i32, _f32, _f64, _sz = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_double, ctypes.c_size_t)

# name -> (restype, argtypes); mirrors include/dva_b200.h one to one
SIGNATURES = {
    "dva_abi_version": (_i32, []),
    "dva_last_error": (ctypes.c_char_p, []),
    "dva_launch_count": (_i64, []),
    "dva_segment_csr_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp]),
    "dva_segment_csr_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp]),
    "dva_gather_csr": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "dva_segment_softmax_csr_fwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _f32, _i32, _i32, _vp]),
    "dva_segment_softmax_csr_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp]),
    "dva_view_attention_fwd": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _i64, _i64, _i64, _i64, _i64, _i32, _f32, _i32, _vp]),
    "dva_view_attention_set_path": (_i32, [_i32]),
    "dva_view_attention_bwd_workspace_bytes": (_sz, [_i64]),
    "dva_view_attention_bwd": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i32, _i32,
                                      _vp, _sz, _vp]),
    "dva_qk_scores_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp]),
    "dva_qk_scores_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp]),
    "dva_heuristic_pool_fwd": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _i32,
                                      _i32, _vp]),
    "dva_gather_pool_fwd": (_i32, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                   _i64, _i64, _i32, _i32, _vp]),
    "dva_gather_pool_bwd": (_i32, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                   _i64, _i64, _i32, _i32, _vp]),
    "dva_interp_pool_fwd": (_i32, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                   _i64, _i64, _i64, _i64, _i32, _i32, _vp]),
    "dva_interp_pool_bwd": (_i32, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                   _i64, _i64, _i64, _i64, _i32, _i32, _vp]),
    "dva_transpose_last2": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "dva_knn_cell_ids": (_i32, [_vp, _vp, _i64, _f32, _f32, _f32, _f32, _i32, _i32, _i32, _vp]),
    "dva_knn_grid": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _i32, _i32, _i32,
                            _vp, _vp, _vp]),
    "dva_neighborhood_features": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _f64, _i32, _i32, _vp,
                                         _i64, _i64, _vp]),
    "dva_scatter_add_rows": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "dva_mapping_build_workspace_bytes": (_sz, [_i64, _i64]),
    "dva_mapping_build": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _sz, _vp]),
    "dva_view_cat_sorting": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "dva_linear_gemm_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "dva_linear_gemm": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _sz, _vp]),
    "dva_linear_bnstats_supported": (_i32, [_i64, _i64, _i64]),
    "dva_linear_bnstats_workspace_bytes": (_sz, [_i64, _i64]),
    "dva_linear_bnstats_fwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dva_bn_workspace_bytes": (_sz, [_i64, _i64]),
    "dva_bn_act_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _i32, _i32,
                              _vp, _sz, _vp]),
    "dva_bn_act_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _i32, _i32, _vp, _sz,
                              _vp]),
    "dva_mlp_layer_bwd_supported": (_i32, [_i64, _i64, _i64]),
    "dva_mlp_layer_bwd_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "dva_mlp_layer_bwd": (_i32, [_vp] * 11 + [_i64, _i64, _i64, _f32, _vp, _sz, _vp]),
    "dva_zbuffer_splat": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64,
                                 _i32, _vp]),
    "dva_splat_boxes": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _f64, _f64, _f64,
                               _i32, _f64, _f64, _vp]),
    "dva_splat_boxes_from_width": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "dva_project_equirectangular": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                           _i64, _f32, _f32, _vp]),
    "dva_project_camera": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _f32,
                                  _f32, _vp]),
    "dva_csr_pointers_from_sorted": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "dva_csr_select_values": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
}

_lib = None


def build(verbose=False):
    """Compile libdva_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libdva_b200.so failed")
    return LIB_PATH


def load():
    """Load the shared library (once). Raises if it has not been built -- no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C deepviewagg_b200/csrc`). deepviewagg_b200 has no CPU "
            f"fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.dva_abi_version() != 1:
        raise RuntimeError("libdva_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().dva_last_error().decode()


def launch_count():
    return int(load().dva_launch_count())


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "deepviewagg_b200 operators run on CUDA tensors only (sm_100a kernels, no CPU "
                "fallback); got a tensor on " + str(t.device))


def dtype_code(t):
    try:
        return DTYPE_CODES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported feature dtype {t.dtype}; expected float32/bfloat16/float16")
