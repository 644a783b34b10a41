"""Builds and loads ``libiou_b200.so`` (the C-ABI CUDA library, include/iou_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised.  The library is built in-tree with nvcc for sm_100a
only (``-gencode arch=compute_100a,code=sm_100a``).
"""
import ctypes
import glob
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libiou_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "iou_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]

MAX_LEVELS, MAX_ANCHORS = 8, 16
CONV_MAX_SEG, CONV_MAX_TAPS, CONV_MAX_SRC = 8, 9, 4
OUT_PADDED, OUT_DENSE = 0, 1
RES_NONE, RES_SAME, RES_UPSAMPLE2 = 0, 1, 2


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libiou_b200.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.isfile(nvcc):
        nvcc = "nvcc"
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose:
            print(out)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    subprocess.check_call(cmd)
    return LIB_PATH


DECODE_DELTA, DECODE_DISTANCE = 0, 1      # iou_postproc_cfg.decode_mode
DTYPE_F32, DTYPE_F16, DTYPE_F64 = 0, 1, 2   # iou_sigmoid_focal_loss_*_dtype


class PostprocCfg(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("num_anchors", ctypes.c_int32),
                ("num_classes", ctypes.c_int32), ("nms_pre", ctypes.c_int32),
                ("max_per_img", ctypes.c_int32),
                ("feat_h", ctypes.c_int32 * MAX_LEVELS), ("feat_w", ctypes.c_int32 * MAX_LEVELS),
                ("stride", ctypes.c_int32 * MAX_LEVELS),
                ("base_anchors", ((ctypes.c_float * 4) * MAX_ANCHORS) * MAX_LEVELS),
                ("target_means", ctypes.c_float * 4), ("target_stds", ctypes.c_float * 4),
                ("alpha", ctypes.c_float), ("score_thr", ctypes.c_float), ("iou_thr", ctypes.c_float),
                ("wh_ratio_clip", ctypes.c_float), ("decode_mode", ctypes.c_int32)]


class ConvSegment(ctypes.Structure):
    _fields_ = [("row_start", ctypes.c_int32), ("n_img", ctypes.c_int32), ("h", ctypes.c_int32),
                ("w", ctypes.c_int32)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("cout_pad", ctypes.c_int32),
                ("block_n", ctypes.c_int32), ("num_taps", ctypes.c_int32),
                ("tap_src", ctypes.c_int32 * CONV_MAX_TAPS), ("tap_dy", ctypes.c_int32 * CONV_MAX_TAPS),
                ("tap_dx", ctypes.c_int32 * CONV_MAX_TAPS), ("num_src", ctypes.c_int32),
                ("src", ctypes.c_void_p * CONV_MAX_SRC), ("src_rows", ctypes.c_int64),
                ("weight", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("relu", ctypes.c_int32), ("res_mode", ctypes.c_int32), ("residual", ctypes.c_void_p),
                ("res_seg", ConvSegment * CONV_MAX_SEG), ("out_mode", ctypes.c_int32),
                ("out", ctypes.c_void_p), ("out_dense", ctypes.c_void_p * CONV_MAX_SEG),
                ("dense_split", ctypes.c_int32), ("out_dense2", ctypes.c_void_p * CONV_MAX_SEG),
                ("num_seg", ctypes.c_int32), ("seg", ConvSegment * CONV_MAX_SEG),
                ("passes", ctypes.c_int32), ("out_rows", ctypes.c_int64), ("res_rows", ctypes.c_int64),
                ("diag_k", ctypes.c_int32), ("two_cta", ctypes.c_int32),
                ("phase_out", ctypes.c_void_p * 4), ("phase_only", ctypes.c_int32),
                ("src_cin", ctypes.c_int32 * CONV_MAX_SRC), ("k_split", ctypes.c_int32), ("wide", ctypes.c_int32),
                ("group_max_out", ctypes.c_void_p * CONV_MAX_SEG), ("group_max_cols", ctypes.c_int32)]


_SIGS = {
    "iou_last_error": (ctypes.c_char_p, []),
    "iou_abi_version": (ctypes.c_int, []),
    "iou_sizeof": (ctypes.c_size_t, [ctypes.c_int]),
    "iou_postproc_num_candidates": (ctypes.c_int, [ctypes.POINTER(PostprocCfg)]),
    "iou_postproc_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(PostprocCfg), ctypes.c_int]),
    "iou_decode_candidates": (ctypes.c_int, [ctypes.POINTER(PostprocCfg), ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "iou_batched_nms": (ctypes.c_int, [ctypes.POINTER(PostprocCfg), ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "iou_get_bboxes": (ctypes.c_int, [ctypes.POINTER(PostprocCfg), ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_void_p]),
    "iou_get_bboxes_premax": (ctypes.c_int, [ctypes.POINTER(PostprocCfg), ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_void_p]),
    "iou_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "iou_nms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "iou_soft_nms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "iou_batched_soft_nms": (ctypes.c_int, [ctypes.POINTER(PostprocCfg), ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_size_t, ctypes.c_void_p]),
    "iou_sigmoid_focal_loss_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                      ctypes.c_void_p, ctypes.c_void_p]),
    "iou_sigmoid_focal_loss_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                       ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "iou_sigmoid_focal_loss_forward_dtype": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                            ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                            ctypes.c_void_p, ctypes.c_void_p]),
    "iou_sigmoid_focal_loss_backward_dtype": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                             ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "iou_conv_plan_create": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_void_p)]),
    "iou_conv_chain_plan_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "iou_conv_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "iou_conv_plan_destroy": (None, [ctypes.c_void_p]),
    "iou_conv_plan_flops": (ctypes.c_double, [ctypes.c_void_p]),
    "iou_conv_plan_epilogue_warps": (ctypes.c_int, [ctypes.c_void_p]),
    "iou_pack_nchw": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "iou_unpack_nchw": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "iou_pack_nchw_fmt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "iou_unpack_nchw_fmt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "iou_stem_pack_v_fmt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "iou_maxpool3x3s2_fmt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "iou_im2col_stem": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "iou_stem_pack_v": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "iou_stem_pack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "iou_maxpool3x3s2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "iou_preprocess_u8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "iou_preprocess_resize_u8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p]),
    "iou_group_norm_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "iou_group_norm_relu": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "iou_group_norm_relu_fmt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "iou_sum_channel_groups": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p]),
    "iou_range_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "iou_scale_exp": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_void_p]),
    "iou_phase_split": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "iou_phase_split_fmt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
}
EXPORTED_SYMBOLS = sorted(_SIGS)

_lib = None
launch_count = 0      # kernels launched through this module (bench.py reports it)


def load():
    """dlopen the library; raises if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("libiou_b200.so is missing: run __graft_entry__.build() "
                               "(there is no CPU/PyTorch fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        raise RuntimeError("libiou_b200: %s (code %d)" % (load().iou_last_error().decode(), code))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
