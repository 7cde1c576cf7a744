"""ctypes binding of libhvpr_b200.so (the C ABI declared in include/hvpr_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhvpr_b200.so")

# every symbol include/hvpr_b200.h declares (tests/test_abi_and_host.py checks the .so exports each one)
SYMBOLS = [
    "hvpr_strerror", "hvpr_last_cuda_error", "hvpr_version", "hvpr_init",
    "hvpr_voxelize_workspace_bytes", "hvpr_voxelize", "hvpr_frame_offsets",
    "hvpr_pfn", "hvpr_pfn_pack", "hvpr_pfn_packed_bytes",
    "hvpr_mem_attn_workspace_bytes", "hvpr_mem_pack_bf16", "hvpr_mem_attn", "hvpr_mem_train_forward", "hvpr_mse_loss",
    "hvpr_bev_fill", "hvpr_build_cell_map",
    "hvpr_conv_packed_bytes", "hvpr_conv_pack_weights", "hvpr_conv2d", "hvpr_nchw_to_nhwc_bf16", "hvpr_attention_gate",
    "hvpr_bev_fill_nhwc_bf16", "hvpr_head_decode",
    "hvpr_post_process_workspace_bytes", "hvpr_post_process", "hvpr_boxes_iou3d",
]

OVERFLOW = {"continue": 0, "break": 1}
VOXELIZE_FORCE_HASH = 0x100
MEM_FP32, MEM_BF16_RESCORE = 0, 1


class HvprGeom(ctypes.Structure):
    _fields_ = [("lo", c_float * 3), ("vs", c_float * 3), ("grid", c_int32 * 3)]


class HvprPfnWeights(ctypes.Structure):
    _fields_ = [("w0", c_float * 160), ("b0", c_float * 16), ("w1a", c_float * 1024), ("w1b", c_float * 1024),
                ("b1", c_float * 64), ("ws0", c_float * 80), ("bs0", c_float * 16), ("ws1", c_float * 512),
                ("bs1", c_float * 32)]


class HvprLaunchCfg(ctypes.Structure):
    _fields_ = [("blocks_per_sm", c_int32), ("variant", c_int32)]


def launch_cfg(cfg):
    """(blocks_per_sm, variant) tuple or None -> pointer argument for hvpr_pfn / hvpr_bev_fill (None = library defaults)."""
    if cfg is None:
        return None
    return ctypes.byref(HvprLaunchCfg(int(cfg[0]), int(cfg[1])))


class HvprZeroFill(ctypes.Structure):
    _fields_ = [("ptr", c_void_p * 4), ("bytes", ctypes.c_uint64 * 4), ("n", c_int32)]


def zero_fill(tensors):
    """List of (contiguous, 16-byte aligned) tensors or None -> the zero_fill argument of hvpr_mem_attn."""
    if not tensors:
        return None
    z = HvprZeroFill()
    assert len(tensors) <= 4
    for i, t in enumerate(tensors):
        assert t.is_contiguous()
        z.ptr[i] = t.data_ptr()
        z.bytes[i] = t.numel() * t.element_size()
    z.n = len(tensors)
    return ctypes.byref(z)


class HvprConvArgs(ctypes.Structure):
    _fields_ = [("in_", c_void_p), ("n", c_int32), ("h_in", c_int32), ("w_in", c_int32), ("in_cs", c_int32),
                ("c_in", c_int32), ("ksize", c_int32), ("stride", c_int32), ("w_packed", c_void_p),
                ("n_total", c_int32), ("bn", c_int32), ("bias", c_void_p), ("relu", c_int32), ("gate", c_void_p),
                ("residual", c_void_p), ("res_cs", c_int32), ("out_mode", c_int32), ("out", c_void_p),
                ("out_cs", c_int32), ("out_c_off", c_int32), ("up", c_int32), ("c_out", c_int32), ("out_ctot", c_int32)]


class HvprError(RuntimeError):
    pass


_lib = None
_inited_devices = set()


def build(verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise HvprError("building libhvpr_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HvprError("%s not found — build it with `make -C hvpr_b200/csrc` or __graft_entry__.build(); "
                        "there is no CPU fallback for this path" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.hvpr_strerror.restype = c_char_p
    L.hvpr_strerror.argtypes = [c_int]
    L.hvpr_last_cuda_error.restype = c_char_p
    L.hvpr_version.restype = c_int
    L.hvpr_init.restype = c_int
    L.hvpr_voxelize_workspace_bytes.restype = c_size_t
    L.hvpr_voxelize_workspace_bytes.argtypes = [c_int64, c_int, c_void_p, c_int]
    L.hvpr_voxelize.restype = c_int
    L.hvpr_voxelize.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int64, c_void_p, c_int, c_int,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.hvpr_frame_offsets.restype = c_int
    L.hvpr_frame_offsets.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]
    L.hvpr_pfn.restype = c_int
    L.hvpr_pfn.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                           c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(HvprLaunchCfg), c_void_p]
    L.hvpr_pfn_pack.restype = c_int
    L.hvpr_pfn_pack.argtypes = [c_void_p, c_void_p, c_void_p]
    L.hvpr_pfn_packed_bytes.restype = c_size_t
    L.hvpr_mem_attn_workspace_bytes.restype = c_size_t
    L.hvpr_mem_attn_workspace_bytes.argtypes = [c_int64, c_int, c_int]
    L.hvpr_mem_pack_bf16.restype = c_int
    L.hvpr_mem_pack_bf16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.hvpr_mem_attn.restype = c_int
    L.hvpr_mem_attn.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_size_t, ctypes.POINTER(HvprZeroFill), c_void_p]
    L.hvpr_mem_train_forward.restype = c_int
    L.hvpr_mem_train_forward.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
    L.hvpr_mse_loss.restype = c_int
    L.hvpr_mse_loss.argtypes = [c_void_p, c_void_p, c_int64, ctypes.c_double, c_void_p, c_void_p, c_void_p]
    L.hvpr_bev_fill.restype = c_int
    L.hvpr_bev_fill.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_void_p, ctypes.POINTER(HvprLaunchCfg), c_void_p]
    L.hvpr_build_cell_map.restype = c_int
    L.hvpr_build_cell_map.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]
    L.hvpr_conv_packed_bytes.restype = c_size_t
    L.hvpr_conv_packed_bytes.argtypes = [c_int, c_int, c_int]
    L.hvpr_conv_pack_weights.restype = c_int
    L.hvpr_conv_pack_weights.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    L.hvpr_conv2d.restype = c_int
    L.hvpr_conv2d.argtypes = [ctypes.POINTER(HvprConvArgs), c_void_p]
    L.hvpr_bev_fill_nhwc_bf16.restype = c_int
    L.hvpr_bev_fill_nhwc_bf16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_int, c_void_p, c_int, c_void_p]
    L.hvpr_post_process_workspace_bytes.restype = c_size_t
    L.hvpr_post_process_workspace_bytes.argtypes = [c_int, c_int64]
    L.hvpr_post_process.restype = c_int
    L.hvpr_post_process.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.hvpr_boxes_iou3d.restype = c_int
    L.hvpr_boxes_iou3d.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]
    L.hvpr_head_decode.restype = c_int
    L.hvpr_head_decode.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_float, c_float, c_void_p, c_void_p, c_void_p]
    L.hvpr_nchw_to_nhwc_bf16.restype = c_int
    L.hvpr_nchw_to_nhwc_bf16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
    L.hvpr_attention_gate.restype = c_int
    L.hvpr_attention_gate.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float), c_float,
                                      c_void_p, c_void_p, c_void_p]
    _lib = L
    return L


def check(status: int, what: str = ""):
    if status != 0:
        L = lib()
        msg = L.hvpr_strerror(status).decode()
        if status == -4:
            msg += ": " + L.hvpr_last_cuda_error().decode()
        raise HvprError("%s failed: %s (%d)" % (what or "hvpr call", msg, status))


def init_device():
    """hvpr_init() once per (process, device): opt-in shared-memory attributes.  Must precede graph capture."""
    import torch
    if not torch.cuda.is_available():
        raise HvprError("hvpr_b200 needs a CUDA device (sm_100a); there is no CPU path")
    d = torch.cuda.current_device()
    if d not in _inited_devices:
        check(lib().hvpr_init(), "hvpr_init")
        _inited_devices.add(d)


def make_geom(pc_range, voxel_size, grid_size) -> HvprGeom:
    g = HvprGeom()
    for i in range(3):
        g.lo[i] = float(pc_range[i])
        g.vs[i] = float(voxel_size[i])
        g.grid[i] = int(grid_size[i])
    return g


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
