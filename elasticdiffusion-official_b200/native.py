"""ctypes binding of `libelastic_b200.so` (C ABI declared in include/elastic_b200.h) + the in-tree build recipe.

The reference has no native layer to mirror; this module is the only place where Python touches
the CUDA library.  There is NO fallback: if the shared object is missing or a call fails, a `NativeError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libelastic_b200.so")
SOURCES = ["abi.cu", "gather.cu", "epilogue.cu", "tiles.cu", "unet_ops.cu"]
ED_MAX_RENOISE = 1000
ED_F32, ED_F16, ED_BF16 = 0, 1, 2
FLAG_RENOISE, FLAG_RRG, FLAG_FP16_SEM = 1, 2, 4
EPILOGUE_AUTO, EPILOGUE_DIRECT, EPILOGUE_STAGED, EPILOGUE_HALF = 0, 1, 2, 3
PLAN_HALF_FAST = 1
ABI_VERSION = 4


class NativeError(RuntimeError):
    pass


class Plan(C.Structure):
    """ed_plan_t"""
    _fields_ = [(n, C.c_int32) for n in
                ("B", "C", "H", "W", "dH", "dW", "lh", "lw", "g_tp", "g_lp", "nv", "nvr", "nvc", "vh", "vw",
                 "v_tp", "v_lp", "flags")] + \
               [(n, C.c_void_p) for n in
                ("row_src", "col_src", "mrow_lo", "mrow_n", "mcol_lo", "mcol_n", "up_row", "up_col", "down_row",
                 "down_col", "views", "vrow_first", "vrow_cnt", "vcol_first", "vcol_cnt", "pix_ref", "cell_cand",
                 "cell_down", "vrow_off", "vcol_off")]


class StepParams(C.Structure):
    """ed_step_params_t"""
    _fields_ = [("guidance", C.c_float), ("sqrt_beta_t", C.c_float), ("sqrt_alpha_t", C.c_float),
                ("sqrt_alpha_prev", C.c_float), ("sqrt_dir", C.c_float), ("rrg_weight", C.c_float),
                ("rrg_norm", C.c_float), ("flags", C.c_int32), ("n_renoise", C.c_int32), ("R1", C.c_int32),
                ("reserved", C.c_int32 * 2), ("renoise_a", C.c_float * ED_MAX_RENOISE),
                ("renoise_b", C.c_float * ED_MAX_RENOISE)]


class Tiles(C.Structure):
    """ed_tiles_t"""
    _fields_ = [(n, C.c_int32) for n in ("ntiles", "ntc", "core", "pad", "scale", "B", "CH", "H", "W", "reserved")] + \
               [(n, C.c_void_p) for n in ("tiles", "trow_first", "trow_cnt", "tcol_first", "tcol_cnt")]


EXPORTS = {
    "ed_abi_version": (C.c_int, []),
    "ed_strerror": (C.c_char_p, [C.c_int]),
    "ed_last_cuda_error": (C.c_int, []),
    "ed_device_check": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "ed_upload_step_params": (C.c_int, [C.c_void_p, C.POINTER(StepParams), C.c_void_p]),
    "ed_gather_views": (C.c_int, [C.POINTER(Plan), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ed_random_pick_gather": (C.c_int, [C.POINTER(Plan), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                        C.c_void_p, C.c_int, C.c_void_p]),
    "ed_pad_views": (C.c_int, [C.POINTER(Plan), C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ed_owner_map": (C.c_int, [C.POINTER(Plan), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ed_wave_epilogue": (C.c_int, [C.POINTER(Plan), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ed_wave_epilogue_peer": (C.c_int, [C.POINTER(Plan), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ed_set_epilogue_mode": (C.c_int, [C.c_int]),
    "ed_epilogue_launch_counts": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ed_epilogue_launch_counts3": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ed_renoise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ed_gather_cond": (C.c_int, [C.POINTER(Plan), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_void_p]),
    "ed_tile_gather": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p]),
    "ed_tile_blend": (C.c_int, [C.POINTER(Tiles), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ed_geglu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "ed_bias_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ed_bias_add_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ed_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "ed_groupnorm_split": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ed_groupnorm_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "ed_tile_blend_peer": (C.c_int, [C.POINTER(Tiles), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libelastic_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", s) for s in SOURCES]
    deps = srcs + [os.path.join(_HERE, "csrc", "common.cuh"), os.path.join(_HERE, "csrc", "epilogue_staged.cuh"),
                   os.path.join(_HERE, "csrc", "epilogue_half.cuh"),
                   os.path.join(_ROOT, "include", "elastic_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH] + srcs
    cmd[1:1] = os.environ.get("ED_NVCC_FLAGS", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise NativeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB_PATH


def lib():
    """The loaded shared library (loud failure when absent - there is no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA extension is mandatory, there is no fallback path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)           # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if l.ed_abi_version() != ABI_VERSION:
            raise NativeError("libelastic_b200 ABI version mismatch")
        _lib = l
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        l = lib()
        msg = l.ed_strerror(status).decode()
        raise NativeError(f"libelastic_b200 {what}: {msg} (status {status}, cuda error {l.ed_last_cuda_error()})")


def epilogue_launch_counts():
    """(direct, staged, half): how many wave-epilogue launches of this process took each kernel."""
    d, s, h = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    check(lib().ed_epilogue_launch_counts3(C.byref(d), C.byref(s), C.byref(h)), "ed_epilogue_launch_counts3")
    return d.value, s.value, h.value


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dtype_code(dt):
    import torch
    return {torch.float32: ED_F32, torch.float16: ED_F16, torch.bfloat16: ED_BF16}[dt]


def stream_handle():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def plan_from_geometry(geo, device, flags=None):
    """(ed_plan_t, keep): uploads the int32 tables of a `geometry.WaveGeometry` to `device` and fills the struct; `keep`
    holds the table tensors and must outlive every launch that uses the plan.  `flags`: override of geo.flags (tests)."""
    import torch
    keep = {k: torch.tensor(v if len(v) else [0], dtype=torch.int32, device=device) for k, v in geo.tables.items()}
    lp, rp, tp, bp = geo.g_pad
    vlp, vrp, vtp, vbp = geo.v_pad
    plan = Plan(B=geo.B, C=geo.C, H=geo.H, W=geo.W, dH=geo.native, dW=geo.native, lh=geo.lh, lw=geo.lw, g_tp=tp, g_lp=lp,
                nv=geo.nv, nvr=geo.nvr, nvc=geo.nvc, vh=geo.vh, vw=geo.vw, v_tp=vtp, v_lp=vlp,
                flags=geo.flags if flags is None else flags, **{k: v.data_ptr() for k, v in keep.items()})
    return plan, keep


def strips_array(strips):
    """4 device pointers (left, right, top, bottom); None -> NULL."""
    arr = (C.c_void_p * 4)()
    for i, s in enumerate(strips):
        arr[i] = s.data_ptr() if s is not None and s.numel() > 0 else None
    return arr
