"""CPU: the staged wave-epilogue kernel SOURCE (csrc/epilogue_staged.cuh) compiled for the host (tests/emu/emu_shim.h:
std::threads for CUDA threads, a synchronous box copy for TMA) against the contract emulation oracle/wave_spec.py,
bit-exact - index math, box geometry, shared-memory layout, fall-back-to-global decisions and launch geometry
(`staged_config`, shared with the CUDA launcher) are checked here without a GPU.  The GPU suite repeats the same cases on
the real kernel (tests/test_gpu_kernels.py)."""
import ctypes
import os
import subprocess

import pytest
import torch

from conftest import PKG, ROOT
from oracle import wave_spec as ws

geometry, native = PKG.geometry, PKG.native
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libed_emu.so")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, "emu_epilogue.cpp"), os.path.join(EMU_DIR, "emu_shim.h"),
            os.path.join(ROOT, "elasticdiffusion-official_b200", "csrc", "epilogue_staged.cuh"),
            os.path.join(ROOT, "elasticdiffusion-official_b200", "csrc", "epilogue_half.cuh"),
            os.path.join(ROOT, "include", "elastic_b200.h")]
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(EMU_LIB) < os.path.getmtime(s) for s in srcs):
        cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-DED_HOST_EMU", "-I", EMU_DIR, "-I", os.path.join(ROOT, "include"),
               "-I", os.path.join(ROOT, "elasticdiffusion-official_b200", "csrc"), "-shared", "-fPIC", "-pthread", srcs[0],
               "-o", EMU_LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(EMU_LIB)
    lib.emu_wave_epilogue.restype = ctypes.c_int
    lib.emu_wave_epilogue.argtypes = [ctypes.POINTER(native.Plan), ctypes.POINTER(native.StepParams), ctypes.c_int] + \
        [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.emu_wave_epilogue_half.restype = ctypes.c_int
    lib.emu_wave_epilogue_half.argtypes = [ctypes.POINTER(native.Plan), ctypes.POINTER(native.StepParams), ctypes.c_int] + \
        [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4 + [ctypes.POINTER(ctypes.c_int)]
    return lib


def host_plan(geo, flags=None):
    return native.plan_from_geometry(geo, "cpu", flags)


# (B, H, W, native, ds, window) - W % 4 == 0 (the staged kernel's domain; other widths take the direct kernel)
GEOS = [(1, 128, 256, 128, (64, 128), 64),      # cfg3
        (1, 64, 128, 64, (32, 64), 32),         # cfg2
        (1, 64, 64, 64, (64, 64), 32),          # cfg1: identity ratio, 1 view
        (2, 192, 192, 128, (128, 128), 64),     # 2/3 ratio, B = 2
        (1, 135, 240, 128, (72, 128), 64),      # 1080x1920: odd height, overlapping views (multi-cover walk)
        (1, 80, 112, 64, (45, 64), 32),         # ragged
        (1, 128, 256, 64, (32, 64), 32),        # downsample factor 4
        (1, 192, 192, 64, (64, 64), 32),        # downsample factor 3
        (1, 128, 256, 128, (64, 128), 32),      # patch_size 32: overlapping last windows
        (2, 96, 256, 128, (48, 128), 64),       # window collapse: padded views
        (1, 72, 100, 64, (36, 50), 32)]         # unaligned low-res offset inside the canvas (g_lp = 7): boxes at odd coordinates


def run_case(emu, cfg, mode, dtype, R1, sms, shrink=0, scalar_views=0, origin=3, cpt=2):
    B, H, W, nat, ds, window = cfg
    geo = geometry.build_geometry(B, 4, H, W, nat, ds, window, window, nat - window)
    plan, keep = host_plan(geo)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, H, W, generator=g)
    idx = torch.randint(0, 4, (R1, geo.lh * geo.lw), generator=g, dtype=torch.uint8)
    idx[0] = 0
    n = 2 * B * R1 + geo.nv * B
    out = torch.randn(n, 4, nat, nat, generator=g).to(dtype)
    out[2 * B * R1:][torch.rand(geo.nv * B, 4, nat, nat, generator=g) < 0.05] = 0    # the "!= 0" first-writer rule
    n_re = 6
    noise = torch.randn(n_re, *x.shape, generator=g)
    flags = {"plain": 0, "renoise": 1, "rrg": 2}[mode] | (4 if dtype == torch.float16 else 0)
    prm = dict(guidance=7.5, sqrt_beta_t=0.9637, sqrt_alpha_t=0.2669, sqrt_alpha_prev=0.3316, sqrt_dir=0.9434,
               rrg_weight=731.25, rrg_norm=2.0 / (4 * H * W), flags=flags, n_renoise=n_re if mode == "renoise" else 0, R1=R1)
    sp = native.StepParams(**prm)
    for k in range(n_re):
        sp.renoise_a[k], sp.renoise_b[k] = 0.99 - 0.001 * k, 0.1 + 0.002 * k
    for k in ("guidance", "sqrt_beta_t", "sqrt_alpha_t", "sqrt_alpha_prev", "sqrt_dir", "rrg_weight", "rrg_norm"):
        prm[k] = float(getattr(sp, k))
    prm["renoise_a"] = [float(sp.renoise_a[k]) for k in range(n_re)]
    prm["renoise_b"] = [float(sp.renoise_b[k]) for k in range(n_re)]
    owner = ws.owner_map(geo, R1, idx, "cpu").to(torch.uint8).contiguous().view(-1)
    y, x0 = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
    info = (ctypes.c_int * 16)()
    info[5], info[6], info[7], info[9] = origin, shrink, scalar_views, cpt
    # like the pipeline, only re-noise launches pass a noise buffer (it selects the kernel instantiation)
    rc = emu.emu_wave_epilogue(ctypes.byref(plan), ctypes.byref(sp), R1, x.data_ptr(), out.data_ptr(), native.dtype_code(dtype),
                               idx.data_ptr(), owner.data_ptr(), noise.data_ptr() if mode == "renoise" else None, y.data_ptr(),
                               x0.data_ptr(), sms, info)
    assert rc == 0, rc
    want, want_x0 = ws.spec_epilogue(geo, prm, x, out, idx, noise)
    assert torch.equal(x0, want_x0), f"x0 max diff {(x0 - want_x0).abs().max().item():.3e} geometry {list(info)}"
    assert torch.equal(y, want), f"latent max diff {(y - want).abs().max().item():.3e} geometry {list(info)}"
    return list(info)


@pytest.mark.parametrize("cfg", GEOS)
@pytest.mark.parametrize("mode,dtype", [("rrg", torch.float32), ("renoise", torch.bfloat16), ("rrg", torch.float16),
                                        ("plain", torch.float32)])
def test_staged_epilogue_source_matches_spec_on_host(emu, cfg, mode, dtype):
    R1 = 1 if (mode == "rrg" and cfg[0] == 2) else 4
    run_case(emu, cfg, mode, dtype, R1, sms=148)


@pytest.mark.parametrize("origin", [0, 1, 2])
def test_staged_epilogue_box_origin_policies(emu, origin):
    """0: box at the first needed cell (unaligned, may hang over the canvas edge: zero fill), 1: 16-byte aligned start,
    2: shifted back inside the canvas; the default 3 = both runs everywhere else."""
    for cfg in (GEOS[3], GEOS[5], GEOS[7], GEOS[10]):
        run_case(emu, cfg, "rrg", torch.bfloat16, 1 if cfg[0] == 2 else 3, sms=148, origin=origin)


def test_staged_epilogue_large_cta_geometry_and_many_iterations(emu):
    # sms = 1: "fills the GPU twice" holds at once -> the 256-thread CTA (8 x 128 tile) is chosen
    info = run_case(emu, GEOS[0], "rrg", torch.bfloat16, 8, sms=1, cpt=4)
    assert info[:4] == [32, 8, 64, 4] and info[4] == 16 * 2048 + 64 * 4 * 4 * 4
    info = run_case(emu, GEOS[0], "rrg", torch.bfloat16, 8, sms=1, cpt=2)
    assert info[:4] == [32, 8, 64, 4] and info[4] == 16 * 1024 + 64 * 4 * 2 * 4
    # R1 = 21 (the signature default resampling_steps = 20) in fp32: boxes only fit with a smaller tile
    info = run_case(emu, GEOS[1], "rrg", torch.float32, 21, sms=1)
    assert info[4] <= 200 * 1024


def test_staged_epilogue_path_coverage(emu):
    """counters of the emulation (info[8:]): staged / unstaged threads, RRG down-cell inside / outside the boxes, vector /
    scalar view loads, multi-cover walks - every branch of the kernel runs somewhere in this suite."""
    tot = [0] * 7
    for cfg, kw in [(GEOS[0], {}), (GEOS[4], {}), (GEOS[5], {}), (GEOS[7], {}), (GEOS[0], dict(shrink=1)), (GEOS[3], dict(shrink=1)),
                    (GEOS[1], dict(scalar_views=1))]:
        info = run_case(emu, cfg, "rrg", torch.bfloat16, 1 if cfg[0] == 2 else 3, sms=148, **kw)
        tot = [a + b for a, b in zip(tot, info[8:15])]
        print(cfg, kw, info[:6], info[8:15])
    assert all(t > 0 for t in tot), tot


def test_reciprocal_division_equals_ieee_division(tmp_path):
    """struct DivBy (csrc/epilogue_staged.cuh): q = a*y, two FMA corrections == a / b for every divisor the DDIM schedule
    produces and random ones (tests/emu/div_check.c; 6e7 quotients here, 1.2e9 with the default argument)."""
    exe = str(tmp_path / "div_check")
    r = subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(EMU_DIR, "div_check.c"), "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "20000"], capture_output=True, text=True)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout


# ---- half kernels (csrc/epilogue_half.cuh): plans that carry ED_PLAN_HALF_FAST -------------------------------------------
HALF_GEOS = [(1, 128, 256, 128, (64, 128), 64),     # cfg3
             (1, 64, 128, 64, (32, 64), 32),        # cfg2
             (1, 256, 256, 128, (128, 128), 64),    # cfg4
             (2, 96, 256, 128, (48, 128), 64),      # window collapse: padded views (v_tp = 16), B = 2
             (1, 128, 256, 128, (64, 128), 32),     # patch_size 32: 4 x 8 windows
             (3, 32, 48, 64, (16, 24), 32)]         # narrow: W / 8 = 6 column groups (guard lanes), low-res latent padded (g_lp = 20)


def run_half_case(emu, cfg, mode, dtype, R1, world=None, grid_z=0, x0_out=True, poison=False):
    B, H, W, nat, ds, window = cfg
    geo = geometry.build_geometry(B, 4, H, W, nat, ds, window, window, nat - window)
    assert geo.flags & native.PLAN_HALF_FAST, cfg
    plan, keep = host_plan(geo)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 4, H, W, generator=g)
    if poison:      # quotients outside the reciprocal division's range: the tile must be redone with IEEE division
        flat = x.view(-1)
        where = torch.randperm(flat.numel(), generator=g)[:64]
        flat[where[:16]] = float("inf")
        flat[where[16:32]] = -float("inf")
        flat[where[32:48]] = 1e38                     # 1e38 / 0.2669 overflows
        flat[where[48:]] = float("nan")
    idx = torch.randint(0, 4, (R1, geo.lh * geo.lw), generator=g, dtype=torch.uint8)
    if mode != "rrg-anypick":
        idx[0] = 0                                   # the pipeline's k = 0 pick; "rrg-anypick": R1 = 1 with arbitrary picks
    n = 2 * B * R1 + geo.nv * B
    out = torch.randn(n, 4, nat, nat, generator=g).to(dtype)
    out[2 * B * R1:][torch.rand(geo.nv * B, 4, nat, nat, generator=g) < 0.05] = 0
    flags = (2 if mode.startswith("rrg") else 0) | (4 if dtype == torch.float16 else 0)
    prm = dict(guidance=7.5, sqrt_beta_t=0.9637, sqrt_alpha_t=0.2669, sqrt_alpha_prev=0.3316, sqrt_dir=0.9434,
               rrg_weight=731.25, rrg_norm=2.0 / (4 * H * W), flags=flags, n_renoise=0, R1=R1)
    sp = native.StepParams(**prm)
    for k in ("guidance", "sqrt_beta_t", "sqrt_alpha_t", "sqrt_alpha_prev", "sqrt_dir", "rrg_weight", "rrg_norm"):
        prm[k] = float(getattr(sp, k))
    owner = ws.owner_map(geo, R1, idx, "cpu").to(torch.uint8).contiguous().view(-1)
    y, x0 = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
    info = (ctypes.c_int * 8)()
    info[0] = grid_z
    peers, bufs, per = None, [], 0
    if world:
        per = (n + world - 1) // world
        for r in range(world):
            buf = torch.full((per,) + tuple(out.shape[1:]), float("nan"), dtype=dtype)
            lo, hi = min(r * per, n), min((r + 1) * per, n)
            buf[:hi - lo] = out[lo:hi]
            bufs.append(buf)
        peers = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
    rc = emu.emu_wave_epilogue_half(ctypes.byref(plan), ctypes.byref(sp), R1, x.data_ptr(), None if world else out.data_ptr(),
                                    peers, world or 0, per, native.dtype_code(dtype), idx.data_ptr(), owner.data_ptr(),
                                    y.data_ptr(), x0.data_ptr() if x0_out else None, info)
    assert rc == 0, rc
    want, want_x0 = ws.spec_epilogue(geo, prm, x, out, idx, None)
    bits = lambda t: t.view(torch.int32) if not poison else torch.where(torch.isnan(t), torch.zeros_like(t), t).view(torch.int32)
    if x0_out:
        assert torch.equal(bits(x0), bits(want_x0)), f"x0 max diff {(x0 - want_x0).abs().max().item():.3e} {list(info)}"
        assert torch.equal(torch.isnan(x0), torch.isnan(want_x0))
    assert torch.equal(bits(y), bits(want)), f"latent max diff {(y - want).abs().max().item():.3e} {list(info)}"
    assert torch.equal(torch.isnan(y), torch.isnan(want))
    return list(info)


def test_half_epilogue_redoes_tiles_with_ieee_division_when_a_quotient_leaves_the_fast_range(emu):
    for mode, dtype, R1 in (("plain", torch.bfloat16, 1), ("rrg", torch.float32, 1), ("rrg", torch.bfloat16, 3)):
        run_half_case(emu, HALF_GEOS[1], mode, dtype, R1, poison=True)


@pytest.mark.parametrize("cfg", HALF_GEOS)
@pytest.mark.parametrize("mode,dtype,R1", [("plain", torch.bfloat16, 1), ("rrg", torch.bfloat16, 1), ("rrg", torch.float16, 1),
                                           ("rrg-anypick", torch.float32, 1), ("plain", torch.float32, 3),
                                           ("rrg", torch.bfloat16, 4), ("rrg", torch.float16, 2)])
def test_half_epilogue_source_matches_spec_on_host(emu, cfg, mode, dtype, R1):
    run_half_case(emu, cfg, mode, dtype, R1)


def test_half_epilogue_peer_mapping_grid_stride_and_no_x0(emu):
    run_half_case(emu, HALF_GEOS[0], "rrg", torch.bfloat16, 1, world=8)          # wave 2 of cfg3 over 8 ranks (2 idle)
    run_half_case(emu, HALF_GEOS[1], "rrg", torch.float16, 3, world=3)
    run_half_case(emu, HALF_GEOS[3], "plain", torch.float32, 2, world=2)
    run_half_case(emu, HALF_GEOS[5], "rrg", torch.bfloat16, 2, grid_z=5)          # 12 (b, c) planes over a grid z of 5
    run_half_case(emu, HALF_GEOS[1], "plain", torch.bfloat16, 1, x0_out=False)


def test_half_flag_is_only_set_on_exact_half_tiling_geometries():
    for cfg in GEOS:
        B, H, W, nat, ds, window = cfg
        geo = geometry.build_geometry(B, 4, H, W, nat, ds, window, window, nat - window)
        exact = (2 * ds[0] == H and 2 * ds[1] == W and W % 8 == 0 and H % 2 == 0)
        assert bool(geo.flags & native.PLAN_HALF_FAST) == exact, cfg
