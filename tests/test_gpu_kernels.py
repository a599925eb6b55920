"""GPU: every C-ABI kernel against the oracle's contract emulation (oracle/wave_spec.py), bit-exact, on seeded
inputs - BASELINE shapes, ragged / general-ratio shapes, every output dtype, and L2-exceeding sizes."""
import ctypes

import pytest
import torch

from conftest import PKG
from oracle import wave_spec as ws

pytestmark = pytest.mark.gpu
geometry, native = PKG.geometry, PKG.native
DEV = "cuda"


def upload(geo, flags=None):
    return native.plan_from_geometry(geo, DEV, flags)


def make_strips(geo, inner_h, inner_w, seed):
    g = torch.Generator().manual_seed(seed)
    nat = geo.native
    l, r = geometry.pad_split(nat, inner_w)
    t, b = geometry.pad_split(nat, inner_h)
    mk = lambda h, w: torch.randn(1, geo.C, h, w, generator=g).to(DEV) if h and w else None
    return [mk(inner_h, l), mk(inner_h, r), mk(t, inner_w + l + r), mk(b, inner_w + l + r)]


def rand_idx(R1, n, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, 4, (R1, n), generator=g, dtype=torch.uint8)
    idx[0] = 0
    return idx.to(DEV)


# (B, H, W, native, ds, window)
GEOS = [(1, 128, 256, 128, (64, 128), 64),      # cfg3  SDXL 1024x2048
        (1, 64, 128, 64, (32, 64), 32),         # cfg2  SD2.1 512x1024
        (1, 64, 64, 64, (64, 64), 32),          # cfg1  SD1.5 512x512 (identity ratio, 1 view)
        (1, 256, 256, 128, (128, 128), 64),     # cfg4  SDXL 2048x2048
        (2, 192, 192, 128, (128, 128), 64),     # 2/3 ratio, B=2
        (1, 135, 240, 128, (72, 128), 64),      # 1080x1920: odd width -> non-TMA / non-vector paths, overlapping views
        (1, 80, 112, 64, (45, 64), 32),         # ragged SD
        (1, 96, 128, 64, (48, 64), 32),
        (1, 128, 256, 128, (64, 128), 32),      # patch_size=32: overlapping last windows
        (2, 96, 256, 128, (48, 128), 64),       # window collapse: views 96 rows < native 128 -> ed_pad_views, v_tp offsets
        (1, 128, 256, 64, (32, 64), 32),        # downsample factor 4 (SD 1024x2048; the reference's "factor > 2" TODO path)
        (1, 192, 192, 64, (64, 64), 32),        # downsample factor 3
        (1, 72, 100, 64, (36, 50), 32)]         # W % 4 == 0 but unaligned low-res offsets (g_lp = 7): TMA boxes at odd coordinates


def build(cfg):
    B, H, W, nat, ds, window = cfg
    return geometry.build_geometry(B, 4, H, W, nat, ds, window, window, nat - window)


@pytest.mark.parametrize("cfg", GEOS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_gather_kernels_match_spec(cfg, dtype):
    L = native.lib()
    geo = build(cfg)
    plan, keep = upload(geo)
    R1 = 3
    torch.manual_seed(0)
    x = torch.randn(geo.B, geo.C, geo.H, geo.W, device=DEV)
    idx = rand_idx(R1, geo.lh * geo.lw, 1)
    sg = make_strips(geo, geo.lh, geo.lw, 2)
    sv = make_strips(geo, geo.vh, geo.vw, 3)
    n = 2 * geo.B * R1 + geo.nv * geo.B
    canvas = torch.full((n, geo.C, geo.native, geo.native), float("nan"), device=DEV, dtype=dtype)
    st = native.stream_handle()
    native.check(L.ed_random_pick_gather(ctypes.byref(plan), R1, native.ptr(x), native.ptr(idx), native.strips_array(sg),
                                         native.ptr(canvas), native.dtype_code(dtype), st))
    native.check(L.ed_gather_views(ctypes.byref(plan), native.ptr(x), native.ptr(canvas), native.dtype_code(dtype),
                                   2 * geo.B * R1, st))
    if any(s is not None for s in sv):
        native.check(L.ed_pad_views(ctypes.byref(plan), native.strips_array(sv), native.ptr(canvas),
                                    native.dtype_code(dtype), 2 * geo.B * R1, st))
    torch.cuda.synchronize()
    want = torch.cat([ws.spec_pick_gather(geo, R1, x, idx, sg), ws.spec_gather_views(geo, x, sv)]).to(dtype)
    assert torch.equal(canvas, want)


@pytest.fixture
def epilogue_kernel(request):
    """direct = scattered-load kernel, staged = TMA tile-staged kernel, half = exact-1/2-ratio kernels (forced: an
    unsupported shape is an error, not a silent fall-back to another kernel)."""
    L = native.lib()
    native.check(L.ed_set_epilogue_mode({"direct": native.EPILOGUE_DIRECT, "staged": native.EPILOGUE_STAGED,
                                         "half": native.EPILOGUE_HALF}[request.param]))
    yield request.param
    native.check(L.ed_set_epilogue_mode(native.EPILOGUE_AUTO))


@pytest.mark.parametrize("cfg", GEOS)
@pytest.mark.parametrize("mode", ["plain", "renoise", "rrg"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("epilogue_kernel", ["direct", "staged"], indirect=True)
def test_wave_epilogue_matches_spec(cfg, mode, dtype, epilogue_kernel, R1=None):
    _epilogue_case(cfg, mode, dtype, R1)


def _epilogue_case(cfg, mode, dtype, R1=None, peer_world=None, plan_flags=None, poison=False):
    """One fused-epilogue launch against the contract emulation.  `peer_world`: go through ed_wave_epilogue_peer with the
    wave's samples split over `peer_world` per-"rank" buffers (sample s lives in buffer s // per at index s % per) - the
    multi-GPU entry point exercised on ONE GPU: the pointer table simply points at separate local allocations."""
    L = native.lib()
    geo = build(cfg)
    plan, keep = upload(geo, plan_flags)
    R1 = R1 or (1 if (mode == "rrg" and cfg[0] == 2) else 4)
    torch.manual_seed(1)
    x = torch.randn(geo.B, geo.C, geo.H, geo.W, device=DEV)
    if poison:      # quotients outside the reciprocal division's range (inf, nan, overflow): IEEE-division results expected
        flat = x.view(-1)
        where = torch.randperm(flat.numel(), device=DEV)[:64]
        flat[where[:16]] = float("inf")
        flat[where[16:32]] = -float("inf")
        flat[where[32:48]] = 1e38
        flat[where[48:]] = float("nan")
    idx = rand_idx(R1, geo.lh * geo.lw, 5)
    n = 2 * geo.B * R1 + geo.nv * geo.B
    out = torch.randn(n, geo.C, geo.native, geo.native, device=DEV).to(dtype)
    # exercise the "!= 0" first-writer rule: zero some view outputs exactly
    out[2 * geo.B * R1:][torch.rand(geo.nv * geo.B, geo.C, geo.native, geo.native, device=DEV) < 0.05] = 0
    n_re = 20
    noise = torch.randn(n_re, *x.shape, device=DEV)
    flags = {"plain": 0, "renoise": 1, "rrg": 2}[mode] | (4 if dtype == torch.float16 else 0)
    prm = dict(guidance=7.5, sqrt_beta_t=0.9637, sqrt_alpha_t=0.2669, sqrt_alpha_prev=0.3316, sqrt_dir=0.9434,
               rrg_weight=731.25, rrg_norm=2.0 / (geo.C * geo.H * geo.W), flags=flags, n_renoise=n_re if mode == "renoise" else 0,
               R1=R1, renoise_a=[0.99 - 0.001 * k for k in range(n_re)], renoise_b=[0.1 + 0.002 * k for k in range(n_re)])
    sp = native.StepParams(**{k: v for k, v in prm.items() if not k.startswith("renoise_")})
    for k in range(n_re):
        sp.renoise_a[k], sp.renoise_b[k] = prm["renoise_a"][k], prm["renoise_b"][k]
    # the spec must see the float32-rounded scalars the struct holds
    for k in ("guidance", "sqrt_beta_t", "sqrt_alpha_t", "sqrt_alpha_prev", "sqrt_dir", "rrg_weight", "rrg_norm"):
        prm[k] = float(getattr(sp, k))
    prm["renoise_a"] = [float(sp.renoise_a[k]) for k in range(n_re)]
    prm["renoise_b"] = [float(sp.renoise_b[k]) for k in range(n_re)]
    d_prm = torch.empty(ctypes.sizeof(native.StepParams), dtype=torch.uint8, device=DEV)
    st = native.stream_handle()
    native.check(L.ed_upload_step_params(native.ptr(d_prm), ctypes.byref(sp), st))
    y = torch.empty_like(x)
    x0 = torch.empty_like(x)
    owner = torch.full((geo.H * geo.W,), 255, dtype=torch.uint8, device=DEV)
    native.check(L.ed_owner_map(ctypes.byref(plan), R1, native.ptr(idx), native.ptr(owner), st))
    assert torch.equal(owner.view(geo.H, geo.W).long(), ws.owner_map(geo, R1, idx, DEV))
    # like the pipeline, only re-noise launches pass a noise buffer: NULL selects the staged kernel's lighter instantiation
    # (no noise stream, channel pairs spread over the grid), a buffer its 4-channels-per-thread one
    if peer_world is None:
        native.check(L.ed_wave_epilogue(ctypes.byref(plan), native.ptr(d_prm), R1, native.ptr(x), native.ptr(out),
                                        native.dtype_code(dtype), native.ptr(idx), native.ptr(owner),
                                        native.ptr(noise) if mode == "renoise" else None, native.ptr(y), native.ptr(x0), st))
    else:
        per = (n + peer_world - 1) // peer_world          # pipeline._unet's partition: ragged last rank, idle ranks
        bufs = []
        for r in range(peer_world):
            buf = torch.full((per,) + tuple(out.shape[1:]), float("nan"), device=DEV, dtype=dtype)   # never-written slots
            lo, hi = min(r * per, n), min((r + 1) * per, n)
            buf[:hi - lo] = out[lo:hi]
            bufs.append(buf)
        ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
        native.check(L.ed_wave_epilogue_peer(ctypes.byref(plan), native.ptr(d_prm), R1, native.ptr(x), native.ptr(ptrs),
                                             peer_world, per, native.dtype_code(dtype), native.ptr(idx), native.ptr(owner),
                                             native.ptr(noise) if mode == "renoise" else None, native.ptr(y),
                                             native.ptr(x0), st))
    torch.cuda.synchronize()
    want, want_x0 = ws.spec_epilogue(geo, prm, x, out, idx, noise)
    if poison:      # nan payloads are not compared
        for got, ref in ((x0, want_x0), (y, want)):
            assert torch.equal(torch.isnan(got), torch.isnan(ref))
            assert torch.equal(torch.nan_to_num(got, nan=0.0, posinf=3e38, neginf=-3e38), torch.nan_to_num(ref, nan=0.0, posinf=3e38, neginf=-3e38))
            assert torch.equal(torch.isinf(got), torch.isinf(ref))
        return y
    assert torch.equal(x0, want_x0), f"x0 max diff {(x0 - want_x0).abs().max().item():.3e}"
    assert torch.equal(y, want), f"latent max diff {(y - want).abs().max().item():.3e}"
    return y


# plans that carry ED_PLAN_HALF_FAST (exact 1/2 ratio, tiling views)
HALF_GEOS = [GEOS[0], GEOS[1], GEOS[3], GEOS[8], GEOS[9],
             (3, 32, 48, 64, (16, 24), 32),          # narrow: 6 column groups (guard lanes), low-res latent at g_lp = 20
             (24, 128, 256, 128, (64, 128), 64)]     # a batch that fills the GPU several times over


@pytest.mark.parametrize("cfg", HALF_GEOS)
@pytest.mark.parametrize("mode", ["plain", "rrg"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("R1", [1, 4, 8])
@pytest.mark.parametrize("epilogue_kernel", ["half"], indirect=True)
def test_half_epilogue_matches_spec(cfg, mode, dtype, R1, epilogue_kernel):
    """the half kernels (csrc/epilogue_half.cuh), forced: R1 == 1 streaming kernel and R1 > 1 per-thread-slot kernel."""
    assert build(cfg).flags & native.PLAN_HALF_FAST
    h0 = native.epilogue_launch_counts()[2]
    _epilogue_case(cfg, mode, dtype, R1=R1)
    assert native.epilogue_launch_counts()[2] == h0 + 1


@pytest.mark.parametrize("epilogue_kernel", ["half"], indirect=True)
def test_half_mode_refuses_what_it_cannot_do(epilogue_kernel):
    with pytest.raises(native.NativeError):
        _epilogue_case(GEOS[5], "plain", torch.float32, R1=2)         # general ratio: no ED_PLAN_HALF_FAST
    with pytest.raises(native.NativeError):
        _epilogue_case(GEOS[0], "renoise", torch.bfloat16, R1=4)      # noise stream: the staged kernel's job


@pytest.mark.parametrize("cfg,mode,R1,world", [(GEOS[0], "rrg", 1, 8), (GEOS[0], "plain", 1, 4), (GEOS[0], "rrg", 8, 8),
                                               (GEOS[3], "rrg", 1, 8), (GEOS[1], "plain", 5, 3), (GEOS[9], "rrg", 1, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("epilogue_kernel", ["half"], indirect=True)
def test_half_epilogue_peer_matches_spec_on_one_gpu(cfg, mode, R1, world, dtype, epilogue_kernel):
    _epilogue_case(cfg, mode, dtype, R1=R1, peer_world=world)


@pytest.mark.parametrize("epilogue_kernel", ["half", "staged", "direct"], indirect=True)
def test_non_finite_quotients_follow_ieee_division(epilogue_kernel):
    """inf / nan / overflowing quotients: the reciprocal-based division of the staged and half kernels must fall back to
    IEEE division (per quotient in the staged kernel, per tile in the half kernels)."""
    for mode, R1 in (("plain", 1), ("rrg", 1), ("rrg", 3)):
        _epilogue_case(GEOS[1], mode, torch.bfloat16, R1=R1, poison=True)


def test_plan_flag_does_not_change_results():
    """AUTO with and without ED_PLAN_HALF_FAST on the same inputs: bit-identical outputs (different kernels)."""
    for mode, R1 in (("plain", 1), ("rrg", 1), ("rrg", 8)):
        a = _epilogue_case(GEOS[0], mode, torch.bfloat16, R1=R1)
        b = _epilogue_case(GEOS[0], mode, torch.bfloat16, R1=R1, plan_flags=0)
        assert torch.equal(a, b)


@pytest.mark.parametrize("cfg", [GEOS[0], GEOS[3], GEOS[4], GEOS[5], GEOS[8], GEOS[9]])
@pytest.mark.parametrize("mode", ["plain", "renoise", "rrg"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("epilogue_kernel", ["direct"], indirect=True)    # the generic PEER kernel (AUTO would take the half
def test_wave_epilogue_peer_matches_spec_on_one_gpu(cfg, mode, dtype, world, epilogue_kernel):   # kernels where flagged)
    """ed_wave_epilogue_peer (the PEER = true instantiation that the N-GPU pipeline launches) with world in {2, 3, 8}:
    ragged `per`, idle ranks (world 8 with 12 samples -> ranks 6, 7 own nothing), bit-exact against the same spec."""
    _epilogue_case(cfg, mode, dtype, R1=(1 if mode == "rrg" and cfg[0] == 2 else 4), peer_world=world)


def test_wave_epilogue_peer_cfg3_partition():
    """the exact partitions of the cfg3 bench at 8 GPUs: wave 1 = 20 samples (per 3, rank 7 idle beyond 2), wave 2 = 6."""
    _epilogue_case(GEOS[0], "renoise", torch.bfloat16, R1=8, peer_world=8)
    _epilogue_case(GEOS[0], "rrg", torch.bfloat16, R1=1, peer_world=8)
    _epilogue_case(GEOS[0], "plain", torch.bfloat16, R1=1, peer_world=4)


@pytest.mark.parametrize("cfg,R1,dtype", [(GEOS[0], 8, torch.bfloat16),      # cfg3 wave 1
                                          (GEOS[1], 21, torch.float32),      # signature default resampling_steps = 20, fp32 boxes
                                          (GEOS[1], 21, torch.float16),
                                          (GEOS[3], 8, torch.float16),       # cfg4
                                          ((16, 128, 256, 128, (64, 128), 64), 8, torch.bfloat16)])   # batch large enough for the 256-thread CTA
@pytest.mark.parametrize("mode", ["renoise", "rrg"])
@pytest.mark.parametrize("epilogue_kernel", ["staged"], indirect=True)
def test_staged_epilogue_many_iterations_and_large_ctas(cfg, R1, dtype, mode, epilogue_kernel):
    _epilogue_case(cfg, mode, dtype, R1=R1)


def test_auto_mode_takes_the_direct_kernel_where_staging_does_not_apply():
    """W % 4 != 0 is outside the staged kernel's domain: forced staging must say so, AUTO must still produce the result."""
    L = native.lib()
    cfg = (1, 128, 250, 128, (64, 125), 64)
    geo = build(cfg)
    plan, keep = upload(geo)
    R1 = 2
    x = torch.randn(1, 4, 128, 250, device=DEV)
    idx = rand_idx(R1, geo.lh * geo.lw, 5)
    n = 2 * R1 + geo.nv
    out = torch.randn(n, 4, 128, 128, device=DEV)
    prm = dict(guidance=7.5, sqrt_beta_t=0.9637, sqrt_alpha_t=0.2669, sqrt_alpha_prev=0.3316, sqrt_dir=0.9434,
               rrg_weight=0.0, rrg_norm=0.0, flags=0, n_renoise=0, R1=R1)
    sp = native.StepParams(**prm)
    for k in ("guidance", "sqrt_beta_t", "sqrt_alpha_t", "sqrt_alpha_prev", "sqrt_dir"):
        prm[k] = float(getattr(sp, k))
    d_prm = torch.empty(ctypes.sizeof(native.StepParams), dtype=torch.uint8, device=DEV)
    st = native.stream_handle()
    native.check(L.ed_upload_step_params(native.ptr(d_prm), ctypes.byref(sp), st))
    owner = torch.empty(geo.H * geo.W, dtype=torch.uint8, device=DEV)
    native.check(L.ed_owner_map(ctypes.byref(plan), R1, native.ptr(idx), native.ptr(owner), st))
    y = torch.empty_like(x)
    args = (ctypes.byref(plan), native.ptr(d_prm), R1, native.ptr(x), native.ptr(out), native.ED_F32, native.ptr(idx),
            native.ptr(owner), None, native.ptr(y), None, st)
    try:
        native.check(L.ed_set_epilogue_mode(native.EPILOGUE_STAGED))
        assert L.ed_wave_epilogue(*args) == -2          # ED_ERR_UNSUPPORTED
    finally:
        native.check(L.ed_set_epilogue_mode(native.EPILOGUE_AUTO))
    native.check(L.ed_wave_epilogue(*args))
    torch.cuda.synchronize()
    want, _ = ws.spec_epilogue(geo, prm, x, out, idx, None)
    assert torch.equal(y, want)
    assert L.ed_set_epilogue_mode(7) == -1


def test_renoise_kernel_matches_sequential_axpy_and_is_linear():
    L = native.lib()
    n_re, numel = 20, 4 * 128 * 256
    x = torch.randn(numel, device=DEV)
    noise = torch.randn(n_re, numel, device=DEV)
    sp = native.StepParams(n_renoise=n_re)
    for k in range(n_re):
        sp.renoise_a[k], sp.renoise_b[k] = 0.995 - 1e-4 * k, 0.1 + 1e-3 * k
    d_prm = torch.empty(ctypes.sizeof(native.StepParams), dtype=torch.uint8, device=DEV)
    st = native.stream_handle()
    native.check(L.ed_upload_step_params(native.ptr(d_prm), ctypes.byref(sp), st))
    y = torch.empty_like(x)
    native.check(L.ed_renoise(native.ptr(d_prm), native.ptr(x), native.ptr(noise), native.ptr(y), numel, st))
    want = x.clone()
    for k in range(n_re):
        want = torch.tensor(float(sp.renoise_a[k])) * want + torch.tensor(float(sp.renoise_b[k])) * noise[k]
    assert torch.equal(y, want)
    # size-independent property at an L2-exceeding size: with zero noise the op is a pure scale by prod(a_k)
    big = torch.randn(64 * 1024 * 1024, device=DEV)            # 256 MiB
    z = torch.zeros(1, device=DEV).expand(n_re, big.numel())
    sp1 = native.StepParams(n_renoise=1)
    sp1.renoise_a[0], sp1.renoise_b[0] = 0.5, 0.25
    native.check(L.ed_upload_step_params(native.ptr(d_prm), ctypes.byref(sp1), st))
    out = torch.empty_like(big)
    native.check(L.ed_renoise(native.ptr(d_prm), native.ptr(big), native.ptr(big), native.ptr(out), big.numel(), st))
    assert torch.equal(out, big * 0.75)


@pytest.mark.parametrize("B,H,W,sample,low_vram", [(1, 256, 256, 128, False), (2, 135, 240, 128, False),
                                                   (1, 64, 128, 64, False), (1, 96, 96, 64, True)])
def test_tile_gather_and_blend_match_spec(B, H, W, sample, low_vram):
    L = native.lib()
    tg = geometry.build_tiles(H, W, sample, 8, low_vram)
    tabs = {k: torch.tensor(v, dtype=torch.int32, device=DEV) for k, v in tg.tables.items()}
    torch.manual_seed(2)
    z = torch.randn(B, 4, H, W, device=DEV)
    T = tg.core + 2 * tg.pad
    nt = len(tg.tiles)
    boxes = torch.full((nt * B, 4, T, T), float("nan"), device=DEV)
    st = native.stream_handle()
    native.check(L.ed_tile_gather(native.ptr(z), B, 4, H, W, native.ptr(tabs["tiles"]), nt, tg.core, tg.pad,
                                  native.ptr(boxes), st))
    torch.cuda.synchronize()
    assert torch.equal(boxes, ws.spec_tile_gather(z, tg.tiles, tg.core, tg.pad))
    scale = 2 if H >= 200 else 8          # keep the synthetic "decoded" patches small for the big case
    patches = torch.randn(nt * B, 3, T * scale, T * scale, device=DEV) * 1.5
    image = torch.empty(B, 3, H * scale, W * scale, device=DEV)
    tt = native.Tiles(ntiles=nt, ntc=tg.ntc, core=tg.core, pad=tg.pad, scale=scale, B=B, CH=3, H=H, W=W,
                      tiles=tabs["tiles"].data_ptr(), trow_first=tabs["trow_first"].data_ptr(),
                      trow_cnt=tabs["trow_cnt"].data_ptr(), tcol_first=tabs["tcol_first"].data_ptr(),
                      tcol_cnt=tabs["tcol_cnt"].data_ptr())
    native.check(L.ed_tile_blend(ctypes.byref(tt), native.ptr(patches), native.ED_F32, native.ptr(image), st))
    torch.cuda.synchronize()
    want = ws.spec_tile_blend(patches, tg.tiles, B, H, W, tg.core, tg.pad, scale)
    assert torch.equal(image, want)
    # tiles sharded over ranks, exercised on one GPU: ed_tile_blend_peer with the pointer table pointing at separate local
    # buffers (patch p lives in buffer p // per), ragged last rank and idle ranks; for every dtype of decoded patches
    for world, dt in ((2, torch.float32), (3, torch.bfloat16), (8, torch.float16)):
        pd = patches.to(dt)
        n = nt * B
        per = (n + world - 1) // world
        bufs = []
        for r in range(world):
            buf = torch.full((per,) + tuple(pd.shape[1:]), float("nan"), device=DEV, dtype=dt)
            lo, hi = min(r * per, n), min((r + 1) * per, n)
            buf[:hi - lo] = pd[lo:hi]
            bufs.append(buf)
        ptrs = torch.tensor([b_.data_ptr() for b_ in bufs], dtype=torch.int64, device=DEV)
        img2 = torch.full_like(image, float("nan"))
        native.check(L.ed_tile_blend_peer(ctypes.byref(tt), native.ptr(ptrs), world, per, native.dtype_code(dt), native.ptr(img2), st))
        torch.cuda.synchronize()
        assert torch.equal(img2, ws.spec_tile_blend(pd, tg.tiles, B, H, W, tg.core, tg.pad, scale)), (world, dt)
    # the "nccl" exchange of the sharded decode all-gathers centre crops and blends with pad = 0: same image
    p0, c = tg.pad * scale, tg.core * scale
    crops = patches[:, :, p0:p0 + c, p0:p0 + c].contiguous()
    tt0 = native.Tiles(ntiles=nt, ntc=tg.ntc, core=tg.core, pad=0, scale=scale, B=B, CH=3, H=H, W=W,
                       tiles=tabs["tiles"].data_ptr(), trow_first=tabs["trow_first"].data_ptr(),
                       trow_cnt=tabs["trow_cnt"].data_ptr(), tcol_first=tabs["tcol_first"].data_ptr(),
                       tcol_cnt=tabs["tcol_cnt"].data_ptr())
    img3 = torch.full_like(image, float("nan"))
    native.check(L.ed_tile_blend(ctypes.byref(tt0), native.ptr(crops), native.ED_F32, native.ptr(img3), st))
    torch.cuda.synchronize()
    assert torch.equal(img3, want)


def test_view_gather_roundtrip_at_l2_exceeding_size():
    """BASELINE full-size property: gather (TMA) then first-writer scatter of the same data reproduces the latent
    (windows tile the latent exactly) - checked on a 1 GiB working set (B=128 SDXL 1024x2048 latents... in chunks)."""
    L = native.lib()
    B = 256                                              # 256 x 512 KiB = 128 MiB latent, 256 MiB of view crops
    geo = geometry.build_geometry(B, 4, 128, 256, 128, (64, 128), 64, 64, 64)
    plan, keep = upload(geo)
    x = torch.randn(B, 4, 128, 256, device=DEV)
    canvas = torch.empty(geo.nv * B, 4, 128, 128, device=DEV)
    native.check(L.ed_gather_views(ctypes.byref(plan), native.ptr(x), native.ptr(canvas), native.ED_F32, 0,
                                   native.stream_handle()))
    torch.cuda.synchronize()
    vt = geo.tables["views"]
    for v in range(geo.nv):
        h0, h1, w0, w1, r0, c0, n_t, n_l = vt[v * 8:v * 8 + 8]
        assert torch.equal(canvas[v * B:(v + 1) * B], x[:, :, r0:r0 + 128, c0:c0 + 128])
        assert torch.equal(canvas[v * B:(v + 1) * B, :, n_t:n_t + (h1 - h0), n_l:n_l + (w1 - w0)], x[:, :, h0:h1, w0:w1])


def test_bad_arguments_are_rejected():
    L = native.lib()
    geo = build(GEOS[0])
    plan, keep = upload(geo)
    x = torch.randn(1, 4, 128, 256, device=DEV)
    assert L.ed_gather_views(ctypes.byref(plan), None, native.ptr(x), 0, 0, native.stream_handle()) == -1
    assert L.ed_gather_views(ctypes.byref(plan), native.ptr(x), native.ptr(x), 7, 0, native.stream_handle()) == -1
    assert L.ed_renoise(None, native.ptr(x), native.ptr(x), native.ptr(x), 16, native.stream_handle()) == -1
    # strips missing although the plan needs padding
    none = (ctypes.c_void_p * 4)()
    idx = rand_idx(1, geo.lh * geo.lw, 0)
    canvas = torch.empty(2, 4, 128, 128, device=DEV)
    assert L.ed_random_pick_gather(ctypes.byref(plan), 1, native.ptr(x), native.ptr(idx), none, native.ptr(canvas), 0,
                                   native.stream_handle()) == -1


@pytest.mark.parametrize("cfg", [GEOS[0], GEOS[1], GEOS[5], GEOS[7]])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gather_cond_matches_spec(cfg, dtype):
    """K12: ControlNet condition batch (zero-padded global condition, nearest-upsampled view crops)."""
    L = native.lib()
    geo = build(cfg)
    plan, keep = upload(geo)
    window = cfg[5]
    rm, cm, org = geometry.cond_geometry(geo, 8, window, cfg[3] - window)
    tabs = [torch.tensor(v, dtype=torch.int32, device=DEV) for v in (rm, cm, org)]
    R1 = 2
    cond = torch.rand(2, 3, geo.lh * 8, geo.lw * 8, device=DEV)
    n = 2 * geo.B * R1 + geo.nv * geo.B
    out = torch.full((n, 3, geo.native * 8, geo.native * 8), float("nan"), device=DEV, dtype=dtype)
    native.check(L.ed_gather_cond(ctypes.byref(plan), R1, native.ptr(cond), 3, 8, native.ptr(tabs[0]), native.ptr(tabs[1]),
                                  native.ptr(tabs[2]), native.ptr(out), native.dtype_code(dtype), native.stream_handle()))
    torch.cuda.synchronize()
    want = ws.spec_gather_cond(geo, R1, cond, 8, rm, cm, org).to(dtype)
    assert torch.equal(out, want)
