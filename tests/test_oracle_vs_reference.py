"""CPU, build container only: the oracle against the LIVE unmodified reference (skipped where /root/reference is
absent, e.g. on the GPU box - there the committed goldens stand in)."""
import os

import numpy as np
import pytest
import torch

from conftest import components, oracle_models
from oracle import reference_port as rp
from oracle.ddim_restated import DDIMRestated
from oracle.ref_shim import build_reference, reference_available, run_reference

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference sources not present")


def _ref(sd, vb):
    unet, vae, txt, proj = components(sd)
    return build_reference(unet, vae, DDIMRestated(), txt, sd_version=sd, view_batch_size=vb, projection_dim=proj)


@pytest.mark.parametrize("H,W,ws,stride", [(512, 512, 32, 32), (1024, 2048, 64, 64), (1080, 1920, 64, 64),
                                            (768, 1024, 32, 32), (2048, 2048, 32, 16), (640, 896, 48, 48)])
def test_view_grid(H, W, ws, stride):
    o = _ref("2.1", 1)
    assert rp.view_grid(H, W, ws, ws, stride, 8) == o.get_views(H, W, h_ws=ws, w_ws=ws, stride=stride)


@pytest.mark.parametrize("H,W", [(64, 128), (135, 240), (96, 128), (256, 256), (80, 112)])
def test_context_box_equals_crop_with_context(H, W):
    o = _ref("XL1.0", 1)
    X = torch.randn(1, 4, H, W)
    for ws, n in [(64, 32), (32, 16), (48, 8)]:
        for v in rp.view_grid(H * 8, W * 8, min(ws, H), min(ws, W), ws, 8):
            crop, ctx = o.crop_with_context(X, *v, S=1, n=n)
            (r0, r1, c0, c1), ctx2 = rp.context_box(v, n, H, W)
            assert tuple(ctx) == tuple(ctx2)
            assert torch.equal(crop, X[:, :, r0:r1, c0:c1])


@pytest.mark.parametrize("H,W,ds", [(128, 256, (64, 128)), (192, 192, (128, 128)), (135, 240, (72, 128)),
                                     (64, 64, (64, 64)), (80, 112, (45, 64)), (96, 128, (48, 64))])
def test_pick_and_mask_equals_random_nearest_downsample(H, W, ds):
    o = _ref("2.1", 1)
    o.random_downasmple_pre = {}
    X = torch.randn(1, 4, H, W)
    torch.manual_seed(3)
    tabs = rp.ResampleTables(H, W, ds)
    # nearest (top-left) pass
    low, mask, idx = o.random_nearest_downsample(X, ds, nearest=True)
    low2, mask2 = rp.pick_and_mask(X, tabs, torch.zeros(tabs.lh * tabs.lw, dtype=torch.long))
    assert torch.equal(low, low2) and torch.equal(mask, mask2[:mask.shape[0], :mask.shape[1]])
    # random pass with an exclude mask and previous indices: same RNG stream on both sides
    excl = torch.zeros(len(idx), 4, dtype=torch.bool)
    excl[torch.arange(len(idx)), idx] = True
    torch.manual_seed(11)
    low, mask, idx_r = o.random_nearest_downsample(X, ds, prev_random_indices=idx, exclude_mask=excl, drop_p=0.7)
    torch.manual_seed(11)
    idx_m = rp.mix_with_previous(rp.draw_cell_indices(len(idx), excl), idx, 0.7, "cpu")
    assert torch.equal(idx_r, idx_m)
    low2, mask2 = rp.pick_and_mask(X, tabs, idx_m)
    assert torch.equal(low, low2) and torch.equal(mask, mask2[:mask.shape[0], :mask.shape[1]])


def test_rrg_closed_form_vs_reference_autograd():
    o = _ref("2.1", 1)
    o.scheduler.set_timesteps(10)
    t = o.scheduler.timesteps[2]
    x0 = torch.randn(2, 4, 64, 128)
    lat, un, di = torch.randn(2, 4, 32, 64), torch.randn(2, 4, 32, 64), torch.randn(2, 4, 32, 64)
    g, _ = o.reduced_resolution_guidance(x0, t, None, x0, None, None, None, guidance_scale=3.3, rrg_scale=np.float64(417.3),
                                         downsample_size=(32, 64),
                                         donwsampled_scores={"latent": lat, "uncond_score": un, "direction": di})
    m = rp.Models(o.unet, o.vae, o.scheduler, None, "2.1")
    g2, _ = rp.rrg_gradient(m, t, x0, lat, un, di, 3.3, np.float64(417.3))
    assert torch.equal(g, g2)


@pytest.mark.parametrize("sd,H,W,T,R,vb", [("2.1", 512, 768, 2, 2, 3), ("XL1.0", 1024, 1536, 2, 1, 2)])
def test_end_to_end_live(sd, H, W, T, R, vb):
    o = _ref(sd, vb)
    o.seed_everything(5)
    _, _, lat = run_reference(o, prompts="a", negative_prompts="b", height=H, width=W, num_inference_steps=T,
                              resampling_steps=R, progress=lambda it: it)
    unet, vae, txt, proj = components(sd)
    m = rp.Models(unet, vae, DDIMRestated(), txt, sd, "cpu", vb, projection_dim=proj)
    rp.seed_all(5, "cpu")
    mine = rp.denoise(m, "a", "b", H, W, T, resampling_steps=R)
    assert torch.equal(mine, lat)


def test_module_level_schedulers_and_timelog_match_reference():
    """ed:33-109: the RRG weight schedules and the TimeIt accumulator callers import next to the class."""
    import importlib
    from oracle.ref_shim import load_reference_module
    ref = load_reference_module()
    mine = importlib.import_module("elastic_diffusion")
    for steps, scale, factor in [(40, 10.0, 1000), (40, 3.0, 1000), (8, 1.0, 0.01)]:
        a, b = ref.CosineScheduler(steps, scale, factor), mine.CosineScheduler(steps, scale, factor)
        assert [a(t) for t in range(steps + 3)] == [b(t) for t in range(steps + 3)]
    for cls in ("LinearScheduler", "ConstScheduler"):
        a, b = getattr(ref, cls)(10, 1000, 0), getattr(mine, cls)(10, 1000, 0)
        assert [a(t) for t in range(13)] == [b(t) for t in range(13)]
    assert hasattr(mine, "timelog") and sorted(n for n in dir(ref.timelog) if not n.startswith("_")) == \
        sorted(n for n in dir(mine.timelog) if not n.startswith("_"))

    @mine.timelog.time_function
    def f(x):
        return x + 1
    assert f(1) == 2 and "FUNCTION_f" in mine.timelog.total_time
    with mine.timelog.time_block("blk"):
        pass
    assert "BLOCK_blk" in mine.timelog.total_time


def test_error_behaviour_matches_reference_on_the_host_side():
    """ed:200-201 raises a str (= TypeError) for sizes that are not multiples of 8; ed:239-242 ValueError for an XL UNet whose
    added-embedding width disagrees with the text encoder's projection dim."""
    o = _ref("XL1.0", 1)
    unet, vae, txt, proj = components("XL1.0")
    from conftest import PKG
    ed = PKG.ElasticDiffusion.from_components("cpu", unet, vae, None, txt, sd_version="XL1.0", projection_dim=proj)
    for obj in (o, ed):
        with pytest.raises(TypeError):
            obj.get_views(1001, 512)
    assert ed.get_views(1080, 1920, 64, 64, 64) == o.get_views(1080, 1920, h_ws=64, w_ws=64, stride=64)
    good = ed._get_add_time_ids((4096, 8192), (0, 0), (4096, 8192), dtype=torch.float32)
    assert torch.equal(good, o._get_add_time_ids((4096, 8192), (0, 0), (4096, 8192), dtype=torch.float32))
    bad = PKG.ElasticDiffusion.from_components("cpu", unet, vae, None, txt, sd_version="XL1.0", projection_dim=proj + 1)
    with pytest.raises(ValueError, match="Model expects an added time embedding vector"):
        bad._get_add_time_ids((4096, 8192), (0, 0), (4096, 8192), dtype=torch.float32)
    assert ed.get_downsample_size(1024, 2048) == o.get_downsample_size(1024, 2048)
    assert ed.get_downsample_size(1080, 1920) == o.get_downsample_size(1080, 1920)


def test_sizes_not_divisible_by_8_are_floored():
    """generate_image never raises on sizes that are not multiples of 8: ed:998 draws a (height // 8, width // 8) latent and
    get_views only sees latent * 8 (ed:817, 827).  516 x 1028 runs like 512 x 1024 in the reference, in the oracle port and
    in the product's wave-form host logic (no CUDA: spec kernels)."""
    from conftest import make_ed
    from oracle import wave_spec as ws
    kw = dict(prompts="a cat", negative_prompts="blurry", guidance_scale=10.0, new_p=0.3, rrg_stop_t=0.2, rrg_init_weight=1000,
              cosine_scale=10, repaint_sampling=True, num_inference_steps=2, resampling_steps=1)
    o = _ref("2.1", 4)
    o.seed_everything(0)
    _, _, want = run_reference(o, progress=lambda it: it, height=516, width=1028, **kw)
    assert want.shape == (1, 4, 64, 128)
    m = oracle_models("2.1", 4)
    rp.seed_all(0, "cpu")
    assert torch.equal(rp.denoise(m, height=516, width=1028, **kw), want)
    ed = make_ed("2.1", 4)
    ed.seed_everything(0)
    got = ws.denoise_wave_form(ed, height=516, width=1028, **kw)
    assert (got - want).abs().max().item() <= 2e-5


@pytest.mark.parametrize("fname,twin", [("elastic_diffusion.py", False), ("elastic_diffusion_w_controlnet.py", True)])
def test_cli_flags_and_defaults_match_the_reference_command_line(fname, twin):
    """ed:1134-1161 / cn:1342-1372: every `--flag` of the reference's __main__ parser exists in the drop-in module's parser
    with the same type and default (the reference's parser lives under `if __name__ == '__main__'`, so its source text is
    read, not imported)."""
    import ast
    import importlib
    from oracle.ref_shim import reference_dir
    tree = ast.parse(open(os.path.join(reference_dir(), fname)).read())
    cli = importlib.import_module("elasticdiffusion-official_b200.cli")
    mine = {a.dest: a for a in cli.build_parser(twin)._actions if a.option_strings and a.dest != "help"}
    seen = 0
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument"):
            continue
        flag = node.args[0].value.lstrip("-")
        kws = {k.arg: k.value for k in node.keywords}
        assert flag in mine, flag
        assert mine[flag].type.__name__ == kws["type"].id, (flag, kws["type"].id, mine[flag].type)
        assert mine[flag].default == ast.literal_eval(kws["default"]), (flag, mine[flag].default)
        if "choices" in kws and flag == "sd_version":
            assert list(mine[flag].choices) == ast.literal_eval(kws["choices"])
        seen += 1
    assert seen == len(mine) == (27 if twin else 24), (seen, len(mine))


def test_process_condition_image_matches_the_twin():
    """cn:1102-1117 (canny / depth pre-processing, outside the hot path) with stand-ins for cv2.Canny and the depth
    estimator: same PIL image out of the drop-in class and the unmodified twin."""
    import sys
    import types
    from PIL import Image
    from conftest import PKG
    import standins
    fake_cv2 = sys.modules.get("cv2") or types.ModuleType("cv2")
    had = hasattr(fake_cv2, "Canny")
    if not had:
        fake_cv2.Canny = lambda img, lo, hi: (np.asarray(img)[:, :, 0] > (lo + hi) // 2).astype(np.uint8) * 255
        sys.modules["cv2"] = fake_cv2
    try:
        unet, vae, txt, proj = components("2.1")
        o = build_reference(unet, vae, DDIMRestated(), txt, sd_version="2.1", controlnet=standins.StubControlNet())
        ed = PKG.controlnet.ElasticDiffusion.from_components("cpu", unet, vae, None, txt, sd_version="2.1",
                                                             controlnet=standins.StubControlNet())
        depth = lambda im: {"depth": Image.fromarray((np.asarray(im)[:, :, 1] // 2).astype(np.uint8))}
        o.depth_estimator = ed.depth_estimator = depth
        img = Image.fromarray(np.random.RandomState(0).randint(0, 255, (64, 96, 3), dtype=np.uint8))
        for model in ("canny", "depth"):
            a, b = o.process_condition_image(img, model), ed.process_condition_image(img, model)
            assert a.size == b.size and a.mode == b.mode and np.array_equal(np.asarray(a), np.asarray(b)), model
        with pytest.raises(AssertionError):
            ed.process_condition_image(img, "pose")
    finally:
        if not had:
            del fake_cv2.Canny
