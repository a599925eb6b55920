"""GPU: the product's generate_image / denoise (CUDA kernels through the C ABI) against
  (a) the committed goldens of the unmodified reference (CPU fp32) - product run with rng_device=cpu so that the
      random draws are the reference's CPU draws; tolerance = BASELINE.json north_star's 1e-3 latent MSE, and a much
      tighter practical bound;
  (b) the oracle port executed eagerly on the same CUDA device (the reference's own PyTorch GPU path restated), with
      the default device RNG - checks the draw order for Philox streams and the fp16-autocast semantics."""
import pytest
import torch

from conftest import (PKG, cn_golden_names, condition_tensor, golden_names, load_golden, make_ed, oracle_kwargs,
                      oracle_models, scheduler_kw)
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu
NOBAR = dict(progress=lambda it: it)


@pytest.mark.parametrize("name", golden_names())
def test_cuda_path_reproduces_reference_goldens(name):
    g = load_golden(name)
    ed = make_ed(g["sd_version"], g["view_batch_size"], "cuda")
    ed.rng_device = torch.device("cpu")
    ed.autocast = False                       # goldens are CPU fp32
    ed.seed_everything(g["seed"])
    lat, _ = ed.denoise(**oracle_kwargs(g["kwargs"]), **scheduler_kw(g["kwargs"], "product"), **NOBAR)
    ref = g["latent"].cuda()
    mse = torch.mean((lat - ref) ** 2).item()
    assert mse < 1e-3, f"latent MSE {mse:.3e} exceeds the north-star tolerance 1e-3"
    assert mse < 1e-8 and (lat - ref).abs().max().item() < 5e-4, f"mse {mse:.3e} max {(lat - ref).abs().max().item():.3e}"
    assert ed.last_run["kernel_launches"] >= 3 * g["kwargs"]["num_inference_steps"]


def test_generate_image_returns_pil_and_matches_reference_image_stats():
    import numpy as np
    for name in ("sd21_512x1024_T4_R4", "xl_2048x2048_T2_R2_tiled"):
        g = load_golden(name)
        ed = make_ed(g["sd_version"], g["view_batch_size"], "cuda")
        ed.rng_device = torch.device("cpu")
        ed.autocast = False
        ed.seed_everything(g["seed"])
        imgs, log = ed.generate_image(**g["kwargs"], **NOBAR)
        assert len(imgs) == 1 and imgs[0].size == tuple(g["image_size"]) and log == {}
        a = torch.from_numpy(np.asarray(imgs[0]).copy()).float().permute(2, 0, 1) / 255.0
        stats = torch.nn.functional.adaptive_avg_pool2d(a[None], 16)[0]
        assert torch.allclose(stats, g["image_stats"][0], atol=2e-3)


@pytest.mark.parametrize("sd,H,W,T,R,vb", [("2.1", 512, 1024, 3, 3, 8), ("XL1.0", 1024, 2048, 2, 3, 16),
                                            ("XL1.0", 2048, 2048, 2, 2, 16), ("1.5", 512, 512, 3, 2, 1)])
def test_device_rng_mode_matches_oracle_on_cuda_fp32(sd, H, W, T, R, vb):
    kw = dict(prompts="a cat", negative_prompts="blurry", height=H, width=W, num_inference_steps=T,
              resampling_steps=R, cosine_scale=10.0)
    m = oracle_models(sd, vb, "cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rp.seed_all(3, "cuda")
    with torch.autocast("cuda", enabled=False):
        pass
    # the oracle's loop enables autocast like the reference; run it in fp32 by patching the flag through a no-autocast UNet
    ref = _oracle_cuda(m, kw, autocast=False)
    ed = make_ed(sd, vb, "cuda")
    ed.autocast = False
    ed.seed_everything(3)
    lat, _ = ed.denoise(**kw, **NOBAR)
    mse = torch.mean((lat - ref) ** 2).item()
    assert mse < 1e-8, f"mse {mse:.3e} max {(lat - ref).abs().max().item():.3e}"


def test_fp16_autocast_semantics_match_oracle_on_cuda():
    sd, vb = "XL1.0", 16
    kw = dict(prompts="a cat", negative_prompts="blurry", height=1024, width=2048, num_inference_steps=3,
              resampling_steps=3, cosine_scale=10.0)
    m = oracle_models(sd, vb, "cuda")
    rp.seed_all(4, "cuda")
    ref = _oracle_cuda(m, kw, autocast=True)
    ed = make_ed(sd, vb, "cuda")
    ed.seed_everything(4)                      # autocast default True, like the reference on CUDA
    lat, _ = ed.denoise(**kw, **NOBAR)
    mse = torch.mean((lat - ref) ** 2).item()
    assert mse < 1e-3, f"fp16 path: mse {mse:.3e}"


def _oracle_cuda(m, kw, autocast):
    """oracle.reference_port.denoise on CUDA; autocast=False wraps the UNet so that it runs outside autocast."""
    if not autocast:
        inner = m.unet

        class NoAutocast(torch.nn.Module):
            config = inner.config

            def __getattr__(self, k):
                return getattr(inner, k)

            def forward(self, *a, **k):
                with torch.autocast("cuda", enabled=False):
                    return inner(*a, **k)

        m.unet = NoAutocast()
    return rp.denoise(m, **kw)


def test_cuda_graph_and_lookahead_mode_is_bit_identical_to_eager():
    """use_cuda_graphs replays each wave's UNet forward from static buffers; results must not change."""
    g = load_golden("xl_1024x2048_T3_R7")
    outs = []
    for graphs in (False, True):
        ed = make_ed(g["sd_version"], g["view_batch_size"], "cuda")
        ed.rng_device = torch.device("cpu")
        ed.autocast = False
        ed.use_cuda_graphs = graphs
        ed.seed_everything(g["seed"])
        lat, _ = ed.denoise(**oracle_kwargs(g["kwargs"]), **NOBAR)
        outs.append(lat.clone())
    assert torch.equal(outs[0], outs[1])
    assert torch.mean((outs[1] - g["latent"].cuda()) ** 2).item() < 1e-8


def test_bf16_unet_input_path_stays_within_north_star_tolerance():
    """bench configuration: gather kernels write the UNet batch in bf16 (inputs rounded to bf16, UNet fp32 stub)."""
    g = load_golden("sd21_512x1024_T4_R4")
    ed = make_ed(g["sd_version"], g["view_batch_size"], "cuda")
    ed.rng_device = torch.device("cpu")
    ed.autocast = False
    ed.unet_input_dtype = torch.bfloat16
    ed.unet = ed.unet.to(torch.bfloat16)
    ed.seed_everything(g["seed"])
    lat, _ = ed.denoise(**oracle_kwargs(g["kwargs"]), **NOBAR)
    mse = torch.mean((lat - g["latent"].cuda()) ** 2).item()
    assert mse < 1e-3, mse            # BASELINE.json north_star tolerance (bf16 rounding of UNet inputs/outputs)


@pytest.mark.parametrize("low_vram", [False, True])
def test_product_tiled_decode_matches_oracle(low_vram):
    g = load_golden("xl_1080x1920_T2_R2")           # ragged: last tile row / column shifted back -> overlapping tiles
    ed = make_ed("XL1.0", 4, "cuda")
    ed.low_vram = low_vram
    z = g["latent"].cuda()
    img = ed.tiled_decode(z)
    m = oracle_models("XL1.0", 4, "cuda")
    want = rp.decode_tiled(m, z, low_vram=low_vram)
    assert img.shape == want.shape
    assert (img - want).abs().max().item() < 1e-5


def test_dedup_decode_geometry_is_the_same_algorithm_with_larger_cores():
    """SURVEY 8 row f1 (opt-in, default off): `decode_tile_geometry = (core, pad)` runs the reference's blend algorithm
    (ed:287-308) on a different tile grid through the same kernels; checked against the torch restatement of that tiling.
    With the stub VAE (no normalisation layers) and a halo covering its receptive field the image equals the reference
    tiling's; with a GroupNorm decoder it does not (scripts/decode_dedup_study.py, profiles/r2_decode_dedup.json)."""
    import torch.nn.functional as F
    ed = make_ed("XL1.0", 4, "cuda")
    z = torch.randn(1, 4, 256, 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    ref_img = ed.tiled_decode(z)                                   # reference tiles: core 32, pad 48 -> 64 tiles
    n_ref = ed.last_run["decode_tiles"]
    ed.decode_tile_geometry = (64, 16)                              # 16 tiles of 96^2 latent: 2.25x instead of 16x decoded
    ed.last_run = {}
    img = ed.tiled_decode(z)
    assert (n_ref, ed.last_run["decode_tiles"]) == (64, 16)
    core, pad, s = 64, 16, 8
    zp = F.pad(z, (pad, pad, pad, pad))
    want = torch.zeros_like(img)
    for h0 in range(0, 256, core):
        for w0 in range(0, 256, core):
            dec = ed.decode_latents(zp[:, :, h0:h0 + core + 2 * pad, w0:w0 + core + 2 * pad])
            want[:, :, h0 * s:(h0 + core) * s, w0 * s:(w0 + core) * s] = dec[:, :, pad * s:(pad + core) * s, pad * s:(pad + core) * s]
    assert (img - want).abs().max().item() < 1e-5
    assert (img - ref_img).abs().max().item() < 1e-5               # stub VAE: receptive field << halo, no GroupNorm


def test_verbose_mode_logs_intermediate_x0_grid():
    ed = make_ed("2.1", 4, "cuda")
    ed.verbose, ed.log_freq, ed.autocast = True, 1, False
    ed.seed_everything(0)
    imgs, log = ed.generate_image("a", "b", height=512, width=768, num_inference_steps=2, resampling_steps=1, **NOBAR)
    assert "intermediate_x0_imgs" in log and imgs[0].size == (768, 512)
    assert {"global_img", "global_img_inter_x0_imgs", "intermediate_cascade_x0_imgs"} <= set(log)


def test_verbose_image_log_matches_the_reference():
    """ed:1093-1118: global_img (a plain run on the first low-res latent, ed:759-796), its intermediate x0 grid, the
    per-step x0 grid and the RRG reference x0 grid - same keys, same image sizes, same pixels (up to the 8-bit rounding of
    values that differ in the last float bits) as the unmodified reference run on CPU with the same seeds."""
    import numpy as np
    from oracle.ddim_restated import DDIMRestated
    from oracle.ref_shim import build_reference, reference_available, run_reference
    if not reference_available():
        pytest.skip("unmodified reference not installed (baseline/_ref)")
    from conftest import components
    kw = dict(prompts="a cat", negative_prompts="blurry", height=512, width=1024, num_inference_steps=3, resampling_steps=2,
              guidance_scale=10.0, new_p=0.3, rrg_stop_t=0.2, rrg_init_weight=1000, cosine_scale=10.0, repaint_sampling=True)
    unet, vae, txt, proj = components("2.1")
    o = build_reference(unet, vae, DDIMRestated(), txt, sd_version="2.1", view_batch_size=4, verbose=True)
    o.log_freq = 1
    o.seed_everything(0)
    ref_imgs, ref_log, _ = run_reference(o, progress=lambda it: it, **kw)
    ed = make_ed("2.1", 4, "cuda")
    ed.verbose, ed.log_freq, ed.autocast, ed.rng_device = True, 1, False, torch.device("cpu")
    ed.seed_everything(0)
    imgs, log = ed.generate_image(**kw, **NOBAR)
    assert set(log) == set(ref_log) and set(log["intermediate_cascade_x0_imgs"]) == set(ref_log["intermediate_cascade_x0_imgs"])

    def close(a, b, what):
        assert a.size == b.size, (what, a.size, b.size)
        d = np.abs(np.asarray(a).astype(np.int32) - np.asarray(b).astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.01, (what, d.max(), (d > 0).mean())
    close(imgs[0], ref_imgs[0], "image")
    for k in ("global_img", "global_img_inter_x0_imgs", "intermediate_x0_imgs"):
        close(log[k], ref_log[k], k)
    close(log["intermediate_cascade_x0_imgs"]["rrg"], ref_log["intermediate_cascade_x0_imgs"]["rrg"], "cascade")


@pytest.mark.parametrize("name", cn_golden_names())
def test_controlnet_twin_cuda_path_reproduces_reference_goldens(name):
    g = load_golden(name)
    ed = make_ed(g["sd_version"], g["view_batch_size"], "cuda", controlnet=True)
    ed.rng_device = torch.device("cpu")
    ed.autocast = False
    ed.seed_everything(g["seed"])
    lat, _ = ed.denoise(**oracle_kwargs(g["kwargs"]), condition_image=condition_tensor(g, g["sd_version"]),
                        controlnet_conditioning_scale=g["kwargs"]["controlnet_conditioning_scale"], **NOBAR)
    mse = torch.mean((lat - g["latent"].cuda()) ** 2).item()
    assert mse < 1e-8, f"mse {mse:.3e}"


def test_controlnet_twin_generate_image_accepts_pil_condition():
    from PIL import Image
    import numpy as np
    ed = make_ed("2.1", 4, "cuda", controlnet=True)
    ed.autocast = False
    ed.seed_everything(0)
    pil = Image.fromarray((np.random.RandomState(0).rand(300, 500, 3) * 255).astype("uint8"))
    imgs, log = ed.generate_image("a", "b", pil, height=512, width=768, num_inference_steps=2, resampling_steps=1,
                                  controlnet_conditioning_scale=0.7, **NOBAR)
    assert imgs[0].size == (768, 512)
