"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/elastic_diffusion.py) through
oracle/ref_shim.py on CPU fp32 with the synthetic StubUNet / StubVAE (fixed seeds).

Run in the build container only (the reference does not exist on the GPU box):
    python scripts/make_golden.py
Each fixture holds the call's kwargs, the final latent handed to the VAE (ed:1121) and coarse image statistics.
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import standins as syn  # noqa: E402
from oracle.ddim_restated import DDIMRestated  # noqa: E402
from oracle.ref_shim import build_reference, run_reference  # noqa: E402

# name -> (sd_version, view_batch_size, generate_image kwargs).  BASELINE.json configs at reduced step counts.
CASES = {
    "sd15_512x512_T5_R3": ("1.5", 1, dict(height=512, width=512, num_inference_steps=5, resampling_steps=3)),
    "sd15_512x512_T3_R0": ("1.5", 1, dict(height=512, width=512, num_inference_steps=3, resampling_steps=0)),
    "sd21_512x1024_T4_R4": ("2.1", 8, dict(height=512, width=1024, num_inference_steps=4, resampling_steps=4)),
    "sd21_512x1024_B2_T3_R2": ("2.1", 2, dict(height=512, width=1024, num_inference_steps=3, resampling_steps=2,
                                               prompts=["a cat", "a dog on a bench"])),
    "sd21_640x896_T3_R2_norepaint": ("2.1", 2, dict(height=640, width=896, num_inference_steps=3, resampling_steps=2,
                                                    repaint_sampling=False, cosine_scale=3.0)),
    # non-default new_p whose threshold 100 * (1 - new_p) = 19.999999999999996 is not representable (float32 compare, ed:542)
    "sd21_512x1024_T3_R3_newp08": ("2.1", 8, dict(height=512, width=1024, num_inference_steps=3, resampling_steps=3,
                                                  new_p=0.8)),
    # the two other RRG weight schedulers of the public signature (`rrg_scherduler_cls`, ed:73-94, 972-979): 5 steps so that
    # the weight crosses the `> 10` gate (ed:1062) and the linear ramp / constant plateau differ from the cosine
    "sd21_512x1024_T5_R2_linear": ("2.1", 4, dict(height=512, width=1024, num_inference_steps=5, resampling_steps=2,
                                                  rrg_scheduler="linear", rrg_stop_t=0.4)),
    "xl_1024x2048_T4_R1_const": ("XL1.0", 16, dict(height=1024, width=2048, num_inference_steps=4, resampling_steps=1,
                                                   rrg_scheduler="const", rrg_stop_t=0.5, rrg_init_weight=400)),
    "xl_1024x2048_T3_R7": ("XL1.0", 16, dict(height=1024, width=2048, num_inference_steps=3, resampling_steps=7)),
    "xl_2048x2048_T2_R2_tiled": ("XL1.0", 16, dict(height=2048, width=2048, num_inference_steps=2, resampling_steps=2,
                                                   tiled_decoder=True)),
    "xl_1536x1536_T2_R3": ("XL1.0", 16, dict(height=1536, width=1536, num_inference_steps=2, resampling_steps=3)),
    "xl_1080x1920_T2_R2": ("XL1.0", 4, dict(height=1080, width=1920, num_inference_steps=2, resampling_steps=2)),
    # latent 96x256: window+context >= 96 -> the window collapses to the full height (ed:820-825) and every view is
    # 96 rows < native 128 -> background-padded local views (ed:405-408 from compute_local_uncond_signal), 2 view chunks
    "xl_768x2048_T2_R2_padded_views": ("XL1.0", 2, dict(height=768, width=2048, num_inference_steps=2, resampling_steps=2)),
    # downsample factors > 2 (the reference's "Future TODO" ed:562 - it does run, keeping 2 of every
    # 8 / 6 rows of the 2x-resized grid): factor 4 (latent 128x256 -> low-res 32x64) and factor 3 (192x192 -> 64x64)
    "sd21_1024x2048_T2_R3_factor4": ("2.1", 8, dict(height=1024, width=2048, num_inference_steps=2, resampling_steps=3)),
    "sd21_1536x1536_T2_R2_factor3": ("2.1", 4, dict(height=1536, width=1536, num_inference_steps=2, resampling_steps=2)),
}
# ControlNet twin (elastic_diffusion_w_controlnet.py): condition = fixed-seed uniform image of the prepared size
CN_CASES = {
    "cn_sd21_512x1024_T3_R3": ("2.1", 8, dict(height=512, width=1024, num_inference_steps=3, resampling_steps=3,
                                              controlnet_conditioning_scale=0.8)),
    "cn_xl_1024x2048_T2_R2": ("XL1.0", 16, dict(height=1024, width=2048, num_inference_steps=2, resampling_steps=2,
                                                controlnet_conditioning_scale=1.0)),
    "cn_xl_1080x1920_T2_R1": ("XL1.0", 4, dict(height=1080, width=1920, num_inference_steps=2, resampling_steps=1,
                                               controlnet_conditioning_scale=0.5)),
}
COND_SEED = 7
DEFAULTS = dict(prompts="a cat", negative_prompts="blurry", guidance_scale=10.0, new_p=0.3, rrg_stop_t=0.2,
                rrg_init_weight=1000, cosine_scale=10, repaint_sampling=True)
SEED = 0


def components(sd):
    xl = sd.startswith("XL")
    unet = syn.StubUNet(sample_size=128 if xl else 64, cross_dim=16, xl=xl, pooled_dim=8)
    return unet, syn.StubVAE(), syn.StubTextEncoder(16, 8 if xl else None), (8 if xl else None)


def image_stats(img):
    """3 x 16 x 16 block means of the PIL image as float tensor (cheap stand-in for the full image)."""
    import numpy as np
    a = torch.from_numpy(np.asarray(img)).float().permute(2, 0, 1) / 255.0
    return torch.nn.functional.adaptive_avg_pool2d(a[None], 16)[0]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (sd, vb, kw) in CASES.items():
        unet, vae, txt, proj = components(sd)
        o = build_reference(unet, vae, DDIMRestated(), txt, sd_version=sd, view_batch_size=vb, projection_dim=proj)
        o.seed_everything(SEED)
        args = dict(DEFAULTS)
        args.update(kw)
        call = dict(args)
        sched = call.pop("rrg_scheduler", None)          # stored by NAME in the fixture; the reference takes its own class
        if sched is not None:
            from oracle.ref_shim import load_reference_module
            mod = load_reference_module("elastic_diffusion")
            call["rrg_scherduler_cls"] = {"linear": mod.LinearScheduler, "const": mod.ConstScheduler, "cosine": mod.CosineScheduler}[sched]
        imgs, _, latent = run_reference(o, progress=lambda it: it, **call)
        torch.save(dict(sd_version=sd, view_batch_size=vb, seed=SEED, kwargs=args, latent=latent.clone(),
                        image_stats=torch.stack([image_stats(i) for i in imgs]), image_size=imgs[0].size),
                   os.path.join(out_dir, name + ".pt"))
        print(f"{name}: latent {tuple(latent.shape)} std {latent.std():.4f} images {len(imgs)} x {imgs[0].size}")


def condition_for(o_or_ds, seed=COND_SEED):
    ds = o_or_ds
    return torch.rand(1, 3, ds[0] * 8, ds[1] * 8, generator=torch.Generator().manual_seed(seed))


def main_cn():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (sd, vb, kw) in CN_CASES.items():
        unet, vae, txt, proj = components(sd)
        cn = syn.StubControlNet()
        o = build_reference(unet, vae, DDIMRestated(), txt, sd_version=sd, view_batch_size=vb, projection_dim=proj,
                            controlnet=cn)
        o.seed_everything(SEED)
        args = dict(DEFAULTS)
        args.update(kw)
        cond = condition_for(o.get_downsample_size(args["height"], args["width"]))
        imgs, _, latent = run_reference(o, progress=lambda it: it, condition_image=cond, **args)
        torch.save(dict(sd_version=sd, view_batch_size=vb, seed=SEED, kwargs=args, latent=latent.clone(), cond_seed=COND_SEED,
                        image_stats=torch.stack([image_stats(i) for i in imgs]), image_size=imgs[0].size),
                   os.path.join(out_dir, name + ".pt"))
        print(f"{name}: latent {tuple(latent.shape)} std {latent.std():.4f}")


if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if a in CASES]
    if only:
        CASES = {k: CASES[k] for k in only}
        main_cn = lambda: None
    if "--cn-only" in sys.argv:
        main_cn()
        sys.exit(0)
    main()
