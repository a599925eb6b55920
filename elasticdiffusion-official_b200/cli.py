"""Command-line entry points of the two drop-in modules - the `__main__` blocks of the reference files
(/root/reference/elastic_diffusion.py:1134-1210, elastic_diffusion_w_controlnet.py:1342-1436): same flags, same defaults,
same outputs (`<outdir>/<exp>/<time>_<seed>/{i}.png`, image-log PNGs, `args.txt`).  Note the reference declares its
switches with `type=bool`, i.e. any non-empty string (even "False") switches them ON; kept as is."""
from __future__ import annotations

import argparse
import os
from datetime import datetime

import torch

# flag, type, default (plain pipeline), default (ControlNet twin; None = same), help
_OPTIONS = [
    ("prompt", str, "A realistic portrait of a young black woman. she has a Christmas red hat and a red scarf. Her eyes are "
                    "light brown like they're almost caramel color. Her attire, simple yet dignified.",
     "Envision a dramatic picture of the joker, masterpiece. High resolution, detailed", None),
    ("negative", str, "blurry, ugly, duplicate, no details, deformed", None, None),
    ("H", int, 2048, 1536, None), ("W", int, 2048, 1536, None),
    ("low_vram", bool, False, None, "run with half percision on low memeory mode"),
    ("seed", int, 0, None, None), ("steps", int, 50, None, None), ("num_sampled", int, 1, None, None),
    ("guidance_scale", float, 10.0, None, None),
    ("cosine_scale", float, 10.0, None, "effective only with CosineScheduler"),
    ("rrg_scale", float, 4000, 2000, None), ("resampling_steps", int, 10, 7, None),
    ("new_p", float, 0.3, None, None), ("rrg_stop_t", float, 0.2, None, None), ("view_batch_size", int, 16, None, None),
    ("outdir", str, "results_log/", None, None),
    ("make_grid", bool, False, None, "make a grid of the output images"),
    ("repaint_sampling", bool, True, None, ""), ("tiled_decoder", bool, False, None, ""),
    ("exp", str, "ElasticDiffusion", "ControlNet-ElasticDiffusion", "experiment tag"),
    ("tag", str, "", None, "identifier experiment tag"),
    ("log_freq", int, 5, None, "log frequency of intermediate diffusion steps"),
    ("verbose", bool, False, None, None),
]
_TWIN_ONLY = [("controlnet_conditioning_scale", float, 0.2), ("condition_image", str, "imgs/input/yoga.jpeg"),
              ("controlnet_model", str, "depth")]


def build_parser(twin: bool) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--sd_version", type=str, default="XL1.0", choices=["1.4", "1.5", "2.0", "2.1", "XL1.0"],
                   help="stable diffusion version ['1.4', '1.5', '2.0', '2.1', or 'XL1.0'] or a model key for a huggingface "
                        "stable diffusion version")
    for name, typ, d_plain, d_twin, hlp in _OPTIONS:
        p.add_argument("--" + name, type=typ, default=d_twin if (twin and d_twin is not None) else d_plain, help=hlp)
    if twin:
        for name, typ, default in _TWIN_ONLY:
            p.add_argument("--" + name, type=typ, default=default)
    return p


def main(cls, timelog, twin=False, argv=None):
    opt = build_parser(twin).parse_args(argv)
    device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    if opt.verbose:
        timelog.sync_gpu = opt.verbose
    common = dict(verbose=opt.verbose, log_freq=opt.log_freq, view_batch_size=opt.view_batch_size, low_vram=opt.low_vram)
    sd = cls(device, opt.sd_version, opt.controlnet_model, **common) if twin else cls(device, opt.sd_version, **common)
    sd.seed_everything(opt.seed)
    extra, condition = {}, None
    if twin:
        from PIL import Image
        ds = sd.get_downsample_size(opt.H, opt.W)
        condition = Image.open(opt.condition_image).resize((ds[1] * sd.vae_scale_factor, ds[0] * sd.vae_scale_factor)).convert("RGB")
        condition = sd.process_condition_image(condition, sd.controlnet_model)
        extra = dict(condition_image=condition, controlnet_conditioning_scale=opt.controlnet_conditioning_scale)
    imgs, image_log = sd.generate_image(prompts=[opt.prompt] * opt.num_sampled, negative_prompts=opt.negative, height=opt.H,
                                        width=opt.W, num_inference_steps=opt.steps, grid=opt.make_grid,
                                        guidance_scale=opt.guidance_scale, resampling_steps=opt.resampling_steps,
                                        new_p=opt.new_p, cosine_scale=opt.cosine_scale, rrg_init_weight=opt.rrg_scale,
                                        rrg_stop_t=opt.rrg_stop_t, repaint_sampling=opt.repaint_sampling,
                                        tiled_decoder=opt.tiled_decoder, **extra)
    if opt.verbose:
        timelog.print_results()
    save_dir = os.path.join(opt.outdir, opt.exp, f"{datetime.now().strftime('%Y-%m-%d %H:%M:%S')}_{opt.seed}")
    os.makedirs(save_dir, exist_ok=True)
    if condition is not None:
        condition.save(f"{save_dir}/condition_image.png")
    for i, img in enumerate(imgs):
        img.save(f"{save_dir}/{i}.png")
    for key, val in image_log.items():
        if isinstance(val, dict):
            for label, img in val.items():
                img.save(f"{save_dir}/{key}_{label}.png")
        else:
            val.save(f"{save_dir}/{key}.png")
    with open(f"{save_dir}/args.txt", "w") as f:
        f.write("\n".join(f"{k}: {v}" for k, v in vars(opt).items()))
    return save_dir
