"""ControlNet twin: drop-in for the class in /root/reference/elastic_diffusion_w_controlnet.py ("cn:N").

Same hot path as `pipeline.ElasticDiffusion` (the reference file is a copy of elastic_diffusion.py with the condition
image threaded through every UNet call); the deltas live in `pipeline.denoise` (condition batch
built once per call by `ed_gather_cond`, ControlNet forward inside the wave's batched UNet call).  Constructor and
`generate_image` signatures are the twin's own (cn:119-124, 1120-1134): `controlnet_model` is the third constructor
argument, `condition_image` the third `generate_image` argument, plus `controlnet_conditioning_scale`.
"""
from __future__ import annotations

import numpy as np
import torch

from .pipeline import (MODEL_KEYS, ConstScheduler, CosineScheduler, ElasticDiffusion as _Base, LinearScheduler, TimeIt,
                       _grid, _to_pil, timelog, tqdm)

CONTROLNET_KEYS = {("depth", True): "diffusers/controlnet-depth-sdxl-1.0", ("depth", False): "lllyasviel/sd-controlnet-depth",
                   ("canny", True): "diffusers/controlnet-canny-sdxl-1.0", ("canny", False): "lllyasviel/sd-controlnet-canny"}


class ElasticDiffusion(_Base):
    def __init__(self, device, sd_version='2.0',
                 controlnet_model='canny',
                 verbose=False,
                 log_freq=5,
                 view_batch_size=1,
                 low_vram=False):
        super().__init__(device, sd_version, verbose, log_freq, view_batch_size, low_vram)   # raises without diffusers
        from diffusers.models import ControlNetModel
        self.controlnet_model = controlnet_model
        key = CONTROLNET_KEYS.get((controlnet_model, self.sd_version == 'XL1.0'), controlnet_model)   # cn:172-189
        self.controlnet = ControlNetModel.from_pretrained(key, torch_dtype=self.torch_dtype).to(
            'cpu' if self.low_vram else self.device)
        if controlnet_model == 'depth':
            from transformers import pipeline
            self.depth_estimator = pipeline('depth-estimation')
        print('[INFO] loaded ControlNet!')

    def process_condition_image(self, condition_image, controlnet_model):
        """cn:1102-1117 (canny / depth pre-processing of a PIL image; outside the hot path)."""
        from PIL import Image
        assert controlnet_model in ['canny', 'depth'], f"processing for ControlNet: {controlnet_model} is not implemented"
        if controlnet_model == 'canny':
            import cv2
            edges = cv2.Canny(np.array(condition_image), 100, 200)[:, :, None]
            return Image.fromarray(np.concatenate([edges] * 3, axis=2))
        depth = np.array(self.depth_estimator(condition_image)['depth'])[:, :, None]
        return Image.fromarray(np.concatenate([depth] * 3, axis=2))

    def prepare_image(self, image, width, height, batch_size=1, num_images_per_prompt=1, device=None, dtype=None,
                      do_classifier_free_guidance=False, guess_mode=False):
        """cn:1005-1033.  Accepts a PIL image / array (resized, scaled to [0,1]) or an already prepared tensor."""
        if not torch.is_tensor(image):
            from PIL import Image
            if not isinstance(image, Image.Image):
                image = Image.fromarray(np.asarray(image))
            image = image.convert("RGB").resize((width, height), Image.LANCZOS)     # VaeImageProcessor default resample
            image = torch.from_numpy(np.asarray(image).copy()).float().div(255.0).permute(2, 0, 1)[None]
        image = image.to(dtype=torch.float32)
        if tuple(image.shape[-2:]) != (height, width):
            image = torch.nn.functional.interpolate(image, size=(height, width), mode="bilinear", align_corners=False)
        image = image.repeat_interleave(batch_size if image.shape[0] == 1 else num_images_per_prompt, dim=0)
        image = image.to(device=device or self.device, dtype=dtype or torch.float32)
        if do_classifier_free_guidance and not guess_mode:
            image = torch.cat([image] * 2)
        return image.to(self.device)

    @torch.no_grad()
    def generate_image(self, prompts, negative_prompts='',
                       condition_image=None,
                       height=768, width=768,
                       num_inference_steps=50,
                       guidance_scale=10.0,
                       controlnet_conditioning_scale=1.0,
                       resampling_steps=20,
                       new_p=0.3, rrg_stop_t=0.2,
                       rrg_init_weight=1000,
                       rrg_scherduler_cls=CosineScheduler,
                       cosine_scale=3.0,
                       repaint_sampling=True,
                       progress=tqdm,
                       tiled_decoder=False,
                       grid=False):
        ds = self.get_downsample_size(height, width)
        sf = self.vae_scale_factor
        cond = self.prepare_image(condition_image, width=ds[1] * sf, height=ds[0] * sf, batch_size=1,
                                  num_images_per_prompt=1, device=self.device,
                                  dtype=next(self.controlnet.parameters()).dtype)          # cn:1183-1193 (doubling inside)
        latent, image_log = self.denoise(prompts, negative_prompts, height, width, num_inference_steps, guidance_scale,
                                         resampling_steps, new_p, rrg_stop_t, rrg_init_weight, rrg_scherduler_cls,
                                         cosine_scale, repaint_sampling, progress, condition_image=cond,
                                         controlnet_conditioning_scale=controlnet_conditioning_scale)
        needs_upcasting = self.vae.dtype == torch.float16 and self.vae.config.force_upcast
        if self.low_vram:
            self.unet.cpu()
            self.vae.to(self.device)
        if needs_upcasting:
            self.upcast_vae()
        decode_fn = self.tiled_decode if tiled_decoder else self.decode_latents
        imgs = torch.cat([decode_fn(latent[i:i + 1]) for i in range(len(latent))])
        if grid:
            imgs = [_grid(imgs)]
        imgs = [_to_pil(img) for img in imgs]
        if needs_upcasting:
            self.vae.to(dtype=torch.float16)
        return imgs, image_log


__all__ = ["ElasticDiffusion", "CosineScheduler", "LinearScheduler", "ConstScheduler", "TimeIt", "timelog", "MODEL_KEYS"]
