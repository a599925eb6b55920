"""TEST INFRASTRUCTURE - loads the UNMODIFIED reference (`/root/reference/elastic_diffusion.py`) read-only.

The reference imports `diffusers`, which is absent from this image; its `__init__` calls `from_pretrained`, and no
weights exist offline.  This module (recipe from SURVEY.md Appendix A) registers dummy `diffusers` modules in
`sys.modules`, `importlib`s the reference file where it lies, builds the object with `__new__` and injects the
synthetic UNet / VAE / scheduler.  Every line of the hot path that then executes is the reference's own.

/root/reference exists only in the build container: this file is used by `scripts/make_golden.py` (which writes
the committed fixtures under tests/golden/) and by CPU tests that are skipped when the reference is absent.
Nothing in `-m gpu` tests, smoke() or bench.py reads /root/reference at run time; bench.py's reference legs load the
same unmodified files from `baseline/_ref/` (installed by scripts/install_reference.py, git-ignored, shipped to the box).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import torch
import torch.nn as nn

# /root/reference exists in the build container only; baseline/_ref is the pip-installed copy of the same unmodified files
# (scripts/install_reference.py, git-ignored) that travels to the GPU box - bench.py's reference legs use it there.
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = [os.environ.get("ELASTIC_REFERENCE_DIR", ""), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")]


def reference_dir():
    for d in REF_CANDIDATES:
        if d and os.path.isfile(os.path.join(d, "elastic_diffusion.py")):
            return d
    return None


def reference_available() -> bool:
    return reference_dir() is not None


_cached = {}


def load_reference_module(name="elastic_diffusion"):
    if name in _cached:
        return _cached[name]
    d, dm, da = (types.ModuleType(n) for n in ("diffusers", "diffusers.models", "diffusers.models.attention_processor"))
    for n in ("AutoencoderKL", "UNet2DConditionModel", "DDIMScheduler", "ControlNetModel"):
        setattr(d, n, type(n, (), {}))
    for n in ("AttnProcessor2_0", "LoRAAttnProcessor2_0", "LoRAXFormersAttnProcessor", "XFormersAttnProcessor"):
        setattr(da, n, type(n, (), {}))
    setattr(dm, "ControlNetModel", d.ControlNetModel)
    dip = types.ModuleType("diffusers.image_processor")

    class VaeImageProcessor:   # only `preprocess` is used (cn:1017); tensors pass through, sizes are checked
        def __init__(self, **kw):
            pass

        def preprocess(self, image, height=None, width=None):
            assert torch.is_tensor(image) and image.shape[-2:] == (height, width), "shim: pass a (1,3,h,w) tensor"
            return image
    dip.VaeImageProcessor = VaeImageProcessor
    if "diffusers" not in sys.modules:
        sys.modules.update({"diffusers": d, "diffusers.models": dm, "diffusers.models.attention_processor": da,
                            "diffusers.image_processor": dip})
    if "cv2" not in sys.modules:            # only used by process_condition_image (canny), outside the hot path
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
    path = os.path.join(reference_dir(), name + ".py")
    spec = importlib.util.spec_from_file_location("_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # the reference calls torchvision's make_grid(..., nrows=...) (ed:1099-1124): with its pinned torchvision 0.12 the unknown
    # keyword falls into **kwargs and is ignored (nrow stays 8); the installed torchvision rejects it.  Same behaviour here:
    if hasattr(mod, "make_grid"):
        tv_make_grid = mod.make_grid
        mod.make_grid = lambda tensor, *a, nrows=None, **k: tv_make_grid(tensor, *a, **k)
    _cached[name] = mod
    return mod


def build_reference(unet, vae, scheduler, text_fn, sd_version="2.1", device="cpu", view_batch_size=1,
                    verbose=False, low_vram=False, projection_dim=None, controlnet=None):
    """Reference `ElasticDiffusion` instance with injected components (bypasses `from_pretrained`).
    With `controlnet` the class comes from the unmodified elastic_diffusion_w_controlnet.py."""
    ref = load_reference_module("elastic_diffusion_w_controlnet" if controlnet is not None else "elastic_diffusion")
    o = ref.ElasticDiffusion.__new__(ref.ElasticDiffusion)
    nn.Module.__init__(o)
    o.device = torch.device(device)
    o.sd_version = sd_version
    o.verbose = verbose
    o.torch_dtype = torch.float16 if low_vram else torch.float32
    o.view_batch_size = view_batch_size
    o.log_freq = 5
    o.low_vram = low_vram
    o.unet, o.vae, o.scheduler = unet, vae, scheduler
    if controlnet is not None:
        o.controlnet, o.controlnet_model = controlnet, "canny"
        o.control_image_processor = sys.modules["diffusers.image_processor"].VaeImageProcessor()
    o.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1)
    o.get_text_embeds = text_fn
    if projection_dim is not None:
        o.text_encoder = [None, SimpleNamespace(config=SimpleNamespace(projection_dim=projection_dim))]
    o.set_view_config()
    return o


def run_reference(o, **gen_kwargs):
    """Calls the reference's own generate_image; returns (imgs, image_log, final_latent).

    The final latent is captured at the hand-over to the decoder (ed:1121 decodes one sample at a time)."""
    chunks, depth = [], [0]
    dec, tdec = o.decode_latents, o.tiled_decode

    def _wrap(fn):
        def inner(z):
            if depth[0] == 0:
                chunks.append(z.detach().clone())
            depth[0] += 1
            try:
                return fn(z)
            finally:
                depth[0] -= 1
        return inner

    o.decode_latents, o.tiled_decode = _wrap(dec), _wrap(tdec)
    try:
        imgs, log = o.generate_image(**gen_kwargs)
    finally:
        o.decode_latents, o.tiled_decode = dec, tdec
    # the final latents are decoded last, one prompt at a time (ed:1121); verbose mode decodes its logging latents before
    prompts = gen_kwargs.get("prompts", "")
    n = 1 if isinstance(prompts, str) else len(prompts)
    return imgs, log, (torch.cat(chunks[-n:]) if chunks else None)
