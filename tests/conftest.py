import glob
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import standins  # noqa: E402

PKG = importlib.import_module("elasticdiffusion-official_b200")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


# fp32 parity runs: no TF32 in cuDNN / cuBLAS (the stand-in UNet's convs would otherwise differ from CPU fp32 by ~1e-3)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run with `-m gpu` on the B200 box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden_names():
    return sorted(n for n in (os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))
                  if not n.startswith("cn_"))


def cn_golden_names():
    """Goldens of the ControlNet twin (unmodified elastic_diffusion_w_controlnet.py)."""
    return sorted(n for n in (os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "cn_*.pt"))))


def condition_tensor(g, sd):
    """The condition image scripts/make_golden.py used: fixed-seed uniform (1,3,ds_h*8,ds_w*8)."""
    ds = PKG.geometry.low_res_size(g["kwargs"]["height"], g["kwargs"]["width"], sd, 8)
    return torch.rand(1, 3, ds[0] * 8, ds[1] * 8, generator=torch.Generator().manual_seed(g["cond_seed"]))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)


def components(sd, device="cpu"):
    """Same synthetic modules as scripts/make_golden.py (fixed-seed weights)."""
    import standins as syn
    xl = sd.startswith("XL")
    unet = syn.StubUNet(sample_size=128 if xl else 64, cross_dim=16, xl=xl, pooled_dim=8).to(device)
    vae = syn.StubVAE().to(device)
    txt = syn.StubTextEncoder(16, 8 if xl else None, device=device)
    return unet, vae, txt, (8 if xl else None)


def make_ed(sd, vb, device="cpu", controlnet=False):
    unet, vae, txt, proj = components(sd, device)
    if controlnet:
        return PKG.controlnet.ElasticDiffusion.from_components(
            device, unet, vae, None, txt, sd_version=sd, view_batch_size=vb, projection_dim=proj,
            controlnet=standins.StubControlNet().to(device))
    return PKG.ElasticDiffusion.from_components(device, unet, vae, None, txt, sd_version=sd, view_batch_size=vb,
                                                projection_dim=proj)


def oracle_models(sd, vb, device="cpu", controlnet=False):
    from oracle import reference_port as rp
    from oracle.ddim_restated import DDIMRestated
    unet, vae, txt, proj = components(sd, device)
    cn = standins.StubControlNet().to(device) if controlnet else None
    return rp.Models(unet, vae, DDIMRestated(), txt, sd, device, vb, projection_dim=proj, controlnet=cn)


def scheduler_kw(kw, target):
    """RRG weight scheduler of a golden (`rrg_scheduler` in its kwargs: "cosine" (default) / "linear" / "const", the three
    classes of ed:73-107) as the keyword the target takes: the oracle port a name, the product / the reference a class."""
    name = kw.get("rrg_scheduler", "cosine")
    if target == "port":
        return {"rrg_scheduler": name}
    return {"rrg_scherduler_cls": {"cosine": PKG.CosineScheduler, "linear": PKG.LinearScheduler, "const": PKG.ConstScheduler}[name]}


def oracle_kwargs(kw):
    """generate_image kwargs of a golden -> positional meaning for oracle.reference_port.denoise."""
    return dict(prompts=kw["prompts"], negative_prompts=kw["negative_prompts"], height=kw["height"], width=kw["width"],
                num_inference_steps=kw["num_inference_steps"], guidance_scale=kw["guidance_scale"],
                resampling_steps=kw["resampling_steps"], new_p=kw["new_p"], rrg_stop_t=kw["rrg_stop_t"],
                rrg_init_weight=kw["rrg_init_weight"], cosine_scale=kw["cosine_scale"],
                repaint_sampling=kw["repaint_sampling"])
