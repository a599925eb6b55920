"""B200-native ElasticDiffusion hot path (global/local patched denoising loop) behind the reference's class API.

The directory name carries the reference's repository name (`elasticdiffusion-official_b200`); since it is not a
valid Python identifier, import it with `importlib.import_module("elasticdiffusion-official_b200")` or through the
top-level drop-in module `elastic_diffusion` (same module name as the reference's file).
"""
from .pipeline import (ConstScheduler, CosineScheduler, ElasticDiffusion, LinearScheduler, RngLedger, TimeIt,
                       timelog)
from .ddim import DDIMSchedule
from . import controlnet, geometry, native, unet_ops

__all__ = ["ElasticDiffusion", "CosineScheduler", "LinearScheduler", "ConstScheduler", "TimeIt", "timelog",
           "DDIMSchedule", "RngLedger", "controlnet", "geometry", "native", "unet_ops"]
