"""Synthetic stand-in models (UNet / VAE / text encoder / ControlNet) for tests, smoke() and bench.py.

Not part of the product package: `diffusers` and every checkpoint are absent offline, so the dense modules that the hot
path treats as injected components (reference elastic_diffusion.py:144-153) are replaced by these for parity tests and
throughput runs.  The product (`elasticdiffusion-official_b200`) never imports this package.
"""
from .models import *  # noqa: F401,F403
from .models import StandInControlNet, StandInUNet, StandInVAE, StubControlNet, StubTextEncoder, StubUNet, StubVAE, stub_text_embeds, timestep_embedding  # noqa: F401
