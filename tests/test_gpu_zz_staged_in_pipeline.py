"""GPU: the product path at a batch large enough for AUTO to choose the TMA tile-staged epilogue (both of its
instantiations: wave 1 with the re-noise stream, wave 2 without), against the oracle port executed eagerly on the same
device.  (At the BASELINE shapes with one prompt every launch is small and AUTO keeps the direct kernel; the staged kernel
is otherwise covered kernel-by-kernel in test_gpu_kernels.py.)  Runs last (file name) on purpose."""
import pytest
import torch

from conftest import PKG, make_ed, oracle_models
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


def test_large_batch_takes_the_staged_epilogue_and_matches_the_oracle():
    B = 40                                     # SD2.1 512x1024: (128 / 4) * 64 * 40 threads' worth of pixels >= 2 * 148 * 256
    kw = dict(prompts=[f"prompt {i}" for i in range(B)], negative_prompts="blurry", height=512, width=1024,
              num_inference_steps=2, resampling_steps=2, cosine_scale=10.0)
    m = oracle_models("2.1", 8, "cuda")
    inner = m.unet

    class NoAutocast(torch.nn.Module):         # fp32 comparison: the oracle's loop enables autocast like the reference
        config = inner.config

        def __getattr__(self, k):
            return getattr(inner, k)

        def forward(self, *a, **k):
            with torch.autocast("cuda", enabled=False):
                return inner(*a, **k)

    m.unet = NoAutocast()
    rp.seed_all(3, "cuda")
    ref = rp.denoise(m, **kw)
    ed = make_ed("2.1", 8, "cuda")
    ed.autocast = False
    ed.seed_everything(3)
    d0, s0 = PKG.native.epilogue_launch_counts()
    lat, _ = ed.denoise(**kw, progress=lambda it: it)
    d1, s1 = PKG.native.epilogue_launch_counts()
    assert s1 - s0 == 3 and d1 == d0, f"expected 3 staged epilogue launches (2 waves + 1), got staged {s1 - s0}, direct {d1 - d0}"
    mse = torch.mean((lat - ref) ** 2).item()
    assert mse < 1e-8, f"mse {mse:.3e} max {(lat - ref).abs().max().item():.3e}"


def test_error_behaviour_of_the_cuda_path_matches_the_reference():
    """ed:200-201: sizes that are not multiples of 8 -> TypeError (the reference raises a str), before any kernel runs;
    a condition image without a ControlNet -> ValueError."""
    ed = make_ed("2.1", 4, "cuda")
    with pytest.raises(TypeError):
        ed.generate_image("a", "b", height=515, width=512, num_inference_steps=1, resampling_steps=0, progress=lambda it: it)
    with pytest.raises(ValueError):
        ed.denoise("a", "b", height=512, width=512, num_inference_steps=1, resampling_steps=0, progress=lambda it: it,
                   condition_image=torch.rand(1, 3, 512, 512))
    assert ed.last_run["kernel_launches"] == 0
