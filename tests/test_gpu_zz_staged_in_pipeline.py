"""GPU: the product path at a batch large enough for AUTO to choose the TMA tile-staged epilogue for the re-noise wave,
with the half kernels (exact 1/2 ratio, plan flag ED_PLAN_HALF_FAST) serving the launches without a noise stream, against
the oracle port executed eagerly on the same device.  (At the BASELINE shapes with one prompt the re-noise launch is small
and AUTO keeps the direct kernel for it; every kernel is otherwise covered one by one in test_gpu_kernels.py.)  Runs last
(file name) on purpose."""
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
    d0, s0, h0 = PKG.native.epilogue_launch_counts()
    lat, _ = ed.denoise(**kw, progress=lambda it: it)
    d1, s1, h1 = PKG.native.epilogue_launch_counts()
    # step 0: wave 1 + re-noise -> staged, wave 2 + RRG (R1 = 1) -> half; step 1 (last, no repaint, R1 = 3) -> half
    assert (s1 - s0, h1 - h0, d1 - d0) == (1, 2, 0), f"staged {s1 - s0}, half {h1 - h0}, direct {d1 - d0}"
    mse = torch.mean((lat - ref) ** 2).item()
    assert mse < 1e-8, f"mse {mse:.3e} max {(lat - ref).abs().max().item():.3e}"


def test_error_behaviour_of_the_cuda_path_matches_the_reference():
    """Sizes that are not multiples of 8 are floored like the reference does (ed:998 draws a (height // 8, width // 8)
    latent; only get_views() itself raises, ed:200-201); a condition image without a ControlNet -> ValueError."""
    ed = make_ed("2.1", 4, "cuda")
    ed.rng_device = torch.device("cpu")
    ed.seed_everything(5)
    # 516 x 1028 floors to the 64 x 128 latent and the (32, 64) low-res size of 512 x 1024 (checked against the live
    # reference in tests/test_oracle_vs_reference.py::test_sizes_not_divisible_by_8_are_floored)
    a, _ = ed.denoise("a", "b", height=516, width=1028, num_inference_steps=2, resampling_steps=1, progress=lambda it: it)
    ed.seed_everything(5)
    b, _ = ed.denoise("a", "b", height=512, width=1024, num_inference_steps=2, resampling_steps=1, progress=lambda it: it)
    assert a.shape == (1, 4, 64, 128) and torch.equal(a, b)
    with pytest.raises(TypeError):
        ed.get_views(515, 512)
    with pytest.raises(ValueError):
        ed.denoise("a", "b", height=512, width=512, num_inference_steps=1, resampling_steps=0, progress=lambda it: it,
                   condition_image=torch.rand(1, 3, 512, 512))
