"""CPU: the oracle (oracle/reference_port.py) reproduces the committed goldens, which were produced by the unmodified
reference (scripts/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

from conftest import scheduler_kw, cn_golden_names, condition_tensor, golden_names, load_golden, oracle_kwargs, oracle_models
from oracle import reference_port as rp

FAST = [n for n in golden_names() if "2048x2048" not in n]


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_latent(name):
    g = load_golden(name)
    m = oracle_models(g["sd_version"], g["view_batch_size"])
    rp.seed_all(g["seed"], "cpu")
    lat = rp.denoise(m, **oracle_kwargs(g["kwargs"]), **scheduler_kw(g["kwargs"], "port"))
    assert lat.shape == g["latent"].shape
    assert torch.equal(lat, g["latent"]), f"max abs diff {(lat - g['latent']).abs().max().item():.3e}"


def test_oracle_tiled_decode_matches_reference_image():
    g = load_golden("xl_2048x2048_T2_R2_tiled")
    m = oracle_models(g["sd_version"], g["view_batch_size"])
    img = rp.decode_tiled(m, g["latent"])
    stats = torch.nn.functional.adaptive_avg_pool2d((img * 255).byte().float() / 255.0, 16)
    assert torch.allclose(stats, g["image_stats"], atol=1e-6)


def test_oracle_plain_decode_matches_reference_image():
    g = load_golden("sd21_512x1024_T4_R4")
    m = oracle_models(g["sd_version"], g["view_batch_size"])
    img = rp.decode_plain(m, g["latent"])
    stats = torch.nn.functional.adaptive_avg_pool2d((img * 255).byte().float() / 255.0, 16)
    assert torch.allclose(stats, g["image_stats"], atol=1e-6)


def test_ddim_restated_known_values():
    """Known-answer checks of the DDIM restatement (closed forms of the scaled-linear schedule)."""
    from oracle.ddim_restated import DDIMRestated
    s = DDIMRestated()
    assert abs(float(s.betas[0]) - 0.00085) < 1e-9 and abs(float(s.betas[-1]) - 0.012) < 1e-8
    s.set_timesteps(50)
    assert s.timesteps[:3].tolist() == [981, 961, 941] and s.timesteps[-1].item() == 1
    s.set_timesteps(10)
    assert s.timesteps.tolist() == [901, 801, 701, 601, 501, 401, 301, 201, 101, 1]
    # step() inverts add_noise when the model predicts the true noise
    x0, eps = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    t = s.timesteps[3]
    xt = s.add_noise(x0, eps, t)
    out = s.step(eps, t, xt)
    assert torch.allclose(out["pred_original_sample"], x0, atol=1e-4)
    prev = s.add_noise(x0, eps, t - 100)
    assert torch.allclose(out["prev_sample"], prev, atol=1e-4)


@pytest.mark.parametrize("name", cn_golden_names())
def test_oracle_reproduces_controlnet_twin_latent(name):
    """goldens from the unmodified elastic_diffusion_w_controlnet.py (scripts/make_golden.py --cn-only)"""
    g = load_golden(name)
    m = oracle_models(g["sd_version"], g["view_batch_size"], controlnet=True)
    rp.seed_all(g["seed"], "cpu")
    lat = rp.denoise(m, **oracle_kwargs(g["kwargs"]), condition_image=condition_tensor(g, g["sd_version"]),
                     controlnet_conditioning_scale=g["kwargs"]["controlnet_conditioning_scale"])
    assert torch.equal(lat, g["latent"]), f"max abs diff {(lat - g['latent']).abs().max().item():.3e}"
