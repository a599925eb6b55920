"""Synthetic model components (no `diffusers`, no weights, no network in this image).

The hot path this package accelerates treats the UNet / VAE / text encoder as injected dense modules
(reference: /root/reference/elastic_diffusion.py:144-153 loads them with `from_pretrained`).  Neither
`diffusers` nor any checkpoint exists in the build image or on the GPU box (SURVEY.md headline facts), so
parity tests, `smoke()` and `bench.py` inject the stand-ins below through
`ElasticDiffusion.from_components(...)`.  The SAME module objects are handed to the unmodified reference
(through `oracle/ref_shim.py`), to the oracle port and to the CUDA path, so every comparison is like for like.

* `StubUNet` / `StubVAE`  - tiny deterministic conv nets with the I/O contract the reference uses
  (`unet(x, t, encoder_hidden_states=..., added_cond_kwargs=...)['sample']`, `vae.encode(x).latent_dist.sample()`,
  `vae.decode(z).sample`, `.config.*`).  Weights are drawn from an explicit seeded generator so goldens
  generated in the build container reproduce on the GPU box.
* `StandInUNet`           - an SD/SDXL-*shaped* random-weight UNet (ResNet + cross-attention transformer stages at
  the real channel widths / depths) used only for throughput measurements.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Config(dict):
    """dict with attribute access (diffusers' FrozenDict behaves like this)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _seeded_(module: nn.Module, seed: int, scale: float = 1.0) -> None:
    """fan-in scaled normal init from an explicit generator (CPU generator for CPU modules: reproducible across
    machines; device generator for modules built directly on a GPU: fast, used by the throughput stand-in only)."""
    gens = {}
    with torch.no_grad():
        for p in module.parameters():
            g = gens.get(p.device)
            if g is None:
                g = gens[p.device] = torch.Generator(device=p.device).manual_seed(seed)
            fan_in = p[0].numel() if p.dim() > 1 else p.numel()
            p.copy_(torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
                    * (scale / math.sqrt(max(fan_in, 1))))


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


class StubUNet(nn.Module):
    """Two 3x3 convs modulated by (timestep, mean text embedding[, XL added conditions]).

    Non-linear (SiLU) so that cond != uncond, and with a 2-pixel receptive halo so that view context and
    background padding influence the result (the things the hot path's gather/pad glue must get right).
    """

    def __init__(self, sample_size=64, in_channels=4, cross_dim=16, hidden=16, xl=False, pooled_dim=8,
                 seed=1234):
        super().__init__()
        self.config = _Config(sample_size=sample_size, in_channels=in_channels, cross_attention_dim=cross_dim)
        self.conv_in = nn.Conv2d(in_channels, hidden, 3, padding=1)
        self.conv_out = nn.Conv2d(hidden, in_channels, 3, padding=1)
        self.time_proj = nn.Linear(16, hidden)
        self.text_proj = nn.Linear(cross_dim, hidden)
        self.xl = xl
        if xl:
            # consistent with ElasticDiffusion._get_add_time_ids (reference elastic_diffusion.py:232-246)
            self.config["addition_time_embed_dim"] = 4
            self.add_embedding = nn.Module()   # attribute path `add_embedding.linear_1.in_features` (ed:238)
            self.add_embedding.linear_1 = nn.Linear(4 * 6 + pooled_dim, hidden)
        _seeded_(self, seed, scale=1.5)

    def forward(self, x, t, encoder_hidden_states=None, added_cond_kwargs=None, **kw):
        b = x.shape[0]
        t = torch.as_tensor(t, device=x.device).reshape(-1)
        if t.numel() == 1:
            t = t.expand(b)
        emb = self.time_proj(timestep_embedding(t, 16).to(x.dtype))
        emb = emb + self.text_proj(encoder_hidden_states.to(x.dtype).mean(dim=1))
        if self.xl:
            tid = added_cond_kwargs["time_ids"].to(x.dtype)
            # crude "fourier" features of the six micro-conditioning ints: 4 per id
            feats = torch.stack([torch.sin(tid / 512.0), torch.cos(tid / 512.0),
                                 torch.sin(tid / 4096.0), torch.cos(tid / 4096.0)], dim=-1).flatten(1)
            add = torch.cat([feats, added_cond_kwargs["text_embeds"].to(x.dtype)], dim=-1)
            emb = emb + self.add_embedding.linear_1(add)
        h = self.conv_in(x) + emb[:, :, None, None]
        # ControlNet residuals (reference elastic_diffusion_w_controlnet.py:493-496): added to the hidden state
        down = kw.get("down_block_additional_residuals")
        mid = kw.get("mid_block_additional_residual")
        if down is not None:
            for r in down:
                h = h + r.to(h.dtype)
        if mid is not None:
            h = h + mid.to(h.dtype)
        h = F.silu(h)
        # denoiser-like: mostly "the noise is what you see" plus a conditioned non-linear term, so that the
        # sampled trajectory stays O(1) over 50 DDIM steps (a pure random conv makes latents grow ~1/sqrt(abar_T))
        return {"sample": 0.95 * x + 0.2 * self.conv_out(h)}


class StubControlNet(nn.Module):
    """Tiny ControlNet with the call contract the reference uses (elastic_diffusion_w_controlnet.py:482-491):
    `controlnet(x, t, encoder_hidden_states=..., controlnet_cond=..., conditioning_scale=..., guess_mode=False,
    return_dict=False[, added_cond_kwargs=...]) -> (down_block_res_samples, mid_block_res_sample)`.
    The condition image lives at pixel resolution (8x the latent); a stride-8 conv brings it to the latent grid."""

    def __init__(self, hidden=16, cross_dim=16, seed=99):
        super().__init__()
        self.cond_in = nn.Conv2d(3, hidden, 8, stride=8)
        self.x_in = nn.Conv2d(4, hidden, 3, padding=1)
        self.text_proj = nn.Linear(cross_dim, hidden)
        self.mid = nn.Conv2d(hidden, hidden, 3, padding=1)
        _seeded_(self, seed, scale=0.7)

    @property
    def dtype(self):
        return self.cond_in.weight.dtype

    def forward(self, x, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0, guess_mode=False,
                return_dict=False, added_cond_kwargs=None):
        b = x.shape[0]
        t = torch.as_tensor(t, device=x.device).reshape(-1)
        if t.numel() == 1:
            t = t.expand(b)
        emb = self.text_proj(encoder_hidden_states.to(x.dtype).mean(dim=1)) + timestep_embedding(t, 16).to(x.dtype)
        h = F.silu(self.x_in(x) + self.cond_in(controlnet_cond.to(x.dtype)) + emb[:, :, None, None])
        down = [h * conditioning_scale]
        mid = torch.tanh(self.mid(h)) * conditioning_scale
        return down, mid


class _LatentDist:
    def __init__(self, moments):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))

    def sample(self, generator=None):
        # diffusers 0.21.4 DiagonalGaussianDistribution.sample: randn on the parameters' device/dtype
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


class StubVAE(nn.Module):
    def __init__(self, scaling_factor=0.18215, seed=4321, force_upcast=False):
        super().__init__()
        self.config = _Config(block_out_channels=(8, 8, 8, 8), scaling_factor=scaling_factor,
                              force_upcast=force_upcast)
        self.enc = nn.Conv2d(3, 8, 8, stride=8)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.dec_mid = nn.Conv2d(4, 12, 3, padding=1)
        self.dec_out = nn.Conv2d(12, 3, 3, padding=1)
        _seeded_(self, seed)

    @property
    def dtype(self):
        return self.enc.weight.dtype

    @property
    def device(self):
        return self.enc.weight.device

    def encode(self, x):
        return SimpleNamespace(latent_dist=_LatentDist(self.enc(x)))

    def decode(self, z, return_dict=True):
        h = F.silu(self.dec_mid(self.post_quant_conv(z)))
        h = F.interpolate(h, scale_factor=8, mode="nearest")
        out = self.dec_out(h)
        return SimpleNamespace(sample=out) if return_dict else (out,)


def stub_text_embeds(n_prompts: int, cross_dim: int, pooled_dim: int | None, seed: int, device="cpu",
                     dtype=torch.float32):
    """Fixed-seed stand-in for `get_text_embeds` (reference elastic_diffusion.py:255-265).

    Returns (text_embeddings (n,77,cross_dim), pooled (n,pooled_dim)); for non-XL models the reference returns the
    text embeddings themselves as the "pooled" value (ed:261-262), which the non-XL UNet path never reads.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    emb = torch.randn(n_prompts, 77, cross_dim, generator=g)
    if pooled_dim is None:
        return emb.to(device, dtype), emb.to(device, dtype)
    pooled = torch.randn(n_prompts, pooled_dim, generator=g)
    return emb.to(device, dtype), pooled.to(device, dtype)


class StubTextEncoder:
    """Deterministic prompt -> embedding map: seed = hash of the prompt strings."""

    def __init__(self, cross_dim=16, pooled_dim=None, device="cpu", dtype=torch.float32):
        self.cross_dim, self.pooled_dim, self.device, self.dtype = cross_dim, pooled_dim, device, dtype

    def __call__(self, prompts):
        import hashlib
        if isinstance(prompts, str):
            prompts = [prompts]
        outs = [stub_text_embeds(1, self.cross_dim, self.pooled_dim,
                                 int(hashlib.md5(p.encode()).hexdigest()[:8], 16), self.device, self.dtype)
                for p in prompts]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])


# --------------------------------------------------------------------------------------------------------------
# SD / SDXL-shaped random-weight UNet for throughput runs
# --------------------------------------------------------------------------------------------------------------

class _PlainOps:
    """The element-wise / normalisation patterns of the UNet as plain PyTorch (what diffusers' modules run).  A UNet instance
    can be given another implementation of this interface through `StandInUNet.set_ops` - e.g. the product's fused kernels
    (`elasticdiffusion-official_b200/unet_ops.py`); the stand-ins themselves never import the product."""

    @staticmethod
    def geglu(x):
        a, g = x.chunk(2, dim=-1)
        return a * F.gelu(g)

    @staticmethod
    def group_norm(gn, x):
        return gn(x)

    @staticmethod
    def group_norm_silu(gn, x):
        return F.silu(gn(x))

    @staticmethod
    def layer_norm(ln, x):
        return ln(x)

    @staticmethod
    def conv_add(conv, x, per_nc=None, residual=None):
        """conv(x) [+ per_nc[:, :, None, None]] [+ residual] - a ResNet block's conv with its time-embedding / skip adds"""
        h = conv(x)
        if per_nc is not None:
            h = h + per_nc[:, :, None, None]
        if residual is not None:
            h = h + residual
        return h


class _ResBlock(nn.Module):
    ops = _PlainOps

    def __init__(self, cin, cout, temb):
        super().__init__()
        self.n1 = nn.GroupNorm(32, cin)
        self.c1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.t = nn.Linear(temb, cout)
        self.n2 = nn.GroupNorm(32, cout)
        self.c2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        o = self.ops
        h = o.conv_add(self.c1, o.group_norm_silu(self.n1, x), per_nc=self.t(F.silu(emb)))
        return o.conv_add(self.c2, o.group_norm_silu(self.n2, h), residual=x if self.skip is None else o.conv_add(self.skip, x))


class _Attn(nn.Module):
    def __init__(self, dim, ctx_dim, head_dim):
        super().__init__()
        self.h = dim // head_dim
        self.q = nn.Linear(dim, dim, bias=False)
        self.k = nn.Linear(ctx_dim, dim, bias=False)
        self.v = nn.Linear(ctx_dim, dim, bias=False)
        self.o = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        b, n, d = x.shape
        q = self.q(x).view(b, n, self.h, -1).transpose(1, 2)
        k = self.k(ctx).view(b, ctx.shape[1], self.h, -1).transpose(1, 2)
        v = self.v(ctx).view(b, ctx.shape[1], self.h, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.o(o.transpose(1, 2).reshape(b, n, d))


class _TBlock(nn.Module):
    ops = _PlainOps

    def __init__(self, dim, ctx_dim, head_dim):
        super().__init__()
        self.n1, self.n2, self.n3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.a1 = _Attn(dim, dim, head_dim)
        self.a2 = _Attn(dim, ctx_dim, head_dim)
        self.ff1 = nn.Linear(dim, dim * 8)   # GEGLU: value + gate
        self.ff2 = nn.Linear(dim * 4, dim)

    def forward(self, x, ctx):
        o = self.ops
        y = o.layer_norm(self.n1, x)
        x = x + self.a1(y, y)
        x = x + self.a2(o.layer_norm(self.n2, x), ctx)
        return x + self.ff2(o.geglu(self.ff1(o.layer_norm(self.n3, x))))


class _Transformer2D(nn.Module):
    ops = _PlainOps

    def __init__(self, dim, ctx_dim, head_dim, depth):
        super().__init__()
        self.norm = nn.GroupNorm(32, dim)
        self.pin = nn.Linear(dim, dim)
        self.blocks = nn.ModuleList([_TBlock(dim, ctx_dim, head_dim) for _ in range(depth)])
        self.pout = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        y = self.pin(self.ops.group_norm(self.norm, x).flatten(2).transpose(1, 2))
        for blk in self.blocks:
            y = blk(y, ctx)
        y = self.pout(y).transpose(1, 2).reshape(b, c, h, w)
        return x + y


class StandInUNet(nn.Module):
    """UNet with the stage widths / transformer depths of SDXL-base (or SD2.1-base), random weights.

    SDXL-base (public config): block_out_channels (320, 640, 1280), 2 ResNet layers per block,
    transformer depth (0, 2, 10) on the down path, 10 in the mid block, head dim 64, cross-attention dim 2048,
    added (text_embeds 1280 + 6 x 256 time ids) -> 2816 -> 1280 embedding; sample_size 128.
    SD2.1-base: (320, 640, 1280, 1280), depth 1 on the first three stages, head dim 64, cross dim 1024, sample 64.
    The I/O contract is the one the reference uses (elastic_diffusion.py:422,426).
    """

    PRESETS = {
        "XL1.0": dict(widths=(320, 640, 1280), depths=(0, 2, 10), mid_depth=10, ctx=2048, sample_size=128,
                      add_in=2816, add_time_dim=256),
        "2.1": dict(widths=(320, 640, 1280, 1280), depths=(1, 1, 1, 0), mid_depth=1, ctx=1024, sample_size=64,
                    add_in=None, add_time_dim=None),
        "1.5": dict(widths=(320, 640, 1280, 1280), depths=(1, 1, 1, 0), mid_depth=1, ctx=768, sample_size=64,
                    add_in=None, add_time_dim=None),
        # a narrow variant for quick functional runs of the same topology
        "tiny-xl": dict(widths=(64, 128, 256), depths=(0, 1, 2), mid_depth=2, ctx=64, sample_size=128,
                        add_in=6 * 8 + 32, add_time_dim=8),
    }

    def __init__(self, preset="XL1.0", layers_per_block=2, head_dim=64, seed=7, device=None, dtype=None):
        super().__init__()
        if device is not None or dtype is not None:   # build + initialise directly on the target device / dtype
            old = torch.get_default_dtype()
            try:
                torch.set_default_dtype(dtype or old)
                with torch.device(device or "cpu"):
                    self._build(preset, layers_per_block, head_dim, seed)
            finally:
                torch.set_default_dtype(old)
        else:
            self._build(preset, layers_per_block, head_dim, seed)

    ENCODER_ONLY = False    # StandInControlNet: conv_in + down path + mid block only
    ops = _PlainOps

    def set_ops(self, ops):
        """Route GEGLU / GroupNorm(+SiLU) of every block through `ops` (interface of `_PlainOps`)."""
        self.ops = ops
        for m in self.modules():
            if isinstance(m, (_ResBlock, _TBlock, _Transformer2D)):
                m.ops = ops
        return self

    def _build(self, preset, layers_per_block, head_dim, seed):
        p = self.PRESETS[preset]
        widths, depths, ctx = p["widths"], p["depths"], p["ctx"]
        self.config = _Config(sample_size=p["sample_size"], in_channels=4, cross_attention_dim=ctx)
        temb = widths[0] * 4
        self.temb_dim0 = widths[0]
        self.time1, self.time2 = nn.Linear(widths[0], temb), nn.Linear(temb, temb)
        self.add_time_dim = p["add_time_dim"]
        if p["add_in"] is not None:
            self.config["addition_time_embed_dim"] = p["add_time_dim"]
            self.add_embedding = nn.Module()
            self.add_embedding.linear_1 = nn.Linear(p["add_in"], temb)
            self.add_embedding.linear_2 = nn.Linear(temb, temb)
        if min(widths) < 64:
            head_dim = 32
        self.conv_in = nn.Conv2d(4, widths[0], 3, padding=1)
        self.down = nn.ModuleList()
        chans = [widths[0]]
        c = widths[0]
        for i, (w, d) in enumerate(zip(widths, depths)):
            for _ in range(layers_per_block):
                self.down.append(nn.ModuleList([_ResBlock(c, w, temb),
                                                _Transformer2D(w, ctx, head_dim, d) if d else nn.Identity()]))
                c = w
                chans.append(c)
            if i < len(widths) - 1:
                self.down.append(nn.ModuleList([nn.Conv2d(c, c, 3, stride=2, padding=1), None]))
                chans.append(c)
        self.mid1 = _ResBlock(c, c, temb)
        self.mid_attn = _Transformer2D(c, ctx, head_dim, p["mid_depth"])
        self.mid2 = _ResBlock(c, c, temb)
        if self.ENCODER_ONLY:
            self._build_control(chans, c, widths[0])
            _seeded_(self, seed)
            return
        self.up = nn.ModuleList()
        for i, (w, d) in reversed(list(enumerate(zip(widths, depths)))):
            for j in range(layers_per_block + 1):
                skip = chans.pop()
                self.up.append(nn.ModuleList([_ResBlock(c + skip, w, temb),
                                              _Transformer2D(w, ctx, head_dim, d) if d else nn.Identity(),
                                              nn.Conv2d(w, w, 3, padding=1)
                                              if (j == layers_per_block and i > 0) else None]))
                c = w
        self.norm_out = nn.GroupNorm(32, c)
        self.conv_out = nn.Conv2d(c, 4, 3, padding=1)
        _seeded_(self, seed)

    def _embed(self, x, t, encoder_hidden_states, added_cond_kwargs):
        b = x.shape[0]
        t = torch.as_tensor(t, device=x.device).reshape(-1)
        if t.numel() == 1:
            t = t.expand(b)
        dt = self.conv_in.weight.dtype
        emb = self.time2(F.silu(self.time1(timestep_embedding(t, self.temb_dim0).to(dt))))
        if hasattr(self, "add_embedding"):
            tid = added_cond_kwargs["time_ids"].flatten()
            te = timestep_embedding(tid, self.add_time_dim).reshape(b, -1).to(dt)
            add = torch.cat([added_cond_kwargs["text_embeds"].to(dt), te], dim=-1)
            emb = emb + self.add_embedding.linear_2(F.silu(self.add_embedding.linear_1(add)))
        return emb, encoder_hidden_states.to(dt), dt

    def _encode(self, h, emb, ctx):
        """conv_in output -> (skip list, mid-block output)"""
        skips = [h]
        for mods in self.down:
            if mods[1] is None:
                h = self.ops.conv_add(mods[0], h)
            else:
                h = mods[0](h, emb)
                h = h if isinstance(mods[1], nn.Identity) else mods[1](h, ctx)
            skips.append(h)
        h = self.mid2(self.mid_attn(self.mid1(h, emb), ctx), emb)
        return skips, h

    def forward(self, x, t, encoder_hidden_states=None, added_cond_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, **kw):
        emb, ctx, dt = self._embed(x, t, encoder_hidden_states, added_cond_kwargs)
        skips, h = self._encode(self.ops.conv_add(self.conv_in, x.to(dt)), emb, ctx)
        # ControlNet residuals as diffusers' UNet2DConditionModel applies them (reference elastic_diffusion_w_controlnet.py
        # :493-496 passes them through): one per skip connection + one for the mid block
        if down_block_additional_residuals is not None:
            skips = [s + r.to(s.dtype) for s, r in zip(skips, down_block_additional_residuals)]
        if mid_block_additional_residual is not None:
            h = h + mid_block_additional_residual.to(h.dtype)
        for res, attn, upc in self.up:
            h = res(torch.cat([h, skips.pop()], dim=1), emb)
            h = h if isinstance(attn, nn.Identity) else attn(h, ctx)
            if upc is not None:
                h = self.ops.conv_add(upc, F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return {"sample": self.ops.conv_add(self.conv_out, self.ops.group_norm_silu(self.norm_out, h))}


class StandInControlNet(StandInUNet):
    """SD/SDXL-*shaped* random-weight ControlNet for throughput runs (cfg5): the UNet's encoder half (conv_in, down path,
    mid block at the real widths / transformer depths) + the condition embedding (3 -> 16 -> 32 -> 96 -> 256 -> width0
    3x3 convs, three of them stride 2: pixel resolution -> latent grid, like diffusers' ControlNetConditioningEmbedding)
    + one 1x1 "zero conv" per skip connection and one for the mid block.  Call contract of the reference
    (elastic_diffusion_w_controlnet.py:482-491): returns (down_block_res_samples, mid_block_res_sample)."""

    ENCODER_ONLY = True

    def _build_control(self, chans, c_mid, w0):
        ce = [3, 16, 32, 96, 256]
        self.cond_in = nn.Conv2d(3, 16, 3, padding=1)
        self.cond_blocks = nn.ModuleList()
        for a, b in zip(ce[1:-1], ce[2:]):
            self.cond_blocks.append(nn.Conv2d(a, a, 3, padding=1))
            self.cond_blocks.append(nn.Conv2d(a, b, 3, padding=1, stride=2))
        self.cond_out = nn.Conv2d(256, w0, 3, padding=1)
        self.zero_convs = nn.ModuleList([nn.Conv2d(c, c, 1) for c in chans])
        self.mid_zero = nn.Conv2d(c_mid, c_mid, 1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, x, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0, guess_mode=False,
                return_dict=False, added_cond_kwargs=None):
        emb, ctx, dt = self._embed(x, t, encoder_hidden_states, added_cond_kwargs)
        c = F.silu(self.cond_in(controlnet_cond.to(dt)))
        for blk in self.cond_blocks:
            c = F.silu(blk(c))
        skips, h = self._encode(self.conv_in(x.to(dt)) + self.cond_out(c), emb, ctx)
        down = [z(s) * conditioning_scale for z, s in zip(self.zero_convs, skips)]
        return down, self.mid_zero(h) * conditioning_scale


# --------------------------------------------------------------------------------------------------------------
# SD/SDXL-VAE-shaped decoder for decode-time measurements (tiled decode, cfg4) and the decode de-duplication study
# --------------------------------------------------------------------------------------------------------------
class _VaeRes(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(32, cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.n2, self.c2 = nn.GroupNorm(32, cout), nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.c1(F.silu(self.n1(x)))
        h = self.c2(F.silu(self.n2(h)))
        return h + (x if self.skip is None else self.skip(x))


class _VaeAttn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = nn.GroupNorm(32, c)
        self.q, self.k, self.v, self.o = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.norm(x).flatten(2).transpose(1, 2)
        o = F.scaled_dot_product_attention(self.q(y)[:, None], self.k(y)[:, None], self.v(y)[:, None])[:, 0]
        return x + self.o(o).transpose(1, 2).reshape(b, c, h, w)


class StandInVAE(StubVAE):
    """StubVAE whose `decode` is an SD/SDXL-VAE-*shaped* random-weight decoder: block_out_channels (128, 256, 512, 512),
    3 ResNet layers per up block, single-head mid attention, GroupNorm(32) + SiLU, three nearest-2x upsamples (8x total) -
    the public AutoencoderKL decoder topology.  `encode` stays the stub's (background strips only).  GroupNorm statistics
    are per decoded tile, which is what makes tiled decoding differ from one whole-image decode."""

    def __init__(self, widths=(128, 256, 512, 512), layers=3, seed=4321, scaling_factor=0.13025, force_upcast=True,
                 device=None, dtype=None):
        super().__init__(scaling_factor=scaling_factor, seed=seed, force_upcast=force_upcast)
        self.config["block_out_channels"] = tuple(widths)
        with torch.device(device or "cpu"):
            c = widths[-1]
            self.d_in = nn.Conv2d(4, c, 3, padding=1)
            self.d_mid = nn.ModuleList([_VaeRes(c, c), _VaeAttn(c), _VaeRes(c, c)])
            self.d_up = nn.ModuleList()
            for i, w in enumerate(reversed(widths)):
                blk = nn.ModuleList([_VaeRes(c if j == 0 else w, w) for j in range(layers)])
                c = w
                self.d_up.append(nn.ModuleList([blk, nn.Conv2d(w, w, 3, padding=1) if i < len(widths) - 1 else None]))
            self.d_norm, self.d_out = nn.GroupNorm(32, c), nn.Conv2d(c, 3, 3, padding=1)
        if device is not None:
            self.to(device)
        if dtype is not None:
            self.to(dtype)
        _seeded_(self, seed)

    def decode(self, z, return_dict=True):
        h = self.d_in(self.post_quant_conv(z))
        for m in self.d_mid:
            h = m(h)
        for blk, up in self.d_up:
            for r in blk:
                h = r(h)
            if up is not None:
                h = up(F.interpolate(h, scale_factor=2.0, mode="nearest"))
        out = self.d_out(F.silu(self.d_norm(h)))
        return SimpleNamespace(sample=out) if return_dict else (out,)
