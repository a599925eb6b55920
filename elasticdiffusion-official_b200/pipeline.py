"""`ElasticDiffusion` - drop-in for the reference class of the same name, hot path on hand-written sm_100a kernels.

Boundary: constructor and `generate_image` signatures are identical to
/root/reference/elastic_diffusion.py:111-115 and :953-965 ("ed:N" below); callers also use `seed_everything`,
`set_view_config`, `get_downsample_size`, `view_batch_size`, `view_config`, `vae_scale_factor` and the module-level
`timelog`, `CosineScheduler`, `LinearScheduler`, `ConstScheduler` - all kept.

Design: one denoise step = two *waves*.  All UNet samples of a wave - the 2(R+1) global resampling passes
and the nv local views - depend only on the wave's input latent and on RNG draws that never depend on UNet outputs,
so the host replays the reference's exact RNG ledger first (`RngLedger`), one gather kernel pair builds the whole UNet
batch, PyTorch runs the UNet once (or sharded over ranks), and one fused epilogue kernel turns the outputs into the
next latent (view scatter + direction fill + CFG + DDIM [+ undo_step] [+ RRG]).

There is no CPU / eager fallback: without the CUDA library or on a non-CUDA device `generate_image` raises.
"""
from __future__ import annotations

import ctypes
import hashlib
import time
from contextlib import contextmanager
from typing import Any

import numpy as np
import torch
import torch.nn as nn

from . import native
from .ddim import DDIMSchedule, check_scheduler, renoise_scalars, step_scalars
from .geometry import build_geometry, build_tiles, cond_geometry, low_res_size, pad_split

try:  # progress bar default of the reference signature (ed:963)
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, *a, **k):
        return it


# ---------------------------------------------------------------------------------------------------------------
# module-level API kept from the reference (ed:33-109)
# ---------------------------------------------------------------------------------------------------------------
class TimeIt:
    """Wall-clock accumulator with the reference's interface (ed:33-70)."""

    def __init__(self, sync_gpu=False):
        self.sync_gpu = sync_gpu
        self.total_time = {}

    def _tick(self):
        if self.sync_gpu and torch.cuda.is_available():
            torch.cuda.synchronize()
        return time.time()

    def time_function(self, func):
        def wrapper(*args, **kwargs):
            t0 = self._tick()
            out = func(*args, **kwargs)
            key = f"FUNCTION_{func.__name__}"
            self.total_time[key] = self.total_time.get(key, 0) + (self._tick() - t0)
            return out
        return wrapper

    @contextmanager
    def time_block(self, block_title):
        t0 = self._tick()
        try:
            yield
        finally:
            key = f"BLOCK_{block_title}"
            self.total_time[key] = self.total_time.get(key, 0) + (self._tick() - t0)

    def print_results(self):
        for key, spent in self.total_time.items():
            print(f"{key} took total {spent} seconds to complete.")


class LinearScheduler:
    """RRG weight schedule, ed:73-82."""

    def __init__(self, steps, start_val, stop_val):
        self.steps, self.start_val, self.stop_val = steps, start_val, stop_val

    def __call__(self, t, *args: Any, **kwds: Any) -> Any:
        if t >= self.steps:
            return self.stop_val
        return self.start_val + (self.stop_val - self.start_val) / self.steps * t


class ConstScheduler(LinearScheduler):
    """ed:85-94."""

    def __call__(self, t, *args: Any, **kwds: Any) -> Any:
        return self.stop_val if t >= self.steps else self.start_val


class CosineScheduler:
    """ed:96-107."""

    def __init__(self, steps, cosine_scale, factor=0.01):
        self.steps, self.cosine_scale, self.factor = steps, cosine_scale, factor

    def __call__(self, t, *args: Any, **kwds: Any) -> Any:
        if t >= self.steps:
            return 0
        return self.factor * ((0.5 * (1 + np.cos(np.pi * t / self.steps))) ** self.cosine_scale)


timelog = TimeIt(sync_gpu=False)

MODEL_KEYS = {"2.1": "stabilityai/stable-diffusion-2-1-base", "2.0": "stabilityai/stable-diffusion-2-base",
              "1.5": "runwayml/stable-diffusion-v1-5", "1.4": "CompVis/stable-diffusion-v1-4",
              "XL1.0": "stabilityai/stable-diffusion-xl-base-1.0"}


def shard_range(n, world, rank):
    """(per, lo, hi): contiguous block partition of n units over `world` ranks - rank r owns [lo, hi) = [r*per, (r+1)*per)
    clipped to n, per = ceil(n / world).  Used for the UNet samples of a wave and for the decode tiles; the peer kernels
    resolve unit u to rank u // per."""
    per = (n + world - 1) // world if world > 1 else n
    return per, min(rank * per, n), min((rank + 1) * per, n)


def _i32(dev, values):
    return torch.tensor(list(values) if len(values) else [0], dtype=torch.int32, device=dev)


# ---------------------------------------------------------------------------------------------------------------
# RNG ledger: the reference's ordered sequence of generator operations, replayed on the host
# before any kernel of the wave runs.  Only the *order and shape* of draws matter for parity, not where the UNet runs.
# ---------------------------------------------------------------------------------------------------------------
class RngLedger:
    def __init__(self, owner: "ElasticDiffusion", geo):
        self.o, self.geo = owner, geo
        self.dev = owner.device
        self.rng_dev = owner.rng_device if owner.rng_device is not None else owner.device
        self.strip_cache = {}
        self.host_ms = {"cells": 0.0, "drop": 0.0, "strips": 0.0, "undo": 0.0}   # host wall time per planner component
        self.n_cells = geo.lh * geo.lw
        self._rows = np.arange(self.n_cells)
        # device draws whose VALUES the host needs (drop masks, ed:541) run on a side stream so that reading them back
        # never waits for the UNet work queued on the main stream (Philox offsets are assigned at call time on the
        # host, so the stream a draw runs on does not change its values)
        # HIGH priority: measured on B200 (scripts/side_stream_probe.py), a launch on a default-priority stream blocks the
        # host until the CUDA graph running on the main stream has finished; a high-priority stream issues in ~0.2 ms
        self._side = torch.cuda.Stream(device=self.dev, priority=-1) if self.rng_dev.type == "cuda" else None
        self._side_ev = torch.cuda.Event() if self._side is not None else None
        self._drop_pinned = torch.empty(self.n_cells, dtype=torch.int64).pin_memory() if self._side is not None else None

    def _drop_draw(self):
        """torch.randint(0, 101, (N,), device=...) of ed:541, returned on the host."""
        n = self.n_cells
        if self._side is None:
            return torch.randint(0, 101, (n,), device=self.rng_dev).cpu()
        with torch.cuda.stream(self._side):
            d = torch.randint(0, 101, (n,), device=self.rng_dev)
            self._drop_pinned.copy_(d, non_blocking=True)
            self._side_ev.record(self._side)
        self._side_ev.synchronize()
        return self._drop_pinned.clone()

    # -- seeds (ed:165-171, 321-324) -----------------------------------------------------------------------------
    def _seed(self, seed):
        """seed_everything(seed, seed_np=False) of ed:165-168 for the generators this process draws from: the CPU default
        generator and this device's CUDA generator.  (torch.manual_seed also walks every other visible device and
        backend hook, ~0.7 ms per call x 36 calls per step; seeding the two generators directly is equivalent here.)"""
        torch.default_generator.manual_seed(seed)
        if self.dev.type == "cuda":
            idx = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
            torch.cuda.default_generators[idx].manual_seed(seed)

    @staticmethod
    def _md5_seed(s):
        return int(hashlib.md5(s.encode()).hexdigest()[:8], 16)

    def _randn(self, shape, dtype=torch.float32):
        return torch.randn(shape, device=self.rng_dev, dtype=dtype).to(self.dev)

    def initial_latent(self, shape, dtype):
        return self._randn(shape, dtype).float()                                   # ed:998

    # -- background strips (ed:327-364); pure function of the id string, cached (the reference's TODO ed:340) ------
    def _strip(self, h, w, t, tag):
        if h == 0 or w == 0:
            return None                                                            # ed:332-333: no generator touched
        key = f"{tag}_{h}_{w}_{t}"                                                 # same id string as ed:331
        hit = self.strip_cache.get(key)
        if hit is None:
            o = self.o
            with torch.autocast("cuda", enabled=False):
                self._seed(self._md5_seed(key))
                colour = torch.rand(1, 3, device=self.rng_dev).to(self.dev)
                img = colour[:, :, None, None].repeat(1, 1, h * o.vae_scale_factor, w * o.vae_scale_factor)
                upcast = o.vae.dtype == torch.float16 and o.vae.config.force_upcast
                if o.low_vram:
                    o.vae.to(self.dev)
                if upcast:
                    o.upcast_vae()
                    img = img.float()
                dist = o.vae.encode(img.to(o.vae.dtype) if not upcast else img).latent_dist
                o.last_run["vae_encode_calls"] = o.last_run.get("vae_encode_calls", 0) + 1
                if self.rng_dev == self.dev:
                    z = dist.sample()
                else:   # test mode (CPU generator): same formula as DiagonalGaussianDistribution.sample
                    z = dist.mean + dist.std * torch.randn(dist.mean.shape, device=self.rng_dev,
                                                           dtype=dist.mean.dtype).to(self.dev)
                z = z * o.vae.config.scaling_factor
                noise = self._randn(z.shape, z.dtype)
                # scheduler.add_noise (ed:358) = sqrt(abar_t) z + sqrt(1-abar_t) noise.  Its .to(device) of the schedule
                # table is a synchronous pageable H2D copy, i.e. a full GPU sync per strip that would serialise the
                # look-ahead planning with the UNet; the two fp32 scalars are taken on the host instead (same values).
                ac = o.scheduler.alphas_cumprod[int(t)]
                hit = (float(ac ** 0.5) * z + float((1 - ac) ** 0.5) * noise).float().contiguous()
                if upcast:
                    o.vae.to(dtype=torch.float16)
            self.strip_cache[key] = hit
        # every non-empty use consumes one numpy draw and re-keys torch, hit or miss (ed:359)
        self._seed(int(np.random.randint(100000)))
        return hit

    def _strip_keys(self, inner_h, inner_w, t):
        """(key, h, w) of the non-empty strips of one padded unet_step, in the order pad_events touches them."""
        nat = self.geo.native
        l, r = pad_split(nat, inner_w)
        tp, b = pad_split(nat, inner_h)
        wide = inner_w + l + r
        cand = [("3_1", inner_h, l), ("3_2", inner_h, r), ("2_1", tp, wide), ("2_2", b, wide)]
        return [(f"{tag}_{h}_{w}_{t}", h, w) for tag, h, w in cand if h > 0 and w > 0]

    def precompute_strips(self, ts, chunk=32):
        """SURVEY 8 row f2 (the reference's TODO ed:340): every background strip of the run - a pure function of its id
        string (dim, side, h, w, t) - built BEFORE the loop with batched VAE encodes instead of one fp32 encode per cache
        miss inside the look-ahead planner.  Per strip the generator operations are exactly those of ed:335-358 (re-key
        from the md5 of the id, rand(1,3) colour, the latent distribution's randn, add_noise's randn) in the same order on
        a re-keyed generator; only the VAE encode between the colour draw and the sample draw - which consumes no random
        numbers - is deferred and batched.  The global generators are restored afterwards, so the loop's own draws
        (ed:502-544, 701) and the per-use numpy re-seeds (ed:359) are untouched: RNG-neutral.  Returns the number of
        batched VAE encode calls."""
        o, g = self.o, self.geo
        todo, seen = [], set()
        for t in ts:
            shapes = [(g.lh, g.lw)] + ([(g.vh, g.vw)] if (g.vh < g.native or g.vw < g.native) else [])
            for ih, iw in shapes:
                for key, h, w in self._strip_keys(ih, iw, t):
                    if key not in seen and key not in self.strip_cache:
                        seen.add(key)
                        todo.append((key, h, w, t))
        if not todo:
            return 0
        cpu_state = torch.default_generator.get_state()
        cuda_gen = None
        if self.dev.type == "cuda":
            idx = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
            cuda_gen = torch.cuda.default_generators[idx]
            cuda_state = cuda_gen.get_state()
        calls = 0
        try:
            with torch.autocast("cuda", enabled=False):
                upcast = o.vae.dtype == torch.float16 and o.vae.config.force_upcast
                if o.low_vram:
                    o.vae.to(self.dev)
                if upcast:
                    o.upcast_vae()
                vdt = torch.float32 if upcast else o.vae.dtype
                C = g.C
                draws = {}
                for key, h, w, t in todo:                                   # generator operations, in the reference's order
                    self._seed(self._md5_seed(key))                         # ed:335
                    colour = torch.rand(1, 3, device=self.rng_dev).to(self.dev)              # ed:336
                    eps_sample = self._randn((1, C, h, w), vdt)             # latent_dist.sample() (ed:350)
                    eps_noise = self._randn((1, C, h, w), vdt)              # randn_like (ed:356)
                    draws[key] = (colour, eps_sample, eps_noise)
                by_shape = {}
                for item in todo:
                    by_shape.setdefault((item[1], item[2]), []).append(item)
                sf = o.vae_scale_factor
                for (h, w), items in by_shape.items():
                    for s in range(0, len(items), chunk):
                        part = items[s:s + chunk]
                        imgs = torch.cat([draws[k][0] for k, *_ in part])[:, :, None, None].expand(-1, -1, h * sf, w * sf)
                        dist = o.vae.encode(imgs.to(vdt).contiguous()).latent_dist              # ONE encode for the chunk
                        calls += 1
                        for j, (key, _, _, t) in enumerate(part):
                            _, e1, e2 = draws[key]
                            z = (dist.mean[j:j + 1] + dist.std[j:j + 1] * e1) * o.vae.config.scaling_factor
                            ac = o.scheduler.alphas_cumprod[int(t)]
                            self.strip_cache[key] = (float(ac ** 0.5) * z + float((1 - ac) ** 0.5) * e2).float().contiguous()
                if upcast:
                    o.vae.to(dtype=torch.float16)
        finally:
            torch.default_generator.set_state(cpu_state)
            if cuda_gen is not None:
                cuda_gen.set_state(cuda_state)
        return calls

    def pad_events(self, inner_h, inner_w, t):
        """Background strips of one padded `unet_step` (ed:405-408, 372-389): width strips first (dim 3), then height
        strips over the widened tensor (dim 2).  Returns [left, right, top, bottom] (None where empty)."""
        nat = self.geo.native
        l, r = pad_split(nat, inner_w)
        tp, b = pad_split(nat, inner_h)
        if l + r + tp + b == 0:
            return [None] * 4
        left = self._strip(inner_h, l, t, "3_1")
        right = self._strip(inner_h, r, t, "3_2")
        wide = inner_w + l + r
        top = self._strip(tp, wide, t, "2_1")
        bottom = self._strip(b, wide, t, "2_2")
        return [left, right, top, bottom]

    # -- per-cell pick indices (ed:502-520, 534-544, 673-675) -------------------------------------------------------
    def _draw_cells(self, exclude):
        """ed:502-520 with identical generator consumption: `exclude` is an (n,4) bool numpy array.  The draws are the
        same torch.randint calls (sizes n, then the number of still-invalid cells per round, <= 50 rounds, then the
        final unconditional redraw); the bookkeeping runs in numpy, and once only cells with all four picks excluded
        remain invalid the remaining rounds are just their (parity-relevant) draws without index work."""
        n = self.n_cells
        idx = torch.randint(0, 4, (n,)).numpy()
        # only the still-invalid cells are re-examined each round (ascending cell order = the order the reference's
        # boolean-mask assignment fills them in), so the work shrinks geometrically with the rounds
        bad = np.flatnonzero(exclude[self._rows, idx])
        m = bad.size
        rounds = 50
        hopeless = None
        while m > 0 and rounds > 0:
            if hopeless is None:
                hopeless = exclude.all(axis=1)
            if hopeless[bad].all():
                for _ in range(rounds):
                    last = torch.randint(0, 4, (m,))
                idx[bad] = last.numpy()
                rounds = 0
                break
            new = torch.randint(0, 4, (m,)).numpy()
            idx[bad] = new
            bad = bad[exclude[bad, new]]
            m = bad.size
            rounds -= 1
        if m > 0:
            idx[bad] = torch.randint(0, 4, (m,)).numpy()
        return idx

    def global_pass(self, t, resampling_steps, drop_p):
        """All RNG of one `approximate_latent_direction_w_resampling` call (ed:650-690).
        Returns (idx uint8 host tensor (R+1, n_cells), strips)."""
        g = self.geo
        n = self.n_cells
        out = np.zeros((resampling_steps + 1, n), dtype=np.uint8)
        exclude = np.zeros((n, 4), dtype=bool)
        rows = self._rows
        prev = np.zeros(n, dtype=np.int64)                                         # k = 0: top-left pick (ed:536)
        strips = None
        thr = 100 * drop_p
        tm, hm = time.perf_counter, self.host_ms
        for k in range(resampling_steps + 1):
            if k > 0:
                t0 = tm()
                idx = self._draw_cells(exclude)
                t1 = tm()
                drop = self._drop_draw()                                           # ed:541
                hm["cells"] += 1e3 * (t1 - t0)
                hm["drop"] += 1e3 * (tm() - t1)
                # thresholds applied to the TORCH tensor like the reference: an int64 tensor against a Python float
                # compares in float32, and 100 * drop_p is not always representable (0.8 -> 19.999999999999996)
                drop[drop <= thr] = 0                                              # ed:542
                drop[drop >= thr] = 1                                              # ed:543 -> P(new) = 30/101 at new_p = 0.3
                drop = drop.numpy()
                prev = idx * drop + prev * (1 - drop)                              # ed:544
            exclude[rows, prev] = True                                             # ed:675
            out[k] = prev.astype(np.uint8)
            t0 = tm()
            strips = self.pad_events(g.lh, g.lw, t)                                # unet_step of this iteration
            hm["strips"] += 1e3 * (tm() - t0)
        return torch.from_numpy(out), strips

    def local_pass(self, t, view_batch_size):
        """RNG side effects of `compute_local_uncond_signal` (ed:830-850): one padded unet_step per view chunk."""
        g = self.geo
        strips = [None] * 4
        if g.vh < g.native or g.vw < g.native:
            for _ in range(0, g.nv, max(1, view_batch_size)):
                strips = self.pad_events(g.vh, g.vw, t)
        return strips

    def undo_noise(self, n, shape, out):
        """The n = num_train/num_inference draws of undo_step (ed:695-701), in order."""
        t0 = time.perf_counter()
        for k in range(n):
            out[k].copy_(torch.randn(shape, device=self.rng_dev, dtype=torch.float32), non_blocking=True)
        self.host_ms["undo"] += 1e3 * (time.perf_counter() - t0)
        return out


# ---------------------------------------------------------------------------------------------------------------
class PeerOut:
    """UNet outputs of a sharded wave left in the ranks' symmetric buffers (read by ed_wave_epilogue_peer)."""

    def __init__(self, ptrs, world, per, dtype):
        self.ptrs, self.world, self.per, self.dtype = ptrs, world, per, dtype


class ElasticDiffusion(nn.Module):
    def __init__(self, device, sd_version='2.0',
                 verbose=False,
                 log_freq=5,
                 view_batch_size=1,
                 low_vram=False):
        super().__init__()
        self._init_common(device, sd_version, verbose, log_freq, view_batch_size, low_vram)
        print('[INFO] loading stable diffusion...')
        try:
            from diffusers import AutoencoderKL, DDIMScheduler, UNet2DConditionModel
            from transformers import CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer
        except Exception as e:  # diffusers is not part of this image: say so instead of failing obscurely
            raise ImportError(
                "ElasticDiffusion(device, sd_version, ...) loads Stable Diffusion through `diffusers` like the "
                "reference (ed:144-153); `diffusers` is not importable here. Install it, or inject modules with "
                "ElasticDiffusion.from_components(device, unet=..., vae=..., scheduler=..., text_embeds_fn=...)."
            ) from e
        model_key = MODEL_KEYS.get(self.sd_version)
        if model_key is None:
            print(f'[INFO] using hugging face custom model key: {self.sd_version}')
            model_key = self.sd_version
        home = 'cpu' if self.low_vram else self.device
        self.vae = AutoencoderKL.from_pretrained(model_key, subfolder="vae", torch_dtype=self.torch_dtype).to(home)
        self.tokenizer = [CLIPTokenizer.from_pretrained(model_key, subfolder="tokenizer")]
        self.text_encoder = [CLIPTextModel.from_pretrained(model_key, subfolder="text_encoder",
                                                           torch_dtype=self.torch_dtype).to(home)]
        self.unet = UNet2DConditionModel.from_pretrained(model_key, subfolder="unet", torch_dtype=self.torch_dtype).to(home)
        if self.sd_version == 'XL1.0':
            self.text_encoder.append(CLIPTextModelWithProjection.from_pretrained(
                model_key, subfolder="text_encoder_2", torch_dtype=self.torch_dtype).to(home))
            self.tokenizer.append(CLIPTokenizer.from_pretrained(model_key, subfolder="tokenizer_2"))
        self.scheduler = DDIMScheduler.from_pretrained(model_key, subfolder="scheduler")
        self._finish_init()
        print('[INFO] loaded stable diffusion!')

    # -- construction helpers ---------------------------------------------------------------------------------------
    def _init_common(self, device, sd_version, verbose, log_freq, view_batch_size, low_vram):
        self.device = torch.device(device)
        self.sd_version = sd_version
        self.verbose = verbose
        self.torch_dtype = torch.float16 if low_vram else torch.float32            # ed:121
        self.view_batch_size = view_batch_size
        self.log_freq = log_freq
        self.low_vram = low_vram
        # B200-native knobs (additive; defaults reproduce the reference's behaviour)
        self.rng_device = None        # None: draw on self.device like the reference; cpu: test mode (CPU goldens)
        self.autocast = True          # reference runs the loop under torch.autocast('cuda') (ed:1012)
        self.unet_batch_limit = None  # max samples per UNet call (None: the whole wave in one call)
        self.dist_group = None        # torch.distributed group to shard wave samples over (None: WORLD if initialised)
        self.shard_waves = True
        self.exchange = "p2p"         # multi-GPU exchange of UNet outputs: "p2p" = symmetric-memory buffers read by the fused
                                      # epilogue over NVLink (no collective), "nccl" = all_gather_into_tensor per wave
        self._sym = None
        self._sym_decode = None
        self._low_vram_limit = None
        self.unet_input_dtype = None  # dtype the gather kernels write the UNet batch in (None: fp32 like the reference)
        # tiled decode: None = the reference's tiles (core = sample_size // 4, pad = 3 * sample_size // 8: every pixel is
        # decoded 16x at SDXL); (core, pad) in latent units = opt-in de-duplicated decode with larger cores / smaller halos
        # (row f1).  NOT result-preserving: GroupNorm statistics and the mid-block attention are per decoded tile
        # (measured: profiles/r2_decode_dedup.json), so the default stays the reference geometry.
        self.decode_tile_geometry = None
        self.precompute_strips = True  # background strips of all timesteps built before the loop in batched VAE encodes
        self.use_cuda_graphs = False  # capture each wave's UNet forward in a CUDA graph (static canvas / text buffers)
        self._graphs = {}
        self.profile_kernels = False  # record CUDA-event pairs around every libelastic_b200 launch (bench.py)
        self._kernel_events = {}
        self.last_run = {}            # counters of the last generate_image call (kernel launches, UNet calls ...)
        self._text_embeds_fn = None
        self._projection_dim = None
        self.controlnet = None        # set by the ControlNet twin (controlnet.py)

    def _finish_init(self):
        for p in self.vae.parameters():                                            # ed:154,173-175
            p.requires_grad = False
        self.set_view_config()
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)  # ed:156
        check_scheduler(self.scheduler)

    @classmethod
    def from_components(cls, device, unet, vae, scheduler=None, text_embeds_fn=None, sd_version='2.1',
                        verbose=False, log_freq=5, view_batch_size=1, low_vram=False, projection_dim=None,
                        controlnet=None, controlnet_model='canny'):
        """Build from already-instantiated modules (what `__init__` gets from `from_pretrained`, ed:144-153).
        `text_embeds_fn(prompts) -> (text_embeddings, pooled)` replaces the CLIP encoders (ed:255-265)."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self._init_common(device, sd_version, verbose, log_freq, view_batch_size, low_vram)
        self.unet, self.vae = unet, vae
        self.scheduler = scheduler if scheduler is not None else DDIMSchedule()
        self._text_embeds_fn = text_embeds_fn
        self._projection_dim = projection_dim
        self.tokenizer, self.text_encoder = [], []
        self.controlnet, self.controlnet_model = controlnet, controlnet_model
        self._finish_init()
        return self

    # -- small reference API ------------------------------------------------------------------------------------------
    def set_view_config(self, patch_size=None):
        """ed:159-163."""
        ws = patch_size if patch_size is not None else self.unet.config.sample_size // 2
        self.view_config = {"window_size": ws, "stride": ws}
        self.view_config["context_size"] = self.unet.config.sample_size - self.view_config["window_size"]

    def seed_everything(self, seed, seed_np=True):
        """ed:165-171."""
        torch.manual_seed(seed)
        if self.device.type == 'cuda':
            torch.cuda.manual_seed(seed)
        if seed_np:
            np.random.seed(seed)

    def requires_grad(self, model, flag=True):
        for p in model.parameters():
            p.requires_grad = flag

    def upcast_vae(self):
        """ed:178-195 (fp16 VAE with force_upcast): decode in fp32."""
        self.vae.to(dtype=torch.float32)

    def get_downsample_size(self, H, W):
        """ed:943-950."""
        return low_res_size(H, W, self.sd_version, self.vae_scale_factor)

    def get_views(self, panorama_height, panorama_width, h_ws=64, w_ws=64, stride=32, **kwargs):
        """ed:198-229 (pixel sizes in, latent windows out)."""
        from .geometry import sliding_windows
        sf = self.vae_scale_factor
        if panorama_height % sf or panorama_width % sf:
            raise TypeError(f"height {panorama_height} and Width {panorama_width} must be divisable by {sf}")
        return sliding_windows(panorama_height // sf, panorama_width // sf, h_ws, w_ws, stride)[0]

    def encoder_prompt(self, prompt, encoder_id):
        tok = self.tokenizer[encoder_id](prompt, padding='max_length',
                                         max_length=self.tokenizer[encoder_id].model_max_length,
                                         truncation=True, return_tensors='pt')
        return self.text_encoder[encoder_id](tok.input_ids.to(self.device), output_hidden_states=True)

    @torch.no_grad()
    def get_text_embeds(self, prompt):
        """ed:255-265."""
        if self._text_embeds_fn is not None:
            return self._text_embeds_fn(prompt)
        if self.sd_version == 'XL1.0':
            emb = torch.cat([self.encoder_prompt(prompt, 0).hidden_states[-2],
                             self.encoder_prompt(prompt, 1).hidden_states[-2]], dim=-1)
            return emb, self.encoder_prompt(prompt, 1)[0]
        emb = self.encoder_prompt(prompt, 0)[0]
        return emb, emb

    def _get_add_time_ids(self, original_size, crops_coords_top_left, target_size, dtype):
        """ed:232-246."""
        ids = list(original_size + crops_coords_top_left + target_size)
        proj = self._projection_dim if self._projection_dim is not None else self.text_encoder[1].config.projection_dim
        passed = self.unet.config.addition_time_embed_dim * len(ids) + proj
        expected = self.unet.add_embedding.linear_1.in_features
        if expected != passed:
            raise ValueError(
                f"Model expects an added time embedding vector of length {expected}, but a vector of {passed} was "
                "created. The model has an incorrect config. Please check `unet.config.time_embedding_type` and "
                "`text_encoder_2.config.projection_dim`.")
        return torch.tensor([ids], dtype=dtype)

    # -- decode ---------------------------------------------------------------------------------------------------------
    def decode_latents(self, latents):
        """ed:267-272."""
        latents = latents.to(next(iter(self.vae.post_quant_conv.parameters())).dtype)
        latents = latents / self.vae.config.scaling_factor
        imgs = self.vae.decode(latents).sample
        return (imgs / 2 + 0.5).clamp(0, 1)

    def tiled_decode(self, latents, tile_batch=16):
        """ed:275-310 with the slicing / padding / blending done by ed_tile_gather (TMA, OOB zero fill) and
        ed_tile_blend; the VAE decoder itself runs through PyTorch, `tile_batch` tiles per call.
        With torch.distributed initialised (and `shard_waves`) the tiles are sharded over the ranks: rank r
        decodes patches [r*per, (r+1)*per) into a symmetric-memory buffer and ed_tile_blend_peer reads the centre crops
        straight from their owners over NVLink (exchange "p2p"), or the centre crops are all-gathered ("nccl")."""
        self._require_cuda()
        L = native.lib()
        latents = latents.float().contiguous()
        B, C, H, W = latents.shape
        dc, dp = self.decode_tile_geometry if self.decode_tile_geometry is not None else (None, None)
        tg = build_tiles(H, W, self.unet.config.sample_size, self.vae_scale_factor, self.low_vram, core=dc, pad=dp)
        dev = latents.device
        tabs = {k: _i32(dev, v) for k, v in tg.tables.items()}
        T = tg.core + 2 * tg.pad
        nt = len(tg.tiles)
        n = nt * B
        grp, rank, ws = self._dist()
        per, lo, hi = shard_range(n, ws, rank)
        boxes = torch.empty(n, C, T, T, device=dev, dtype=torch.float32)
        native.check(L.ed_tile_gather(native.ptr(latents), B, C, H, W, native.ptr(tabs["tiles"]), nt, tg.core, tg.pad,
                                      native.ptr(boxes), native.stream_handle()), "ed_tile_gather")
        wdt = next(iter(self.vae.post_quant_conv.parameters())).dtype
        sf = self.vae_scale_factor
        cfg = self.vae.config
        CH = int(cfg.get("out_channels", 3)) if hasattr(cfg, "get") else int(getattr(cfg, "out_channels", 3))
        # the patches THIS rank decodes, (per, CH, T*sf, T*sf) in the decoder's dtype; sharded + p2p: a symmetric buffer
        shape = (max(per, 1), CH, T * sf, T * sf)
        sym = self._decode_symmetric(shape, wdt, grp) if (ws > 1 and self.exchange == "p2p") else None
        patches = sym["buf"] if sym else torch.empty(shape, device=dev, dtype=wdt)
        for s in range(lo, hi, tile_batch):
            e = min(s + tile_batch, hi)
            z = boxes[s:e].to(wdt) / self.vae.config.scaling_factor
            patches[s - lo:e - lo] = self.vae.decode(z).sample
        self.last_run["decode_tiles"] = self.last_run.get("decode_tiles", 0) + (hi - lo)
        image = torch.empty(B, CH, H * sf, W * sf, device=dev, dtype=torch.float32)
        mk = lambda pad: native.Tiles(ntiles=nt, ntc=tg.ntc, core=tg.core, pad=pad, scale=sf, B=B, CH=CH, H=H, W=W,
                                      tiles=tabs["tiles"].data_ptr(), trow_first=tabs["trow_first"].data_ptr(),
                                      trow_cnt=tabs["trow_cnt"].data_ptr(), tcol_first=tabs["tcol_first"].data_ptr(),
                                      tcol_cnt=tabs["tcol_cnt"].data_ptr())
        if ws == 1:
            tt = mk(tg.pad)
            native.check(L.ed_tile_blend(ctypes.byref(tt), native.ptr(patches), native.dtype_code(wdt),
                                         native.ptr(image), native.stream_handle()), "ed_tile_blend")
        elif sym:
            tt = mk(tg.pad)
            sym["hdl"].barrier(channel=0)                                  # every rank's patches are in its buffer
            native.check(L.ed_tile_blend_peer(ctypes.byref(tt), native.ptr(sym["ptrs"]), ws, per, native.dtype_code(wdt),
                                              native.ptr(image), native.stream_handle()), "ed_tile_blend_peer")
            sym["hdl"].barrier(channel=0)                                  # nobody overwrites a buffer a peer still reads
            self.last_run["peer_exchanges"] = self.last_run.get("peer_exchanges", 0) + 1
        else:
            # "nccl": all-gather of the centre crops (1/16 of the padded patches), then the same blend with pad = 0
            import torch.distributed as dist
            p0, c = tg.pad * sf, tg.core * sf
            send = patches[:, :, p0:p0 + c, p0:p0 + c].contiguous()
            crops = torch.empty(ws * per, CH, c, c, device=dev, dtype=wdt)
            dist.all_gather_into_tensor(crops, send, group=grp)
            self.last_run["collectives"] = self.last_run.get("collectives", 0) + 1
            tt = mk(0)
            native.check(L.ed_tile_blend(ctypes.byref(tt), native.ptr(crops), native.dtype_code(wdt), native.ptr(image),
                                         native.stream_handle()), "ed_tile_blend")
        self.last_run["kernel_launches"] = self.last_run.get("kernel_launches", 0) + 2
        return image

    def _decode_symmetric(self, shape, dtype, grp):
        """Symmetric-memory buffer for this rank's decoded patches + the peer pointer table (cached per shape).
        Falls back to the NCCL crop all-gather (with a note in last_run) when symmetric memory is unavailable."""
        ent = self._sym_decode
        if ent is not None and (ent["shape"], ent["dtype"]) == (shape, dtype):
            return ent
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            group = grp if grp is not None else dist.group.WORLD
            buf = symm.empty(shape, dtype=dtype, device=self.device)
            hdl = symm.rendezvous(buf, group)
            ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
            self._sym_decode = dict(buf=buf, hdl=hdl, ptrs=ptrs, shape=shape, dtype=dtype)
        except Exception as e:  # pragma: no cover - depends on the box
            self.last_run["exchange_fallback"] = f"symmetric memory unavailable ({type(e).__name__}: {e}); using nccl all_gather"
            self.exchange = "nccl"
            self._sym_decode = None
        return self._sym_decode

    # -- the hot path ---------------------------------------------------------------------------------------------------
    def kernel_times_ms(self, reset=True):
        """{kernel name: (launches, total ms)} from the CUDA events recorded while `profile_kernels` was on."""
        torch.cuda.synchronize()
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self._kernel_events.items()}
        if reset:
            self._kernel_events = {}
        return out

    def _require_cuda(self):
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise native.NativeError(
                "this ElasticDiffusion runs its denoising loop on sm_100a CUDA kernels only (no CPU / eager "
                f"fallback); got device={self.device}, cuda available={torch.cuda.is_available()}")
        native.lib()

    def _upload_plan(self, geo):
        return native.plan_from_geometry(geo, self.device)

    def _dist(self):
        import torch.distributed as dist
        if not self.shard_waves or not dist.is_available() or not dist.is_initialized():
            return None, 0, 1
        grp = self.dist_group
        ws = dist.get_world_size(grp)
        return (grp, dist.get_rank(grp), ws) if ws > 1 else (None, 0, 1)

    def _graphed(self, key, call):
        """Replay (capturing on first use) `call` as a CUDA graph; its inputs must live in static buffers."""
        ent = self._graphs.get(key)
        if ent is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):                  # warm-up outside capture (cuDNN / cuBLAS plan selection)
                for _ in range(2):
                    call()
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = call()
            ent = self._graphs[key] = (g, out)
        ent[0].replay()
        return ent[1]

    def _unet(self, canvas, t, text, pooled, time_ids, cond=None, cond_scale=1.0):
        """One batched UNet evaluation of a wave (dense part, through PyTorch; ed:417-426 for the XL kwargs).
        With torch.distributed initialised the samples are sharded over ranks and all-gathered (multi-GPU)."""
        grp, rank, ws = self._dist()
        n = canvas.shape[0]
        per, lo, hi = shard_range(n, ws, rank)
        outs = []
        # samples per UNet call: the whole (rank-local part of the) wave by default.  `view_batch_size` is the reference's
        # memory knob for the view passes (ed:830-850, its global passes are always batch 2B): it is honoured as the
        # batch bound when the caller asked for the memory-saving mode (low_vram) or set `unet_batch_limit` explicitly.
        limit = self.unet_batch_limit or (self._low_vram_limit if self.low_vram else None) or max(hi - lo, 1)
        for s in range(lo, hi, limit):
            e = min(s + limit, hi)
            kw = {}
            if time_ids is not None:
                kw["added_cond_kwargs"] = {"text_embeds": pooled[s:e], "time_ids": time_ids[s:e]}

            def call(s=s, e=e, kw=kw):
                with torch.autocast("cuda", enabled=self.autocast):
                    res = {}
                    if cond is not None:   # ControlNet forward feeding residuals into the UNet (cn:482-496, 506-518)
                        down, mid = self.controlnet(canvas[s:e], t, encoder_hidden_states=text[s:e],
                                                    controlnet_cond=cond[s:e], conditioning_scale=cond_scale,
                                                    guess_mode=False, return_dict=False, **kw)
                        res = {"down_block_additional_residuals": down, "mid_block_additional_residual": mid}
                    return self.unet(canvas[s:e], t, encoder_hidden_states=text[s:e], **kw, **res)["sample"]
            if self.use_cuda_graphs:
                outs.append(self._graphed(("unet", canvas.data_ptr(), text.data_ptr(), s, e), call))
            else:
                outs.append(call())
            self.last_run["unet_calls"] += 1
            self.last_run["unet_samples"] += e - s
        if ws == 1:
            out = outs[0] if len(outs) == 1 else torch.cat(outs)
            return out.contiguous()
        import torch.distributed as dist
        mine = torch.cat(outs) if outs else canvas.new_zeros((0,) + tuple(canvas.shape[1:]))
        if self._shard_dtype is None:
            # dtype of the UNet output must agree on every rank, including ranks that got no sample
            probe = torch.tensor([native.dtype_code(mine.dtype) if mine.numel() else -1], device=self.device)
            dist.all_reduce(probe, op=dist.ReduceOp.MAX, group=grp)
            self._shard_dtype = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}[int(probe.item())]
        if self.exchange == "p2p" and self._sym is None:
            self._sym = self._setup_symmetric(per, tuple(canvas.shape[1:]), self._shard_dtype, grp, ws)
        if self.exchange == "p2p" and self._sym:
            # leave the local outputs in this rank's symmetric buffer; the epilogue kernel of every rank reads the samples
            # it needs straight from their owners over NVLink after a device-side barrier (no all-gather)
            sym = self._sym
            slot = sym["slot"] = sym["slot"] ^ 1
            if per > sym["per_max"]:
                raise RuntimeError("symmetric exchange buffer too small for this wave")
            sym["bufs"][slot][:hi - lo].copy_(mine)
            sym["hdls"][slot].barrier(channel=0)
            self.last_run["peer_exchanges"] = self.last_run.get("peer_exchanges", 0) + 1
            return PeerOut(sym["ptrs"][slot], ws, per, self._shard_dtype)
        send = torch.zeros((per,) + tuple(canvas.shape[1:]), device=self.device, dtype=self._shard_dtype)
        send[:hi - lo] = mine
        gathered = torch.empty((ws * per,) + tuple(canvas.shape[1:]), device=self.device, dtype=self._shard_dtype)
        dist.all_gather_into_tensor(gathered, send, group=grp)
        self.last_run["collectives"] += 1
        return gathered[:n]

    def _setup_symmetric(self, per_max, tail, dtype, grp, ws):
        """Two symmetric-memory buffers (alternating per wave) for the rank-local UNet outputs + peer pointer tables.
        Falls back to the NCCL all-gather (with a note in last_run) when symmetric memory is unavailable."""
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            group = grp if grp is not None else dist.group.WORLD
            bufs, hdls, ptrs = [], [], []
            for _ in range(2):
                t = symm.empty((per_max,) + tail, dtype=dtype, device=self.device)
                h = symm.rendezvous(t, group)
                bufs.append(t)
                hdls.append(h)
                ptrs.append(torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64, device=self.device))
            return dict(bufs=bufs, hdls=hdls, ptrs=ptrs, slot=0, per_max=per_max)
        except Exception as e:  # pragma: no cover - depends on the box
            self.last_run["exchange_fallback"] = f"symmetric memory unavailable ({type(e).__name__}: {e}); using nccl all_gather"
            self.exchange = "nccl"
            return {}

    # -- verbose-mode diagnostics (ed:759-796, 1093-1118): plain torch, outside the hot path -----------------------------------
    @torch.no_grad()
    def generate(self, latent, text_embeds, add_text_embeds, guidance_scale=7.5):
        """ed:759-796: a plain CFG + DDIM run on the (low-resolution) latent, logged as `global_img` in verbose mode.
        Every UNet call pads the latent to the native size with the background strips of the timestep (ed:405-408)."""
        vb = getattr(self, "_verbose", None)
        if vb is None or vb.get("ledger") is None:
            raise RuntimeError("generate() replays the run's background strips: call it after generate_image(verbose=True)")
        ledger, inter = vb["ledger"], []
        B = latent.shape[0]
        nat = ledger.geo.native
        xl = self.sd_version.startswith('XL')
        if self.low_vram:
            self.vae.cpu()
            self.unet.to(self.device)
        for i, t in enumerate(tqdm(self.scheduler.timesteps)):
            x2 = torch.cat([latent] * 2)
            left, right, top, bottom = ledger.pad_events(latent.shape[-2], latent.shape[-1], t)     # ed:405-408, 366-391
            lp, _ = pad_split(nat, latent.shape[-1])
            tp, _ = pad_split(nat, latent.shape[-2])
            ex = lambda s: s.to(x2.dtype).expand(x2.shape[0], -1, -1, -1)
            if left is not None or right is not None:
                x2 = torch.cat(([ex(left)] if left is not None else []) + [x2] + ([ex(right)] if right is not None else []), dim=3)
            if top is not None or bottom is not None:
                x2 = torch.cat(([ex(top)] if top is not None else []) + [x2] + ([ex(bottom)] if bottom is not None else []), dim=2)
            kw = {}
            if xl:
                ids = self._get_add_time_ids(self.default_size, (0, 0), self.default_size, dtype=text_embeds.dtype)
                kw["added_cond_kwargs"] = {"text_embeds": add_text_embeds, "time_ids": ids.to(self.device).repeat(2 * B, 1)}
            with torch.autocast("cuda", enabled=self.autocast):
                eps = self.unet(x2, t, encoder_hidden_states=text_embeds, **kw)["sample"]
            eps = eps[:, :, tp:tp + latent.shape[-2], lp:lp + latent.shape[-1]]
            un, co = eps.chunk(2)
            eps = un + guidance_scale * (co - un)
            sc = {k: torch.tensor(v, dtype=torch.float32, device=latent.device) for k, v in step_scalars(self.scheduler, t).items()}
            x0 = (latent - sc["sqrt_beta_t"] * eps) / sc["sqrt_alpha_t"]
            latent = sc["sqrt_alpha_prev"] * x0 + sc["sqrt_dir"] * eps
            if i % self.log_freq == 0:
                inter.append(x0.float().cpu())
        needs_upcasting = self.vae.dtype == torch.float16 and self.vae.config.force_upcast
        if self.low_vram:
            self.unet.cpu()
            self.vae.to(self.device)
        if needs_upcasting:
            self.upcast_vae()
        img = _to_pil(self.decode_latents(latent.float())[0])
        if needs_upcasting:
            self.vae.to(dtype=torch.float16)
        return img, {"inter_x0": inter}

    def _rrg_reference_x0(self, geo, x_in, out, idx_dev, owner, R1, B, g, sc, fp16sem):
        """cascade_info['x0'] of reduced_resolution_guidance (ed:909-921) for the verbose log: the low-res DDIM x0 the fused
        epilogue evaluates per cell, restated with torch ops from the same inputs."""
        dev = x_in.device
        T_ = lambda k: torch.tensor(geo.tables[k], dtype=torch.long, device=dev)
        lp, rp, tp, bp = geo.g_pad
        lh, lw, C = geo.lh, geo.lw, geo.C
        f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
        kl = R1 - 1
        pick = idx_dev[kl].long().view(lh, lw)
        r, c = torch.arange(lh, device=dev)[:, None], torch.arange(lw, device=dev)[None, :]
        xl = x_in[:, :, T_("row_src")[2 * r + pick // 2], T_("col_src")[2 * c + pick % 2]]
        glob = out[:2 * B * R1].reshape(R1, 2, B, C, geo.native, geo.native)[..., tp:tp + lh, lp:lp + lw].float()
        ksel = owner.view(geo.H, geo.W).long()[T_("down_row")][:, T_("down_col")]                    # (lh, lw)
        pickk = lambda s: torch.gather(glob[:, s], 0, ksel[None, None, None].expand(1, B, C, lh, lw))[0]
        d = pickk(1) - pickk(0)
        rnd = (lambda v: v.half().float()) if fp16sem else (lambda v: v)
        gl = rnd(f32(g) * rnd(d))
        el = rnd(glob[kl, 0] + gl)
        return (xl - rnd(f32(sc["sqrt_beta_t"]) * el)) / f32(sc["sqrt_alpha_t"])

    @torch.no_grad()
    def generate_image(self, prompts, negative_prompts='',
                       height=768, width=768,
                       num_inference_steps=50,
                       guidance_scale=10.0,
                       resampling_steps=20,
                       new_p=0.3, rrg_stop_t=0.2,
                       rrg_init_weight=1000,
                       rrg_scherduler_cls=CosineScheduler,
                       cosine_scale=3.0,
                       repaint_sampling=True,
                       progress=tqdm,
                       tiled_decoder=False,
                       grid=False):
        latent, image_log = self.denoise(prompts, negative_prompts, height, width, num_inference_steps, guidance_scale,
                                         resampling_steps, new_p, rrg_stop_t, rrg_init_weight, rrg_scherduler_cls,
                                         cosine_scale, repaint_sampling, progress)
        # ---- decode (ed:1080-1130) ------------------------------------------------------------------------------------
        needs_upcasting = self.vae.dtype == torch.float16 and self.vae.config.force_upcast
        if self.low_vram:
            self.unet.cpu()
            self.vae.to(self.device)
        if needs_upcasting:
            self.upcast_vae()
        if tiled_decoder == "dedup" and self.decode_tile_geometry is None:       # opt-in extension of the boolean flag
            self.decode_tile_geometry = (self.unet.config.sample_size // 2, self.unet.config.sample_size // 4)
        decode_fn = self.tiled_decode if tiled_decoder else self.decode_latents
        if self.verbose:                                                                       # ed:1093-1118
            vb = self._verbose
            if vb.get("init_low") is not None:
                image_log['global_img'], info = self.generate(vb["init_low"], vb["text"], vb["pooled"],
                                                              guidance_scale=guidance_scale)
                if info.get('inter_x0'):
                    x0s = torch.cat([decode_fn(z.to(self.device)) for z in info['inter_x0']])
                    image_log['global_img_inter_x0_imgs'] = _to_pil(_grid(x0s))
            if self._x0_log:
                x0s = torch.cat([decode_fn(z.to(self.device)) for z in self._x0_log]).clip(0, 1)
                image_log['intermediate_x0_imgs'] = _to_pil(_grid(x0s))
            image_log['intermediate_cascade_x0_imgs'] = {}
            if vb.get("cascade"):
                x0s = torch.cat([decode_fn(z.to(self.device)) for z in vb["cascade"]])
                image_log['intermediate_cascade_x0_imgs']['rrg'] = _to_pil(_grid(x0s))
        imgs = torch.cat([decode_fn(latent[i:i + 1]) for i in range(len(latent))])   # ed:1121 (one sample at a time)
        if grid:
            imgs = [_grid(imgs)]
        imgs = [_to_pil(img) for img in imgs]
        if needs_upcasting:
            self.vae.to(dtype=torch.float16)
        return imgs, image_log

    @torch.no_grad()
    def denoise(self, prompts, negative_prompts='', height=768, width=768, num_inference_steps=50,
                guidance_scale=10.0, resampling_steps=20, new_p=0.3, rrg_stop_t=0.2, rrg_init_weight=1000,
                rrg_scherduler_cls=CosineScheduler, cosine_scale=3.0, repaint_sampling=True, progress=tqdm,
                step_callback=None, max_steps=None, condition_image=None, controlnet_conditioning_scale=1.0):
        """The loop of ed:967-1078.  Returns (final latent (B,4,H/8,W/8) fp32 on device, image_log).
        `condition_image`: prepared ControlNet condition (1,3,ds_h*8,ds_w*8) float tensor (ControlNet twin, cn:1119-1322)."""
        self._require_cuda()
        L = native.lib()
        sf = self.vae_scale_factor
        # heights / widths that are not multiples of the VAE factor are floored like the reference does (ed:998 draws a
        # (height // 8, width // 8) latent and get_views only ever sees latent * 8, ed:817,827); only get_views() itself
        # raises on such sizes (ed:200-201)
        self.last_run = dict(kernel_launches=0, unet_calls=0, unet_samples=0, collectives=0, vae_encodes=0, steps=0)
        self._shard_dtype = None
        self._sym = None
        self._x0_log = []
        self._verbose = {}
        self._low_vram_limit = max(2, int(self.view_batch_size)) * (1 if isinstance(prompts, str) else len(prompts))
        ds = self.get_downsample_size(height, width)                                           # ed:968
        self.default_size = (4 * height, 4 * width)                                            # ed:969
        T = num_inference_steps
        n_rrg = T - int(T * rrg_stop_t)
        if rrg_scherduler_cls == CosineScheduler:                                              # ed:972-979
            rrg_w = rrg_scherduler_cls(steps=n_rrg, cosine_scale=cosine_scale, factor=rrg_init_weight)
        else:
            rrg_w = rrg_scherduler_cls(steps=n_rrg, start_val=rrg_init_weight, stop_val=0)
        if isinstance(prompts, str):
            prompts = [prompts]
        if isinstance(negative_prompts, str):
            negative_prompts = [negative_prompts] * len(prompts)
        if self.low_vram:
            self.vae.cpu()
            self.unet.cpu()
            self.text_encoder = [e.to(self.device) for e in self.text_encoder]
        un_text, un_pool = self.get_text_embeds(negative_prompts)                              # ed:992-993
        co_text, co_pool = self.get_text_embeds(prompts)
        un_text, un_pool, co_text, co_pool = (z.to(self.device) for z in (un_text, un_pool, co_text, co_pool))

        B, C = len(prompts), self.unet.config.in_channels
        H, W = height // sf, width // sf
        is_xl = self.sd_version.startswith('XL')
        native_size = 128 if is_xl else 64                                                     # ed:398-400
        vc = self.view_config
        geo = build_geometry(B, C, H, W, native_size, ds, vc["window_size"], vc["stride"], vc["context_size"])
        if (geo.lh, geo.lw) != tuple(ds):
            raise ValueError(f"resampler produced {geo.lh}x{geo.lw}, expected {ds}")
        plan, _keep = self._upload_plan(geo)
        ledger = RngLedger(self, geo)
        dev = self.device

        x = ledger.initial_latent((B, C, H, W), self.torch_dtype)                               # ed:998
        self.scheduler.set_timesteps(T)                                                        # ed:1001
        ts = self.scheduler.timesteps
        if self.low_vram:
            self.text_encoder = [e.cpu() for e in self.text_encoder]
            self.vae.cpu()
            self.unet.to(dev)

        # all background strips of the run, batched, before the loop (row f2); RNG-neutral
        n_steps = len(ts) if max_steps is None else min(len(ts), max_steps)
        self.last_run["vae_encode_calls"] = ledger.precompute_strips(ts[:n_steps]) if self.precompute_strips else 0
        R = resampling_steps
        n_re = self.scheduler.config.num_train_timesteps // T
        if n_re > native.ED_MAX_RENOISE:
            raise ValueError(f"undo_step needs {n_re} forward steps > ED_MAX_RENOISE={native.ED_MAX_RENOISE}")
        nv = geo.nv
        # static per-call buffers (addresses fixed for the whole loop)
        in_dtype = self.unet_input_dtype or torch.float32
        n1, n2 = 2 * B * (R + 1) + nv * B, 2 * B + nv * B
        canvas = torch.empty(max(n1, n2), C, native_size, native_size, device=dev, dtype=in_dtype)
        x_mid, x_next = torch.empty_like(x), torch.empty_like(x)
        x0_buf = torch.empty_like(x) if self.verbose else None
        noise = torch.empty(n_re, B, C, H, W, device=dev, dtype=torch.float32)
        idx1 = torch.empty(R + 1, geo.lh * geo.lw, device=dev, dtype=torch.uint8)
        idx2 = torch.zeros(1, geo.lh * geo.lw, device=dev, dtype=torch.uint8)
        owner1 = torch.empty(H * W, device=dev, dtype=torch.uint8)     # per-wave owner maps (ed_owner_map)
        owner2 = torch.zeros(H * W, device=dev, dtype=torch.uint8)     # wave 2 has a single iteration: owner == 0
        d_params = torch.empty(2, ctypes.sizeof(native.StepParams), device=dev, dtype=torch.uint8)
        t_dev = torch.zeros((), device=dev, dtype=torch.int64)     # timestep for the UNet, static address (CUDA graphs)
        idx_pin = [torch.empty(R + 1, geo.lh * geo.lw, dtype=torch.uint8).pin_memory() for _ in range(3)]
        idx_ev = [None, None, None]
        self._graphs = {}
        text_pair, pool_pair = torch.cat([un_text, co_text]), torch.cat([un_pool, co_pool], dim=0)   # ed:996-997
        text1 = torch.cat([text_pair] * (R + 1) + [un_text] * nv)
        pool1 = torch.cat([pool_pair] * (R + 1) + [un_pool] * nv)
        text2 = torch.cat([text_pair] + [un_text] * nv)
        pool2 = torch.cat([pool_pair] + [un_pool] * nv)
        time_ids = None
        if is_xl:                                                                              # ed:413-420, hoisted
            time_ids = self._get_add_time_ids(self.default_size, (0, 0), self.default_size, dtype=text_pair.dtype)
            time_ids = time_ids.to(dev).repeat(max(n1, n2), 1)
        rrg_norm = float(torch.tensor(2.0 / (C * H * W), dtype=torch.float64).to(torch.float32))
        image_log = {}

        # ---- ControlNet condition batches (static for the whole call): one K12 launch per wave shape ----------------------
        cond1 = cond2 = None
        if condition_image is not None:
            if self.controlnet is None:
                raise ValueError("condition_image given but no ControlNet is loaded (use the elastic_diffusion_w_controlnet class)")
            if B != 1:
                raise ValueError("the ControlNet twin prepares the condition with batch_size=1 (cn:1187): one prompt per call")
            img = condition_image.to(device=dev, dtype=torch.float32)
            if tuple(img.shape) != (1, 3, ds[0] * sf, ds[1] * sf):
                raise ValueError(f"condition image must be (1, 3, {ds[0] * sf}, {ds[1] * sf}), got {tuple(img.shape)}")
            cdt = next(self.controlnet.parameters()).dtype
            prepared = torch.cat([img.to(cdt)] * 2).float().contiguous()                       # cn:1028-1031 (CFG doubling)
            row_map, col_map, origin = cond_geometry(geo, sf, vc["window_size"], vc["context_size"])
            ctabs = [_i32(dev, v) for v in (row_map, col_map, origin)]
            cond_dtype = self.unet_input_dtype or torch.float32

            def build_cond(R1):
                out_c = torch.empty(2 * B * R1 + nv * B, 3, native_size * sf, native_size * sf, device=dev, dtype=cond_dtype)
                native.check(L.ed_gather_cond(ctypes.byref(plan), R1, native.ptr(prepared), 3, sf, native.ptr(ctabs[0]),
                                              native.ptr(ctabs[1]), native.ptr(ctabs[2]), native.ptr(out_c),
                                              native.dtype_code(cond_dtype), native.stream_handle()), "ed_gather_cond")
                self.last_run["kernel_launches"] += 1
                return out_c
            cond1 = build_cond(R + 1)
            cond2 = cond1 if R == 0 else build_cond(1)

        def launch(name, fn, *args):
            """one libelastic_b200 kernel launch (+ optional CUDA-event bracket on the launching stream)"""
            if self.profile_kernels:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                native.check(fn(*args), name)
                e1.record()
                self._kernel_events.setdefault(name, []).append((e0, e1))
            else:
                native.check(fn(*args), name)
            self.last_run["kernel_launches"] += 1

        def run_wave(x_in, t, idx_dev, R1, strips_g, strips_v, prm, slot, noise_buf, x_out, x0_out, text, pool):
            st = native.stream_handle()
            n = 2 * B * R1 + nv * B
            cv = canvas[:n]
            owner = owner2
            if R1 > 1:
                owner = owner1
                launch("ed_owner_map", L.ed_owner_map, ctypes.byref(plan), R1, native.ptr(idx_dev), native.ptr(owner), st)
            launch("ed_random_pick_gather", L.ed_random_pick_gather, ctypes.byref(plan), R1, native.ptr(x_in),
                   native.ptr(idx_dev), native.strips_array(strips_g), native.ptr(cv), native.dtype_code(cv.dtype), st)
            launch("ed_gather_views", L.ed_gather_views, ctypes.byref(plan), native.ptr(x_in), native.ptr(cv),
                   native.dtype_code(cv.dtype), 2 * B * R1, st)
            if any(s is not None for s in strips_v):
                launch("ed_pad_views", L.ed_pad_views, ctypes.byref(plan), native.strips_array(strips_v), native.ptr(cv),
                       native.dtype_code(cv.dtype), 2 * B * R1, st)
            out = self._unet(cv, t_dev, text[:n], pool[:n], None if time_ids is None else time_ids[:n],
                             cond1 if R1 == R + 1 else cond2, controlnet_conditioning_scale)
            if out.dtype == torch.float16:
                prm.flags |= native.FLAG_FP16_SEM
            native.check(L.ed_upload_step_params(native.ptr(d_params[slot]), ctypes.byref(prm), st), "upload")
            tag = "+renoise" if prm.flags & 1 else "+rrg" if prm.flags & 2 else ""
            if isinstance(out, PeerOut):
                launch("ed_wave_epilogue_peer" + tag, L.ed_wave_epilogue_peer, ctypes.byref(plan),
                       native.ptr(d_params[slot]), R1, native.ptr(x_in), native.ptr(out.ptrs), out.world, out.per,
                       native.dtype_code(out.dtype), native.ptr(idx_dev), native.ptr(owner), native.ptr(noise_buf),
                       native.ptr(x_out), native.ptr(x0_out), st)
            else:
                launch("ed_wave_epilogue" + tag, L.ed_wave_epilogue, ctypes.byref(plan), native.ptr(d_params[slot]), R1,
                       native.ptr(x_in), native.ptr(out), native.dtype_code(out.dtype), native.ptr(idx_dev),
                       native.ptr(owner), native.ptr(noise_buf), native.ptr(x_out), native.ptr(x0_out), st)
            return out

        def plan_step(i):
            """All RNG of step i in the reference's order; touches no UNet output, so it is
            issued one step AHEAD and overlaps with the GPU work of step i-1."""
            t_host = time.perf_counter()
            t = ts[i]
            last = i == len(ts) - 1
            repaint = bool(repaint_sampling and R > 0 and not last)                            # ed:1038
            p = dict(t=t, repaint=repaint, w=rrg_w(i), sc=step_scalars(self.scheduler, t))
            idx_host, p["strips_g"] = ledger.global_pass(t, R, 1 - new_p)                       # ed:1016
            p["strips_v"] = ledger.local_pass(t, self.view_batch_size)                          # ed:1027
            slot = i % 3
            if idx_ev[slot] is not None:
                # the H2D copy that last used this pinned buffer has run.  This is also what keeps the host at most ~3 steps
                # ahead of the GPU: the wait is throttling, not planner work, and is accounted separately
                t_w = time.perf_counter()
                idx_ev[slot].synchronize()
                t_w = time.perf_counter() - t_w
                t_host += t_w
                self.last_run["throttle_wait_ms"] = self.last_run.get("throttle_wait_ms", 0.0) + 1e3 * t_w
            idx_pin[slot].copy_(idx_host)
            p["slot"] = slot
            if repaint:
                ledger.undo_noise(n_re, (B, C, H, W), noise)                                   # ed:1040
                _, p["strips_g2"] = ledger.global_pass(t, 0, 1 - new_p)                         # ed:1043
                p["strips_v2"] = ledger.local_pass(t, self.view_batch_size)                     # ed:1049
                p["renoise"] = renoise_scalars(self.scheduler, ts[i + 1])
            self.last_run["plan_host_ms"] = self.last_run.get("plan_host_ms", 0.0) + 1e3 * (time.perf_counter() - t_host)
            return p

        # NOTE on buffer reuse: `noise` and `idx1` are single device buffers; the draws / copies of step i+1 are
        # enqueued on the main stream AFTER the kernels of step i, so stream order protects them.
        nxt = None
        for i, t in enumerate(progress(ts)):       # lazily: the caller's progress bar advances once per finished step (ed:1013)
            if i >= n_steps:
                break
            p = nxt if nxt is not None else plan_step(i)
            nxt = None
            repaint, w, sc = p["repaint"], p["w"], p["sc"]
            rrg_on = w > 10                                                                    # ed:1062
            strips_g, strips_v = p["strips_g"], p["strips_v"]
            idx1.copy_(idx_pin[p["slot"]], non_blocking=True)
            idx_ev[p["slot"]] = torch.cuda.Event()
            idx_ev[p["slot"]].record()
            t_dev.fill_(int(t))

            if self.verbose and i == 0:                # init_downsampled_latent (ed:677-678): the k = 0 (top-left) picks
                rs = torch.tensor(geo.tables["row_src"][0::2], device=dev)
                cs = torch.tensor(geo.tables["col_src"][0::2], device=dev)
                self._verbose = dict(init_low=x[:, :, rs][:, :, :, cs].clone(), text=text_pair, pooled=pool_pair, ledger=ledger,
                                     cascade=[])
            # ---- wave 1 ----------------------------------------------------------------------------------------------------
            prm = native.StepParams(guidance=guidance_scale, rrg_weight=float(w), rrg_norm=rrg_norm, R1=R + 1, **sc)
            if repaint:
                strips_g2, strips_v2 = p["strips_g2"], p["strips_v2"]
                a, b = p["renoise"]
                prm.flags = native.FLAG_RENOISE
                prm.n_renoise = n_re
                for k in range(n_re):
                    prm.renoise_a[k], prm.renoise_b[k] = a[k], b[k]
                run_wave(x, t, idx1, R + 1, strips_g, strips_v, prm, 0, noise, x_mid, None, text1, pool1)
                # ---- wave 2 (repaint: one more estimate at the re-noised latent, guidance / 3; ed:1041-1056) --------
                prm2 = native.StepParams(guidance=guidance_scale / 3, rrg_weight=float(w), rrg_norm=rrg_norm, R1=1, **sc)
                prm2.flags = native.FLAG_RRG if rrg_on else 0
                out_w = run_wave(x_mid, t, idx2, 1, strips_g2, strips_v2, prm2, 1, None, x_next, x0_buf, text2, pool2)
                rrg_in = (x_mid, idx2, owner2, 1, prm2)
            else:
                prm.flags = native.FLAG_RRG if rrg_on else 0
                out_w = run_wave(x, t, idx1, R + 1, strips_g, strips_v, prm, 0, None, x_next, x0_buf, text1, pool1)
                rrg_in = (x, idx1, owner1 if R > 0 else owner2, R + 1, prm)
            if self.verbose and rrg_on and i % self.log_freq == 0 and torch.is_tensor(out_w):   # ed:1074-1077 (not when sharded)
                xi, ix, ow, r1, pr = rrg_in
                self._verbose["cascade"].append(self._rrg_reference_x0(
                    geo, xi, out_w, ix, ow, r1, B, pr.guidance, sc, bool(pr.flags & native.FLAG_FP16_SEM)).cpu())
            if i + 1 < n_steps:
                nxt = plan_step(i + 1)          # look-ahead: host RNG chain of the next step runs under this step's GPU work
            x, x_next = x_next, x                                                              # ed:1078
            self.last_run["steps"] += 1
            if self.verbose and i % self.log_freq == 0:
                self._x0_log.append(x0_buf.clone().cpu())
            if step_callback is not None:
                step_callback(i, x)
        self.last_run["vae_encodes"] = len(ledger.strip_cache)
        self.last_run["plan_host_ms_by_part"] = {k: round(v, 2) for k, v in ledger.host_ms.items()}
        return x, image_log


def _grid(imgs):
    """make_grid(imgs, nrow=8, padding=2) equivalent for the small logging grids of ed:1099-1124."""
    n, c, h, w = imgs.shape
    cols = min(8, n)
    rows = (n + cols - 1) // cols
    out = imgs.new_zeros(c, rows * (h + 2) + 2, cols * (w + 2) + 2)
    for k in range(n):
        r, q = divmod(k, cols)
        out[:, 2 + r * (h + 2):2 + r * (h + 2) + h, 2 + q * (w + 2):2 + q * (w + 2) + w] = imgs[k]
    return out


def _to_pil(img):
    """torchvision ToPILImage for a float CHW tensor in [0,1]: mul(255).byte() (ed:1125)."""
    from PIL import Image
    arr = img.detach().float().cpu().mul(255).byte().permute(1, 2, 0).numpy()
    return Image.fromarray(arr[:, :, 0] if arr.shape[2] == 1 else arr)
