"""TEST INFRASTRUCTURE - CPU oracle for the ElasticDiffusion global/local patched denoising loop.

NOT product code.  Only tests/, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference` leg
may import this module; the product package never does (it fails loudly without its CUDA library instead).

This is a *restatement* (plain torch on CPU, no custom kernels) of the reference algorithm
`ElasticDiffusion.generate_image` and everything it calls - `/root/reference/elastic_diffusion.py` ("ed:N" below).
It is written in table / closed-form style rather than as the reference's chain of tensor ops, but keeps the
reference's floating-point operation order and its exact sequence of RNG draws, so on CPU it
reproduces the reference bit for bit.

PINNING: the reference ships no tests and no golden vectors.  This oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF, run unmodified in the build container through `oracle/ref_shim.py`:
`tests/test_oracle_vs_reference.py` (live, skipped where /root/reference is absent) and the committed fixtures under
`tests/golden/` written by `scripts/make_golden.py`.  The DDIM arithmetic lives in un-vendored `diffusers==0.21.4`
and is restated in `oracle/ddim_restated.py` ("parity unpinned" against diffusers itself, see that header).

The oracle is device-agnostic torch: `device="cuda"` replays the same eager op sequence on the GPU (device Philox
RNG), which is what the reference's own PyTorch path does there; `device="cpu"` is the CPU baseline.
"""
from __future__ import annotations

import hashlib
import math
from fractions import Fraction

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# host-side geometry (integers only)
# ----------------------------------------------------------------------------------------------------------------

def view_grid(height_px, width_px, h_ws, w_ws, stride, scale=8):
    """Sliding-window table in latent units.  ed:198-229 (`get_views`)."""
    if height_px % scale or width_px % scale:
        raise TypeError(f"height {height_px} and width {width_px} must be divisible by {scale}")  # ed:200-201 raises a str
    H, W = height_px // scale, width_px // scale
    nh = math.ceil((H - h_ws) / stride) + 1 if stride else 1                       # ed:206
    nw = math.ceil((W - w_ws) / stride) + 1 if stride else 1                       # ed:207
    out = []
    for i in range(int(nh * nw)):
        h0 = int((i // nw) * stride)
        h1 = h0 + h_ws
        if h1 > H:                                                                 # ed:214-217 shift last window inside
            h0, h1 = max(0, h0 - (h1 - H)), H
        w0 = int((i % nw) * stride)
        w1 = w0 + w_ws
        if w1 > W:                                                                 # ed:222-225
            w0, w1 = max(0, w0 - (w1 - W)), W
        out.append((h0, h1, w0, w1))
    return out


def _context_1d(lo, hi, n, size):
    """Context extents (before, after) for one axis with stride S=1.  ed:718-744."""
    if lo - n < 0:
        before = lo - max(0, lo - n)                      # len(arange(max(0,lo-n), lo))
        after = max(0, min(size, hi + (2 * n - before)) - hi)
    else:
        after = max(0, min(size, hi + n) - hi)
        before = lo - max(0, lo - (2 * n - after))
    return before, after


def context_box(view, n, H, W):
    """`crop_with_context(..., S=1, n)` as a contiguous box.  ed:706-757 (the only call site is ed:838, S=1).

    Returns (r0, r1, c0, c1) of the crop and (n_t, n_b, n_l, n_r)."""
    h0, h1, w0, w1 = view
    n_t, n_b = _context_1d(h0, h1, n, H)
    n_l, n_r = _context_1d(w0, w1, n, W)
    return (h0 - n_t, h1 + n_b, w0 - n_l, w1 + n_r), (n_t, n_b, n_l, n_r)


def downsample_size(height_px, width_px, sd_version, scale=8):
    """ed:943-950 (`get_downsample_size`)."""
    factor = max(height_px, width_px) / (1024 if "XL" in sd_version else 512)
    factor = max(factor, 1)
    return int((height_px // factor) // scale), int((width_px // factor) // scale)


def even_rational(f, max_block=32):
    """ed:468-476 (`to_even_rational`)."""
    fr = Fraction(f).limit_denominator(max_block)
    if fr.numerator % 2 or fr.denominator % 2:
        fr = Fraction(f).limit_denominator(max_block // 2)
    if fr.numerator % 2 or fr.denominator % 2:
        return fr.numerator * 2, fr.denominator * 2
    return fr.numerator, fr.denominator


def keep_offsets(block_sz, n_remove):
    """ed:478-499 (`get_keep_blocks`) on `arange(block_sz)`: offsets kept inside a block + "masked block" ids."""
    pairs = n_remove // 2
    interval = block_sz // (pairs + 1)
    if interval % 2:
        interval += 1
    keep = [True] * block_sz
    marked = []
    for i in range(pairs):
        start = (i + 1) * interval - 1
        marked += [start - 1 - 2 * i, start - 2 * i]                  # ed:493
        for k in (start, start + 1):                                  # mask[start:start+2] = False (slice clips)
            if 0 <= k < block_sz:
                keep[k] = False
    return [o for o in range(block_sz) if keep[o]], marked


def axis_resample_table(n_in, n_out):
    """One axis of ed:568-611: which rows of the 2x-nearest-upsampled input survive (length 2*n_out) and the
    "masked block" list used by the mask restoration (ed:591-593)."""
    n_keep, block = even_rational(n_out / n_in)
    n_remove = block - n_keep
    n_blocks = (n_out * 2) // n_keep
    if n_blocks * block > n_in * 2:
        n_blocks -= 1
    span = n_blocks * block
    offsets, marked = keep_offsets(block, n_remove)
    src = [b + o for b in range(0, span, block) for o in offsets]
    src = [s for s in src if s < n_in * 2]                                        # ed:588
    remain = n_out * 2 - len(src)
    src = src + list(range(n_in * 2))[span:span + remain]                           # ed:612-613 (python slice semantics)
    special = [b + m for b in range(0, n_out * 2, n_keep) for m in marked]         # ed:591-592
    return src, special


def restore_groups(n_resized, special):
    """ed:446-465 (`restore_mask_shape`) as a table: output row -> tuple of resized rows that are OR-ed."""
    groups, i, j = [], 0, 0
    while i < n_resized:
        if j < len(special) and i == special[j]:
            groups.append((i,))
            groups.append((i + 1,))
            j += 2
        else:
            groups.append((i, i + 1))
        i += 2
    return groups


class ResampleTables:
    """Everything `random_nearest_downsample` caches per generate_image call (ed:584-609), as plain int tables."""

    def __init__(self, H, W, ds):
        self.H, self.W, self.ds = H, W, tuple(ds)
        rs, rspecial = axis_resample_table(H, ds[0])
        cs, cspecial = axis_resample_table(W, ds[1])
        self.row_src = torch.tensor(rs, dtype=torch.long) // 2      # latent row feeding each resized row
        self.col_src = torch.tensor(cs, dtype=torch.long) // 2
        self.row_groups = restore_groups(len(rs), rspecial)
        self.col_groups = restore_groups(len(cs), cspecial)
        self.rh, self.rw = len(rs), len(cs)                          # resized size (normally 2*ds)
        self.lh, self.lw = self.rh // 2, self.rw // 2                # low-res size actually produced


# ----------------------------------------------------------------------------------------------------------------
# RNG-bearing pieces (draw order is parity critical - SURVEY.md Appendix B)
# ----------------------------------------------------------------------------------------------------------------

def seed_all(seed, device, seed_np=True):
    """ed:165-171."""
    torch.manual_seed(seed)
    if torch.device(device).type == "cuda":
        torch.cuda.manual_seed(seed)
    if seed_np:
        np.random.seed(seed)


def md5_seed(s, num_bytes=4):
    """ed:321-324."""
    return int(hashlib.md5(s.encode()).hexdigest()[:num_bytes * 2], 16)


def draw_cell_indices(n_cells, exclude, hi=4, max_iteration=50):
    """ed:502-520 - CPU draws; `exclude` is (n_cells, hi) bool (any device) or None."""
    idx = torch.randint(0, hi, (n_cells,))
    if exclude is not None:
        rows = torch.arange(n_cells)
        bad = exclude[rows, idx]
        m = int(bad.sum())
        while m > 0 and max_iteration > 0:
            idx[bad.cpu()] = torch.randint(0, hi, (m,))
            bad = exclude[rows, idx]
            m = int(bad.sum())
            max_iteration -= 1
        bad = exclude[rows, idx]
        m = int(bad.sum())
        if m > 0:
            idx[bad.cpu()] = torch.randint(0, hi, (m,))
    return idx


def mix_with_previous(idx, prev, drop_p, device):
    """ed:540-544: keep a freshly drawn index where randint(0,101) > 100*drop_p, else the previous one."""
    drop = torch.randint(0, 101, (idx.numel(),), device=device)
    drop[drop <= (100 * drop_p)] = 0
    drop[drop >= (100 * drop_p)] = 1
    return idx * drop + prev * (1 - drop)


def pick_and_mask(latent, tabs: ResampleTables, idx):
    """Value gather + sampled-position mask.  ed:565, 612-613 (resize), ed:532-556 (2x2 pick), ed:622-628 (mask)."""
    lh, lw = tabs.lh, tabs.lw
    idx2 = idx.reshape(lh, lw)
    rr = torch.arange(lh, device=latent.device)[:, None] * 2 + idx2 // 2           # row in the resized grid
    cc = torch.arange(lw, device=latent.device)[None, :] * 2 + idx2 % 2
    src_r = tabs.row_src.to(latent.device)[rr]
    src_c = tabs.col_src.to(latent.device)[cc]
    low = latent[:, :, src_r, src_c]
    resized_mask = torch.zeros(tabs.rh, tabs.rw, dtype=torch.bool, device=latent.device)
    resized_mask[rr.reshape(-1), cc.reshape(-1)] = True
    rows = [resized_mask[list(g)].any(dim=0) for g in tabs.row_groups]
    m = torch.stack(rows, dim=0)
    cols = [m[:, list(g)].any(dim=1) for g in tabs.col_groups]
    m = torch.stack(cols, dim=1)
    full = torch.zeros(max(tabs.H, m.shape[0]), max(tabs.W, m.shape[1]), dtype=torch.bool, device=latent.device)
    full[:m.shape[0], :m.shape[1]] = m
    return low, full


def nearest_resize(x, size):
    """ed:869-883 with bottom=right=False (the only way it is called, ed:1071)."""
    return F.interpolate(x, size=size, mode="nearest")


class Models:
    """The injected dense modules + the few attributes the hot path reads from `self` (ed:111-163)."""

    def __init__(self, unet, vae, scheduler, text_fn, sd_version, device="cpu", view_batch_size=1,
                 patch_size=None, projection_dim=None, dtype=torch.float32, controlnet=None):
        self.unet, self.vae, self.scheduler, self.text_fn = unet, vae, scheduler, text_fn
        self.controlnet = controlnet          # ControlNet twin: elastic_diffusion_w_controlnet.py ("cn:N")
        self.sd_version, self.device, self.view_batch_size = sd_version, torch.device(device), view_batch_size
        self.dtype = dtype
        self.scale = 2 ** (len(vae.config.block_out_channels) - 1)                     # ed:156
        ws = patch_size if patch_size is not None else unet.config.sample_size // 2     # ed:159-163
        self.view_config = {"window_size": ws, "stride": ws, "context_size": unet.config.sample_size - ws}
        self.projection_dim = projection_dim
        self.default_size = None


def background_strip(m: Models, size, t, tag):
    """ed:327-364 (`make_denoised_background`): VAE-encoded flat random colour, noised to level t.

    Zero-size strips return before touching any generator (ed:332-333)."""
    h, w = size
    if h == 0 or w == 0:
        return torch.zeros(1, 4, h, w, device=m.device)
    with torch.autocast("cuda", enabled=False):
        seed_all(md5_seed(f"{tag}_{h}_{w}_{t}"), m.device, seed_np=False)              # ed:331,335
        colour = torch.rand(1, 3, device=m.device)[:, :, None, None].repeat(1, 1, h * m.scale, w * m.scale)
        z = m.vae.encode(colour).latent_dist.sample() * m.vae.config.scaling_factor     # ed:350
        noise = torch.randn_like(z)                                                     # ed:356
        z_t = m.scheduler.add_noise(z, noise, t.long())                                 # ed:358
        seed_all(int(np.random.randint(100000)), m.device, seed_np=False)               # ed:359
    return z_t


def unet_call(m: Models, x, t, text, pooled, cond=None, cond_scale=1.0):
    """ed:393-432 (`unet_step`): pad to the native size with background strips, UNet, crop.
    ControlNet twin cn:434-524: the condition image is zero-padded by pad*vae_scale (cn:457-461), a ControlNet forward on
    the padded latent yields residuals that are fed to the UNet (cn:482-496 / 506-518)."""
    native = 128 if m.sd_version.startswith("XL") else 64                               # ed:398-400
    x = m.scheduler.scale_model_input(x, t)
    hp, wp = max(native - x.shape[-2], 0), max(native - x.shape[-1], 0)
    lp, rp, tp, bp = wp // 2, wp - wp // 2, hp // 2, hp - hp // 2                        # ed:406
    xin = x
    if hp > 0 or wp > 0:
        B = x.shape[0]
        # width first (dim 3), then height (dim 2) over the already widened tensor - ed:372-389
        for dim, (before, after) in ((3, (lp, rp)), (2, (tp, bp))):
            shp = list(xin.shape)
            sb = (shp[2], before) if dim == 3 else (before, shp[3])
            sa = (shp[2], after) if dim == 3 else (after, shp[3])
            s1 = background_strip(m, sb, t, f"{dim}_1").repeat(B, 1, 1, 1).to(x)
            s2 = background_strip(m, sa, t, f"{dim}_2").repeat(B, 1, 1, 1).to(x)
            xin = torch.cat([s1, xin, s2], dim=dim)
        if cond is not None:
            sc = m.scale
            cond = F.pad(cond, (lp * sc, rp * sc, tp * sc, bp * sc))                     # cn:457-461
    extra = {}

    def control(added=None):
        if cond is None:
            return {}
        kw = {} if added is None else {"added_cond_kwargs": added}
        down, mid = m.controlnet(xin, t, encoder_hidden_states=text, controlnet_cond=cond[:xin.shape[0]],
                                 conditioning_scale=cond_scale, guess_mode=False, return_dict=False, **kw)   # cn:482-491
        return {"down_block_additional_residuals": down, "mid_block_additional_residual": mid}
    if m.sd_version.startswith("XL"):
        ids = list(m.default_size + (0, 0) + m.default_size)                            # ed:233, 414
        n_expected = m.unet.add_embedding.linear_1.in_features
        n_passed = m.unet.config.addition_time_embed_dim * len(ids) + m.projection_dim
        if n_expected != n_passed:
            raise ValueError(f"Model expects an added time embedding vector of length {n_expected}, but a vector of "
                             f"{n_passed} was created.")
        time_ids = torch.tensor([ids], dtype=text.dtype).to(m.device).repeat(xin.shape[0], 1)
        added = {"text_embeds": pooled, "time_ids": time_ids}
        out = m.unet(xin, t, encoder_hidden_states=text, added_cond_kwargs=added, **control(added))["sample"]
    else:
        out = m.unet(xin, t, encoder_hidden_states=text, **control())["sample"]
    if hp > 0 or wp > 0:
        out = out[:, :, tp:out.shape[-2] - bp, lp:out.shape[-1] - rp]                    # ed:429-430
    return out


def global_direction(m: Models, latent, t, text, pooled, tabs: ResampleTables, resampling_steps, drop_p, trace=None,
                     cimg=None, cond_scale=1.0):
    """ed:650-690 (`approximate_latent_direction_w_resampling`); the ControlNet twin threads the (CFG-doubled) condition
    image through to unet_step (cn:527-537, 744-781)."""
    target = torch.full_like(latent, float("nan")).half()                               # ed:655 (fp16 on purpose)
    exclude, prev = None, None
    info = {"init_downsampled_latent": None}
    n_cells = tabs.lh * tabs.lw
    for k in range(resampling_steps + 1):
        if k == 0:                                                                      # nearest & fix_initial
            idx = torch.zeros(n_cells, device=latent.device, dtype=torch.long)           # ed:536
        else:
            idx = draw_cell_indices(n_cells, exclude).to(latent.device)                  # ed:538
        if prev is not None:
            idx = mix_with_previous(idx, prev, drop_p, latent.device)                    # ed:540-544
        low, mask = pick_and_mask(latent, tabs, idx)
        prev = idx
        if exclude is None:
            exclude = torch.zeros((n_cells, 4), dtype=torch.bool, device=latent.device)  # ed:674
        exclude[torch.arange(n_cells), prev] = True                                      # ed:675
        if info["init_downsampled_latent"] is None:
            info["init_downsampled_latent"] = low.clone()
        both = unet_call(m, torch.cat([low] * 2), t, text, pooled, cimg, cond_scale)     # ed:436-438
        uncond, cond = both.chunk(2)
        direction = cond - uncond                                                        # ed:440
        up = nearest_resize(direction, (target.size(2), target.size(3)))                 # ed:636
        target = torch.where(mask, up, target)                                           # ed:637
        if k == resampling_steps:                                                        # ed:639-644
            target = torch.where(torch.isnan(target), up, target)
        if trace is not None:
            trace.setdefault("idx", []).append(idx.clone())
            trace.setdefault("mask", []).append(mask.clone())
    info["downsampled_latent"] = low
    info["scores"] = {"uncond_score": uncond, "cond_score": cond}
    info["downsampled_direction"] = nearest_resize(target, tabs.ds)                       # ed:688
    return target, info


def local_uncond(m: Models, latent, t, uncond_text, uncond_pooled, cond=None, cond_scale=1.0):
    """ed:814-864 (`compute_local_uncond_signal`): view gather -> UNet -> first-writer-wins scatter.
    ControlNet twin cn:917-983: condition_image[0:1] is nearest-upsampled to the full pixel size (cn:933) and cropped per
    view with the window coordinates x 8 and context*8//2 (cn:946-949, the factor 8 is hard-coded)."""
    H, W = latent.shape[-2:]
    cond_up = None
    if cond is not None:
        cond_up = nearest_resize(cond[0:1], (H * m.scale, W * m.scale))
    vc = m.view_config
    h_ws = H if vc["window_size"] + vc["context_size"] >= H else vc["window_size"]        # ed:820-825
    w_ws = W if vc["window_size"] + vc["context_size"] >= W else vc["window_size"]
    views = view_grid(H * m.scale, W * m.scale, h_ws, w_ws, vc["stride"], m.scale)
    out = torch.zeros_like(latent)
    n = vc["context_size"] // 2
    for s in range(0, len(views), m.view_batch_size):
        chunk = views[s:s + m.view_batch_size]
        boxes = [context_box(v, n, H, W) for v in chunk]
        crops = torch.cat([latent[:, :, r0:r1, c0:c1] for (r0, r1, c0, c1), _ in boxes])
        cviews = None
        if cond_up is not None:
            cb = [context_box(tuple(8 * q for q in v), (vc["context_size"] * 8) // 2, H * 8, W * 8)[0] for v in chunk]
            cviews = torch.cat([cond_up[:, :, r0:r1, c0:c1] for (r0, r1, c0, c1) in cb])
        pred = unet_call(m, crops, t, torch.cat([uncond_text] * len(chunk)), torch.cat([uncond_pooled] * len(chunk)),
                         cviews, cond_scale)
        for (h0, h1, w0, w1), (_, (n_t, n_b, n_l, n_r)), p in zip(chunk, boxes, pred.chunk(len(chunk))):
            centre = p[:, :, n_t:p.shape[-2] - n_b, n_l:p.shape[-1] - n_r]
            dst = out[:, :, h0:h1, w0:w1]
            empty = ~(dst != 0)                                                            # ed:859
            dst[empty] = centre[empty].to(out.dtype)                                       # ed:860-861
    return out


def renoise(m: Models, x, t_next):
    """ed:692-704 (`undo_step`): num_train/num_inference forward-diffusion steps from t_next."""
    n = m.scheduler.config.num_train_timesteps // m.scheduler.num_inference_steps
    for i in range(n):
        beta = m.scheduler.betas[t_next + i]
        noise = torch.randn(x.shape, device=x.device, dtype=x.dtype)
        x = (1 - beta) ** 0.5 * x + beta ** 0.5 * noise
    return x


def rrg_gradient(m: Models, t, x0_full, low_latent, low_uncond, low_direction, cfg, weight):
    """ed:886-940 (`reduced_resolution_guidance`, the `donwsampled_scores` branch ed:909-916)."""
    eps = low_uncond + cfg * low_direction                                                 # ed:918
    ref_x0 = m.scheduler.step(eps, t, low_latent)["pred_original_sample"]                  # ed:920-921
    ref_up = nearest_resize(ref_x0, x0_full.shape[-2:])                                    # ed:922
    grads = []
    for j in range(len(x0_full)):                                                          # ed:927-936
        with torch.enable_grad():
            probe = x0_full[j:j + 1].clone().detach().requires_grad_(True)
            loss = weight * F.mse_loss(ref_up[j:j + 1], probe)
            loss.backward()
            grads.append(probe.grad.clone() * -1.0)
    return torch.cat(grads), ref_x0


class CosineWeight:
    """ed:96-107 (`CosineScheduler`)."""

    def __init__(self, steps, cosine_scale, factor=0.01):
        self.steps, self.cosine_scale, self.factor = steps, cosine_scale, factor

    def __call__(self, i):
        if i >= self.steps:
            return 0
        return self.factor * ((0.5 * (1 + np.cos(np.pi * i / self.steps))) ** self.cosine_scale)


class LinearWeight:
    """ed:73-82."""

    def __init__(self, steps, start_val, stop_val):
        self.steps, self.start_val, self.stop_val = steps, start_val, stop_val

    def __call__(self, i):
        return self.stop_val if i >= self.steps else self.start_val + (self.stop_val - self.start_val) / self.steps * i


class ConstWeight(LinearWeight):
    """ed:85-94."""

    def __call__(self, i):
        return self.stop_val if i >= self.steps else self.start_val


@torch.no_grad()
def denoise(m: Models, prompts, negative_prompts="", height=768, width=768, num_inference_steps=50,
            guidance_scale=10.0, resampling_steps=20, new_p=0.3, rrg_stop_t=0.2, rrg_init_weight=1000,
            rrg_scheduler="cosine", cosine_scale=3.0, repaint_sampling=True, trace=None, step_callback=None,
            condition_image=None, controlnet_conditioning_scale=1.0):
    """The loop of ed:952-1078; returns the final latent (what the reference hands to the VAE at ed:1121).
    With `condition_image` ((1,3,ds_h*8,ds_w*8) float tensor in [0,1]) it is the ControlNet twin's loop cn:1119-1322."""
    ds = downsample_size(height, width, m.sd_version, m.scale)                             # ed:968
    m.default_size = (4 * height, 4 * width)                                               # ed:969
    steps_rrg = num_inference_steps - int(num_inference_steps * rrg_stop_t)
    if rrg_scheduler == "cosine":                                                          # ed:972-979
        w_of = CosineWeight(steps_rrg, cosine_scale, rrg_init_weight)
    elif rrg_scheduler == "linear":
        w_of = LinearWeight(steps_rrg, rrg_init_weight, 0)
    else:
        w_of = ConstWeight(steps_rrg, rrg_init_weight, 0)
    prompts = [prompts] if isinstance(prompts, str) else prompts
    negative_prompts = [negative_prompts] * len(prompts) if isinstance(negative_prompts, str) else negative_prompts
    un_text, un_pool = m.text_fn(negative_prompts)                                         # ed:992-993
    co_text, co_pool = m.text_fn(prompts)
    text, pooled = torch.cat([un_text, co_text]), torch.cat([un_pool, co_pool], dim=0)
    x = torch.randn((len(prompts), m.unet.config.in_channels, height // m.scale, width // m.scale),
                    device=m.device, dtype=m.dtype)                                        # ed:998
    m.scheduler.set_timesteps(num_inference_steps)
    ts = m.scheduler.timesteps
    tabs = ResampleTables(x.shape[-2], x.shape[-1], ds)
    cond, cs = None, controlnet_conditioning_scale
    if condition_image is not None:                                                        # prepare_image, cn:1005-1033, 1183-1193
        img = condition_image.to(dtype=torch.float32)
        assert img.shape[-2:] == (ds[0] * m.scale, ds[1] * m.scale), "condition image must be (ds_h*8, ds_w*8)"
        img = img.repeat_interleave(1, dim=0).to(device=m.device, dtype=m.controlnet.dtype)
        cond = torch.cat([img] * 2).to(m.device)                                           # do_classifier_free_guidance
    with torch.autocast("cuda", enabled=(m.device.type == "cuda")):                        # ed:1012
        for i, t in enumerate(ts):
            d, info = global_direction(m, x, t, text, pooled, tabs, resampling_steps, 1 - new_p, trace, cond, cs)
            u = local_uncond(m, x, t, un_text, un_pool, cond, cs)
            out = m.scheduler.step(u + guidance_scale * d, t, x)                            # ed:1031-1033
            x0, nxt, cfg = out["pred_original_sample"], out["prev_sample"], guidance_scale
            if repaint_sampling and resampling_steps > 0 and i < len(ts) - 1:               # ed:1038
                x = renoise(m, nxt, ts[i + 1])
                cfg = guidance_scale / 3
                d, info = global_direction(m, x, t, text, pooled, tabs, 0, 1 - new_p, None, cond, cs)
                u = local_uncond(m, x, t, un_text, un_pool, cond, cs)
                out = m.scheduler.step(u + cfg * d, t, x)
                x0, nxt = out["pred_original_sample"], out["prev_sample"]
            cascade = torch.zeros_like(nxt)
            w = w_of(i)
            if w > 10:                                                                      # ed:1062
                cascade, _ = rrg_gradient(m, t, x0, info["downsampled_latent"], info["scores"]["uncond_score"],
                                          info["downsampled_direction"], cfg, w)
            x = nxt + cascade                                                               # ed:1078
            if step_callback is not None:
                step_callback(i, x, x0)
    return x


def decode_plain(m: Models, z):
    """ed:267-272."""
    z = z.to(next(iter(m.vae.post_quant_conv.parameters())).dtype) / m.vae.config.scaling_factor
    return (m.vae.decode(z).sample / 2 + 0.5).clamp(0, 1)


def decode_tiled(m: Models, z, low_vram=False):
    """ed:275-310: zero-pad, decode (core+2*pad)^2 tiles, accumulate centres, divide by the hit count."""
    Hpx, Wpx = z.shape[2] * m.scale, z.shape[3] * m.scale
    core = m.unet.config.sample_size // 4
    stride, pad = core, m.unet.config.sample_size // m.scale * 3
    if low_vram:
        stride, pad = core // 2, core
    tiles = view_grid(Hpx, Wpx, core, core, stride, m.scale)
    zp = F.pad(z, (pad, pad, pad, pad), "constant", 0)
    img = torch.zeros(z.size(0), 3, Hpx, Wpx, device=z.device)
    cnt = torch.zeros_like(img)
    for (h0, h1, w0, w1) in tiles:
        patch = decode_plain(m, zp[:, :, h0:h1 + 2 * pad, w0:w1 + 2 * pad])
        s = m.scale
        img[:, :, h0 * s:h1 * s, w0 * s:w1 * s] += patch[:, :, pad * s:patch.size(2) - pad * s, pad * s:patch.size(3) - pad * s]
        cnt[:, :, h0 * s:h1 * s, w0 * s:w1 * s] += 1
    return img / cnt
