"""TEST INFRASTRUCTURE - torch emulation of the C-ABI kernels' contracts (include/elastic_b200.h).

Not product code; only tests/ and smoke() import it.  Each `spec_*` function computes, with plain torch ops on any
device, exactly what the corresponding `ed_*` kernel must write given the same plan tables, in the reference's
floating-point operation order (so CUDA results are compared bit for bit).  `denoise_wave_form` drives the product's
own host logic (geometry tables + RngLedger) with these emulations instead of the CUDA library, which lets the CPU
test-suite check the whole wave-batched dataflow against the goldens generated from the unmodified reference.
"""
from __future__ import annotations

import numpy as np
import torch


def _t(geo, name, dev):
    return torch.tensor(geo.tables[name], dtype=torch.long, device=dev)


def spec_pick_gather(geo, R1, latent, idx, strips):
    """ed_random_pick_gather: (2*B*R1, C, dH, dW) canvases of the global passes."""
    dev = latent.device
    B, C = geo.B, geo.C
    lp, rp, tp, bp = geo.g_pad
    row_src, col_src = _t(geo, "row_src", dev), _t(geo, "col_src", dev)
    out = torch.empty(2 * B * R1, C, geo.native, geo.native, device=dev, dtype=latent.dtype)
    r = torch.arange(geo.lh, device=dev)[:, None]
    c = torch.arange(geo.lw, device=dev)[None, :]
    for k in range(R1):
        pick = idx[k].to(dev).long().view(geo.lh, geo.lw)
        low = latent[:, :, row_src[2 * r + pick // 2], col_src[2 * c + pick % 2]]
        canv = low
        left, right, top, bottom = strips
        if lp or rp:
            parts = ([left.to(dev).expand(B, -1, -1, -1)] if lp else []) + [canv] + \
                    ([right.to(dev).expand(B, -1, -1, -1)] if rp else [])
            canv = torch.cat(parts, dim=3)
        if tp or bp:
            parts = ([top.to(dev).expand(B, -1, -1, -1)] if tp else []) + [canv] + \
                    ([bottom.to(dev).expand(B, -1, -1, -1)] if bp else [])
            canv = torch.cat(parts, dim=2)
        out[(2 * k) * B:(2 * k + 1) * B] = canv
        out[(2 * k + 1) * B:(2 * k + 2) * B] = canv
    return out


def spec_gather_views(geo, latent, strips=(None, None, None, None)):
    """ed_gather_views (+ ed_pad_views): (nv*B, C, dH, dW) canvases of the local views."""
    dev = latent.device
    B, C = geo.B, geo.C
    vlp, vrp, vtp, vbp = geo.v_pad
    vt = geo.tables["views"]
    out = torch.zeros(geo.nv * B, C, geo.native, geo.native, device=dev, dtype=latent.dtype)
    for v in range(geo.nv):
        r0, c0 = vt[v * 8 + 4], vt[v * 8 + 5]
        canv = latent[:, :, r0:r0 + geo.vh, c0:c0 + geo.vw]
        left, right, top, bottom = strips
        if vlp or vrp:
            canv = torch.cat(([left.to(dev).expand(B, -1, -1, -1)] if vlp else []) + [canv] +
                             ([right.to(dev).expand(B, -1, -1, -1)] if vrp else []), dim=3)
        if vtp or vbp:
            canv = torch.cat(([top.to(dev).expand(B, -1, -1, -1)] if vtp else []) + [canv] +
                             ([bottom.to(dev).expand(B, -1, -1, -1)] if vbp else []), dim=2)
        out[v * B:(v + 1) * B] = canv
    return out


def owner_map(geo, R1, idx, dev):
    """(H, W) long: resampling iteration whose fill is the last to touch each pixel (kernel: owner_iteration)."""
    H, W, lh, lw = geo.H, geo.W, geo.lh, geo.lw
    rlo, rn = _t(geo, "mrow_lo", dev), _t(geo, "mrow_n", dev)
    clo, cn = _t(geo, "mcol_lo", dev), _t(geo, "mcol_n", dev)
    ry = torch.arange(2 * lh, device=dev)[:, None]
    rx = torch.arange(2 * lw, device=dev)[None, :]
    code = ((ry & 1) << 1) | (rx & 1)
    owner = torch.full((H, W), -1, dtype=torch.long, device=dev)
    for k in range(R1):
        pick = idx[k].to(dev).long().view(lh, lw)
        M = pick[ry >> 1, rx >> 1] == code                       # sampled positions in the resized grid
        restored = torch.zeros(H, W, dtype=torch.bool, device=dev)
        for a in range(2):
            for e in range(2):
                rr = (rlo + a).clamp(max=2 * lh - 1)
                cc = (clo + e).clamp(max=2 * lw - 1)
                ok = (rn > a)[:, None] & (cn > e)[None, :]
                restored |= M[rr][:, cc] & ok
        owner = torch.where(restored, torch.full_like(owner, k), owner)
    return torch.where(owner < 0, torch.full_like(owner, R1 - 1), owner)


def _direction_full(geo, R1, unet_out, idx, fp16sem):
    dev = unet_out.device
    B = geo.B
    lp, rp, tp, bp = geo.g_pad
    up_r, up_c = _t(geo, "up_row", dev), _t(geo, "up_col", dev)
    own = owner_map(geo, R1, idx, dev)
    dirs = []
    for k in range(R1):
        un = unet_out[(2 * k) * B:(2 * k + 1) * B, :, tp:tp + geo.lh, lp:lp + geo.lw].float()
        co = unet_out[(2 * k + 1) * B:(2 * k + 2) * B, :, tp:tp + geo.lh, lp:lp + geo.lw].float()
        d = co - un
        if fp16sem:
            d = d.half().float()
        dirs.append(d[:, :, up_r][:, :, :, up_c])
    stack = torch.stack(dirs)                                     # (R1, B, C, H, W)
    sel = own[None, None, None].expand(1, *stack.shape[1:])
    return torch.gather(stack, 0, sel)[0]


def _local_uncond_full(geo, R1, unet_out):
    B = geo.B
    vlp, vrp, vtp, vbp = geo.v_pad
    vt = geo.tables["views"]
    base = 2 * B * R1
    out = torch.zeros(B, geo.C, geo.H, geo.W, device=unet_out.device, dtype=torch.float32)
    for v in range(geo.nv):
        h0, h1, w0, w1, r0, c0, n_t, n_l = vt[v * 8:v * 8 + 8]
        p = unet_out[base + v * B:base + (v + 1) * B].float()
        centre = p[:, :, vtp + n_t:vtp + n_t + (h1 - h0), vlp + n_l:vlp + n_l + (w1 - w0)]
        dst = out[:, :, h0:h1, w0:w1]
        empty = ~(dst != 0)
        dst[empty] = centre[empty]
    return out


def spec_epilogue(geo, prm, latent, unet_out, idx, noise=None):
    """ed_wave_epilogue.  `prm`: dict with the ed_step_params_t fields (flags as an int).
    Returns (out_latent, x0)."""
    dev = latent.device
    # 0-dim tensors ON the data's device: torch's CUDA division by a *CPU* scalar multiplies by the reciprocal instead
    # of dividing (1 ulp off the CPU result); a device operand takes the IEEE division kernel on every backend.
    f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
    R1, flags = prm["R1"], prm["flags"]
    fp16sem = bool(flags & 4)
    g, sb, sa = f32(prm["guidance"]), f32(prm["sqrt_beta_t"]), f32(prm["sqrt_alpha_t"])
    sap, sd = f32(prm["sqrt_alpha_prev"]), f32(prm["sqrt_dir"])
    u = _local_uncond_full(geo, R1, unet_out)
    d = _direction_full(geo, R1, unet_out, idx, fp16sem)
    gd = g * d
    if fp16sem:
        gd = gd.half().float()
    eps = u + gd
    x0 = (latent - sb * eps) / sa
    res = sap * x0 + sd * eps
    if flags & 2:   # RRG
        B = geo.B
        lp, rp, tp, bp = geo.g_pad
        kl = R1 - 1
        row_src, col_src = _t(geo, "row_src", dev), _t(geo, "col_src", dev)
        up_r, up_c = _t(geo, "up_row", dev), _t(geo, "up_col", dev)
        dn_r, dn_c = _t(geo, "down_row", dev), _t(geo, "down_col", dev)
        pick = idx[kl].to(dev).long().view(geo.lh, geo.lw)
        r = torch.arange(geo.lh, device=dev)[:, None]
        c = torch.arange(geo.lw, device=dev)[None, :]
        xl = latent[:, :, row_src[2 * r + pick // 2], col_src[2 * c + pick % 2]]
        ul = unet_out[(2 * kl) * B:(2 * kl + 1) * B, :, tp:tp + geo.lh, lp:lp + geo.lw].float()
        dl = d[:, :, dn_r][:, :, :, dn_c]
        gl = g * dl
        if fp16sem:
            gl = gl.half().float()
            el = (ul + gl).half().float()
            t1 = (sb * el).half().float()
        else:
            el = ul + gl
            t1 = sb * el
        ref = ((xl - t1) / sa)[:, :, up_r][:, :, :, up_c]
        grad = (f32(prm["rrg_norm"]) * (x0 - ref)) * f32(prm["rrg_weight"])
        res = res + (-grad)
    if flags & 1:   # re-noise
        for k in range(prm["n_renoise"]):
            res = f32(prm["renoise_a"][k]) * res + f32(prm["renoise_b"][k]) * noise[k]
    return res, x0


def spec_gather_cond(geo, R1, cond, scale, row_map, col_map, origin):
    """ed_gather_cond: (2*B*R1 + nv*B, CH, dH*scale, dW*scale) condition batch (cn:457-461, 933-949)."""
    import torch.nn.functional as F
    dev = cond.device
    B = geo.B
    lp, rp, tp, bp = geo.g_pad
    vlp, vrp, vtp, vbp = geo.v_pad
    s = scale
    padded = F.pad(cond, (lp * s, rp * s, tp * s, bp * s))                       # (2, CH, dH*s, dW*s)
    glob = torch.cat([padded[sidx:sidx + 1].expand(B, -1, -1, -1) for _ in range(R1) for sidx in (0, 1)])
    rm, cm = torch.tensor(row_map, device=dev), torch.tensor(col_map, device=dev)
    up = cond[0:1][:, :, rm][:, :, :, cm]                                         # nearest upsample to full pixel size
    views = []
    for v in range(geo.nv):
        r0, c0 = origin[2 * v], origin[2 * v + 1]
        box = up[:, :, r0:r0 + geo.vh * s, c0:c0 + geo.vw * s]
        box = F.pad(box, (vlp * s, vrp * s, vtp * s, vbp * s))
        views.append(box.expand(B, -1, -1, -1))
    return torch.cat([glob] + views)


def spec_tile_gather(latent, tiles, core, pad):
    import torch.nn.functional as F
    zp = F.pad(latent, (pad, pad, pad, pad), "constant", 0)
    return torch.cat([zp[:, :, h0:h1 + 2 * pad, w0:w1 + 2 * pad] for (h0, h1, w0, w1) in tiles])


def spec_tile_blend(patches, tiles, B, H, W, core, pad, scale):
    s = scale
    img = torch.zeros(B, patches.shape[1], H * s, W * s, device=patches.device)
    cnt = torch.zeros_like(img)
    for j, (h0, h1, w0, w1) in enumerate(tiles):
        p = (patches[j * B:(j + 1) * B].float() / 2 + 0.5).clamp(0, 1)
        img[:, :, h0 * s:h1 * s, w0 * s:w1 * s] += p[:, :, pad * s:p.size(2) - pad * s, pad * s:p.size(3) - pad * s]
        cnt[:, :, h0 * s:h1 * s, w0 * s:w1 * s] += 1
    return img / cnt


# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def denoise_wave_form(ed, prompts, negative_prompts="", height=768, width=768, num_inference_steps=50,
                      guidance_scale=10.0, resampling_steps=20, new_p=0.3, rrg_stop_t=0.2, rrg_init_weight=1000,
                      cosine_scale=3.0, repaint_sampling=True, trace=None, condition_image=None,
                      controlnet_conditioning_scale=1.0, rrg_scherduler_cls=None):
    """The product's wave-batched loop (pipeline.ElasticDiffusion.denoise) with the CUDA ops replaced by the spec
    emulations above; uses the product's geometry, RngLedger and DDIM scalar helpers unchanged."""
    import importlib
    pkg = importlib.import_module(type(ed).__module__.rsplit(".", 1)[0])
    pl = importlib.import_module(pkg.__name__ + ".pipeline")
    from_ddim = importlib.import_module(pkg.__name__ + ".ddim")
    sf = ed.vae_scale_factor
    ds = ed.get_downsample_size(height, width)
    ed.default_size = (4 * height, 4 * width)
    T = num_inference_steps
    if rrg_scherduler_cls in (None, pl.CosineScheduler):                                       # ed:972-979
        rrg_w = pl.CosineScheduler(steps=T - int(T * rrg_stop_t), cosine_scale=cosine_scale, factor=rrg_init_weight)
    else:
        rrg_w = rrg_scherduler_cls(steps=T - int(T * rrg_stop_t), start_val=rrg_init_weight, stop_val=0)
    prompts = [prompts] if isinstance(prompts, str) else prompts
    negative_prompts = [negative_prompts] * len(prompts) if isinstance(negative_prompts, str) else negative_prompts
    un_text, un_pool = ed.get_text_embeds(negative_prompts)
    co_text, co_pool = ed.get_text_embeds(prompts)
    B, C, H, W = len(prompts), ed.unet.config.in_channels, height // sf, width // sf
    is_xl = ed.sd_version.startswith("XL")
    nat = 128 if is_xl else 64
    vc = ed.view_config
    geo = pkg.geometry.build_geometry(B, C, H, W, nat, ds, vc["window_size"], vc["stride"], vc["context_size"])
    ledger = pl.RngLedger(ed, geo)
    x = ledger.initial_latent((B, C, H, W), ed.torch_dtype)
    ed.scheduler.set_timesteps(T)
    ts = ed.scheduler.timesteps
    if getattr(ed, "precompute_strips", False):     # product behaviour: all background strips before the loop, batched
        ed.last_run = dict(getattr(ed, "last_run", {}) or {})
        ed.last_run["vae_encode_calls"] = ledger.precompute_strips(ts)
    R, nv = resampling_steps, geo.nv
    n_re = ed.scheduler.config.num_train_timesteps // T
    text_pair, pool_pair = torch.cat([un_text, co_text]), torch.cat([un_pool, co_pool], dim=0)
    time_ids = ed._get_add_time_ids(ed.default_size, (0, 0), ed.default_size, dtype=text_pair.dtype) if is_xl else None
    rrg_norm = float(torch.tensor(2.0 / (C * H * W), dtype=torch.float64).to(torch.float32))

    conds = {}
    if condition_image is not None:
        prepared = torch.cat([condition_image.float()] * 2)
        rm, cm, org = pkg.geometry.cond_geometry(geo, sf, vc["window_size"], vc["context_size"])
        conds = {R1: spec_gather_cond(geo, R1, prepared, sf, rm, cm, org) for R1 in {resampling_steps + 1, 1}}

    def unet(canvas, t, R1):
        text = torch.cat([text_pair] * R1 + [un_text] * nv)
        pool = torch.cat([pool_pair] * R1 + [un_pool] * nv)
        kw = {}
        if is_xl:
            kw["added_cond_kwargs"] = {"text_embeds": pool, "time_ids": time_ids.to(canvas.device).repeat(len(canvas), 1)}
        res = {}
        if conds:
            down, mid = ed.controlnet(canvas, t, encoder_hidden_states=text, controlnet_cond=conds[R1],
                                      conditioning_scale=controlnet_conditioning_scale, guess_mode=False,
                                      return_dict=False, **kw)
            res = {"down_block_additional_residuals": down, "mid_block_additional_residual": mid}
        return ed.unet(canvas, t, encoder_hidden_states=text, **kw, **res)["sample"]

    def wave(x_in, t, idx, R1, sg, sv, prm, noise):
        canvas = torch.cat([spec_pick_gather(geo, R1, x_in, idx, sg), spec_gather_views(geo, x_in, sv)])
        out = unet(canvas, t, R1)
        return spec_epilogue(geo, prm, x_in, out, idx, noise)

    for i, t in enumerate(ts):
        last = i == len(ts) - 1
        repaint = bool(repaint_sampling and R > 0 and not last)
        w = rrg_w(i)
        sc = from_ddim.step_scalars(ed.scheduler, t)
        idx1, sg = ledger.global_pass(t, R, 1 - new_p)
        sv = ledger.local_pass(t, ed.view_batch_size)
        if trace is not None:
            trace.setdefault("idx", []).append(idx1.clone())
        prm = dict(guidance=guidance_scale, rrg_weight=float(w), rrg_norm=rrg_norm, R1=R + 1, flags=0, n_renoise=0, **sc)
        if repaint:
            noise = torch.empty(n_re, B, C, H, W)
            ledger.undo_noise(n_re, (B, C, H, W), noise)
            _, sg2 = ledger.global_pass(t, 0, 1 - new_p)
            sv2 = ledger.local_pass(t, ed.view_batch_size)
            a, b = from_ddim.renoise_scalars(ed.scheduler, ts[i + 1])
            prm.update(flags=1, n_renoise=n_re, renoise_a=a, renoise_b=b)
            x_mid, _ = wave(x, t, idx1, R + 1, sg, sv, prm, noise)
            prm2 = dict(guidance=guidance_scale / 3, rrg_weight=float(w), rrg_norm=rrg_norm, R1=1,
                        flags=2 if w > 10 else 0, n_renoise=0, **sc)
            x, _ = wave(x_mid, t, torch.zeros(1, geo.lh * geo.lw, dtype=torch.uint8), 1, sg2, sv2, prm2, None)
        else:
            prm["flags"] = 2 if w > 10 else 0
            x, _ = wave(x, t, idx1, R + 1, sg, sv, prm, None)
    return x
