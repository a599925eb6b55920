"""Host-side integer geometry of one `generate_image` call: everything the CUDA kernels index with.

Reference logic restated here ("ed:N" = /root/reference/elastic_diffusion.py line N):
  view grid                        get_views                       ed:198-229
  window collapse                  compute_local_uncond_signal     ed:820-825
  context boxes (S=1)              crop_with_context               ed:706-757
  rational resampling tables       random_nearest_downsample       ed:565-613 (+ to_even_rational ed:468-476,
                                                                    get_keep_blocks ed:478-499)
  mask restoration groups          restore_mask_shape              ed:446-465, 622-628
  nearest up / down index maps     nearest_interpolate             ed:869-883 (F.interpolate, mode='nearest')
  low-res size                     get_downsample_size             ed:943-950
  UNet pad split                   unet_step                       ed:398-406
  decode tiles                     tiled_decode                    ed:276-287

Everything is computed ONCE per call on the host (pure Python ints / small torch CPU tensors) and uploaded as int32
tables; the reference recomputes most of it with tensor ops on every denoise step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np
import torch
import torch.nn.functional as F


def sliding_windows(H, W, h_ws, w_ws, stride):
    """Windows (h0,h1,w0,w1) in latent units, row-major, last row/col shifted back inside (ed:206-227).
    Returns (windows, n_rows, n_cols)."""
    n_r = math.ceil((H - h_ws) / stride) + 1 if stride else 1
    n_c = math.ceil((W - w_ws) / stride) + 1 if stride else 1
    n_r, n_c = int(n_r), int(n_c)
    wins = []
    for r in range(n_r):
        h0 = r * stride
        h1 = h0 + h_ws
        if h1 > H:
            h0, h1 = max(0, h0 - (h1 - H)), H
        for c in range(n_c):
            w0 = c * stride
            w1 = w0 + w_ws
            if w1 > W:
                w0, w1 = max(0, w0 - (w1 - W)), W
            wins.append((h0, h1, w0, w1))
    return wins, n_r, n_c


def context_extent(lo, hi, n, size):
    """(before, after) context pixels along one axis for stride 1; the budget 2n is moved to the far side at a
    border (ed:718-744)."""
    if lo - n < 0:
        before = lo
        after = max(0, min(size, hi + 2 * n - before) - hi)
    else:
        after = max(0, min(size, hi + n) - hi)
        before = lo - max(0, lo - (2 * n - after))
    return before, after


def pad_split(native, size):
    """(before, after) background padding of one axis up to the UNet's native size (ed:405-406)."""
    p = max(native - size, 0)
    return p // 2, p - p // 2


def low_res_size(height_px, width_px, sd_version, scale):
    """ed:943-950."""
    factor = max(height_px, width_px) / (1024 if "XL" in sd_version else 512)
    factor = max(factor, 1)
    return int((height_px // factor) // scale), int((width_px // factor) // scale)


def _even_ratio(f, max_block=32):
    fr = Fraction(f).limit_denominator(max_block)
    if fr.numerator % 2 or fr.denominator % 2:
        fr = Fraction(f).limit_denominator(max_block // 2)
    if fr.numerator % 2 or fr.denominator % 2:
        return fr.numerator * 2, fr.denominator * 2
    return fr.numerator, fr.denominator


def resample_axis(n_in, n_out):
    """One axis of the "2x nearest upsample, then drop rows per block" resampler.

    Returns (src, lo, cnt):
      src[i]  latent row feeding row i of the resized grid (len = resized size, normally 2*n_out)
      lo/cnt  for every latent row y < n_in: the consecutive resized rows that are OR-ed into row y of the restored
              sampled-position mask (cnt = 0 for rows the restoration never produces: they stay False, ed:625-628)
    """
    keep, block = _even_ratio(n_out / n_in)
    drop_pairs = (block - keep) // 2
    n_blocks = (n_out * 2) // keep
    if n_blocks * block > n_in * 2:
        n_blocks -= 1
    span = n_blocks * block
    interval = block // (drop_pairs + 1)
    interval += interval % 2
    dropped, marked = set(), []
    for i in range(drop_pairs):
        s = (i + 1) * interval - 1
        dropped.update(k for k in (s, s + 1) if 0 <= k < block)
        marked += [s - 1 - 2 * i, s - 2 * i]
    offsets = [o for o in range(block) if o not in dropped]
    src2 = [b + o for b in range(0, span, block) for o in offsets if b + o < n_in * 2]
    src2 += list(range(n_in * 2))[span:span + (n_out * 2 - len(src2))]
    special = [b + m for b in range(0, n_out * 2, keep) for m in marked]
    # mask restoration: walk the resized rows in pairs; a pair that starts at the next "special" index keeps both
    # rows, every other pair collapses to one row (ed:446-465)
    groups, i, j = [], 0, 0
    while i < len(src2):
        if j < len(special) and i == special[j]:
            groups += [(i, 1), (i + 1, 1)]
            j += 2
        else:
            groups.append((i, 2))
        i += 2
    if len(groups) > n_in:
        raise ValueError(f"resampling {n_in}->{n_out}: restored mask has {len(groups)} rows > {n_in} "
                         "(the reference's torch.where would fail to broadcast here too)")
    lo = [g[0] for g in groups] + [0] * (n_in - len(groups))
    cnt = [min(g[1], len(src2) - g[0]) for g in groups] + [0] * (n_in - len(groups))
    return [s // 2 for s in src2], lo, cnt


def nearest_index(n_in, n_out):
    """Source index read by F.interpolate(mode='nearest') for each of n_out outputs - taken from torch itself so that
    float rounding of the scale matches (ed:876)."""
    ramp = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1)
    return F.interpolate(ramp, size=(n_out, 1), mode="nearest").view(-1).to(torch.int64).tolist()


def cover_ranges(starts, length, size):
    """For windows [s, s+length) with non-decreasing starts: per position p the first covering window and how many
    consecutive windows cover it."""
    st = np.asarray(starts, dtype=np.int64)
    p = np.arange(size, dtype=np.int64)
    # windows are sorted by start and share one length, so the windows covering p are the consecutive run
    # [first index with start + length > p, first index with start > p)
    lo = np.searchsorted(st + length, p, side="right")
    hi = np.searchsorted(st, p, side="right")
    cnt = np.maximum(hi - lo, 0)
    first = np.where(cnt > 0, lo, 0)
    return first.tolist(), cnt.tolist()


@dataclass
class WaveGeometry:
    B: int
    C: int
    H: int
    W: int
    native: int
    lh: int
    lw: int
    g_pad: tuple            # (l, r, t, b) padding of the low-res latent inside the canvas
    views: list             # windows
    nvr: int
    nvc: int
    vh: int
    vw: int
    v_pad: tuple            # (l, r, t, b) padding of a view crop inside the canvas
    tables: dict = field(default_factory=dict)   # name -> list[int]
    flags: int = 0          # PLAN_* bits (ed_plan_t.flags)

    @property
    def nv(self):
        return len(self.views)


def build_geometry(B, C, H, W, native, ds, window, stride, context) -> WaveGeometry:
    row_src, mrow_lo, mrow_n = resample_axis(H, ds[0])
    col_src, mcol_lo, mcol_n = resample_axis(W, ds[1])
    lh, lw = len(row_src) // 2, len(col_src) // 2
    # views (window collapse when no context fits, ed:820-825)
    h_ws = H if window + context >= H else window
    w_ws = W if window + context >= W else window
    wins, nvr, nvc = sliding_windows(H, W, h_ws, w_ws, stride)
    n = context // 2
    vt = []
    shapes = set()
    for (h0, h1, w0, w1) in wins:
        n_t, n_b = context_extent(h0, h1, n, H)
        n_l, n_r = context_extent(w0, w1, n, W)
        vt += [h0, h1, w0, w1, h0 - n_t, w0 - n_l, n_t, n_l]
        shapes.add((h1 - h0 + n_t + n_b, w1 - w0 + n_l + n_r))
    if len(shapes) != 1:
        raise ValueError(f"views have different crop shapes {shapes}; the reference's torch.cat (ed:845) fails too")
    vh, vw = shapes.pop()
    if vh > native or vw > native:
        raise ValueError(f"view crop {vh}x{vw} exceeds the UNet native size {native}")
    row_starts = [wins[r * nvc][0] for r in range(nvr)]
    col_starts = [wins[c][2] for c in range(nvc)]
    vrow_first, vrow_cnt = cover_ranges(row_starts, h_ws, H)
    vcol_first, vcol_cnt = cover_ranges(col_starts, w_ws, W)
    tp, bp = pad_split(native, lh)
    lp, rp = pad_split(native, lw)
    vtp, vbp = pad_split(native, vh)
    vlp, vrp = pad_split(native, vw)
    g = WaveGeometry(B, C, H, W, native, lh, lw, (lp, rp, tp, bp), wins, nvr, nvc, vh, vw, (vlp, vrp, vtp, vbp))
    up_row, up_col = nearest_index(lh, H), nearest_index(lw, W)
    down_row, down_col = nearest_index(H, lh), nearest_index(W, lw)
    g.tables = dict(row_src=row_src, col_src=col_src, mrow_lo=mrow_lo, mrow_n=mrow_n, mcol_lo=mcol_lo, mcol_n=mcol_n,
                    up_row=up_row, up_col=up_col, down_row=down_row, down_col=down_col,
                    views=vt, vrow_first=vrow_first, vrow_cnt=vrow_cnt, vcol_first=vcol_first, vcol_cnt=vcol_cnt)
    # per-row / per-column view offsets: with the windows on a grid, the canvas row of latent row y inside the view of the
    # FIRST covering grid row depends on y only, likewise for columns (used by the exact-1/2 fast-path epilogue)
    A = lambda v: np.asarray(v, dtype=np.int64)
    vta = A(vt).reshape(-1, 8)
    ys, xs = np.arange(H), np.arange(W)
    vr_first, vc_first = A(vrow_first), A(vcol_first)
    row_view, col_view = vr_first * nvc, vc_first                       # a view of that grid row / column
    vrow_off = vtp + vta[row_view, 6] + (ys - vta[row_view, 0])
    vcol_off = vlp + vta[col_view, 7] + (xs - vta[col_view, 2])
    g.tables.update(vrow_off=vrow_off.tolist(), vcol_off=vcol_off.tolist())
    # static per-pixel / per-cell references (what the epilogue would otherwise re-derive per pixel, per channel)
    ur, uc = A(up_row), A(up_col)
    dir_off = ((tp + ur) * native)[:, None] + (lp + uc)[None, :]                                  # (H, W)
    single = (A(vrow_cnt) == 1)[:, None] & (A(vcol_cnt) == 1)[None, :]
    view = vr_first[:, None] * nvc + vc_first[None, :]
    voff = (vrow_off * native)[:, None] + vcol_off[None, :]
    pix = np.stack([dir_off, np.where(single, view, -1), np.where(single, voff, 0), ur[:, None] * lw + uc[None, :]], axis=-1)
    rs, cs = A(row_src), A(col_src)
    r, c = np.arange(lh), np.arange(lw)
    cand = np.stack([(rs[2 * r + (j >> 1)] * W)[:, None] + cs[2 * c + (j & 1)][None, :] for j in range(4)], axis=-1)
    dr, dc = A(down_row), A(down_col)
    down = np.stack([(dr * W)[:, None] + dc[None, :], dir_off[dr][:, dc]], axis=-1)
    g.tables.update(pix_ref=pix.reshape(-1).tolist(), cell_cand=cand.reshape(-1).tolist(), cell_down=down.reshape(-1).tolist())
    g.flags = PLAN_HALF_FAST if _half_fast(g, H, W, native, lh, lw, rs, cs, ur, uc, dr, dc, A(mrow_lo), A(mrow_n), A(mcol_lo),
                                            A(mcol_n), A(vrow_cnt), A(vcol_cnt), vr_first, vc_first, vrow_off, vcol_off,
                                            tp, lp) else 0
    return g


PLAN_HALF_FAST = 1   # ED_PLAN_HALF_FAST of include/elastic_b200.h


def _half_fast(g, H, W, native, lh, lw, rs, cs, ur, uc, dr, dc, mrow_lo, mrow_n, mcol_lo, mcol_n, vrow_cnt, vcol_cnt,
               vr_first, vc_first, vrow_off, vcol_off, tp, lp):
    """True when the plan is the exact 1/2-ratio geometry with tiling views that the fast-path epilogue kernels index
    arithmetically (every tiled BASELINE config).  The C ABI documents this list at ED_PLAN_HALF_FAST; the kernels trust
    the flag, so every identity is checked here."""
    if g.C != 4 or H % 2 or W % 8 or 2 * lh != H or 2 * lw != W or native % 8 or lp % 4:
        return False
    ys, xs = np.arange(H), np.arange(W)
    ok = (np.array_equal(ur, ys >> 1) and np.array_equal(uc, xs >> 1)              # nearest-up reads cell (y/2, x/2)
          and np.array_equal(rs, ys) and np.array_equal(cs, xs)                    # the 2x-resized grid IS the latent
          and np.array_equal(dr, 2 * np.arange(lh)) and np.array_equal(dc, 2 * np.arange(lw))   # nearest-down reads (2r, 2c)
          and np.array_equal(mrow_lo, ys) and np.all(mrow_n == 1)                  # restored mask = resized mask
          and np.array_equal(mcol_lo, xs) and np.all(mcol_n == 1)
          and np.all(vrow_cnt == 1) and np.all(vcol_cnt == 1))                     # every pixel under exactly one window
    if not ok:
        return False
    # the two rows of a row pair lie in the same view on consecutive canvas rows
    if not (np.array_equal(vr_first[0::2], vr_first[1::2]) and np.array_equal(vrow_off[0::2] + 1, vrow_off[1::2])):
        return False
    # every aligned group of 8 columns lies in one view, contiguous, starting at a canvas column that is a multiple of 8
    vc8, vo8 = vc_first.reshape(-1, 8), vcol_off.reshape(-1, 8)
    return bool(np.all(vc8 == vc8[:, :1]) and np.all(vo8 == vo8[:, :1] + np.arange(8)) and np.all(vo8[:, 0] % 8 == 0))


@dataclass
class TileGeometry:
    tiles: list
    ntr: int
    ntc: int
    core: int
    pad: int
    stride: int
    tables: dict


def build_tiles(H, W, sample_size, scale, low_vram=False, core=None, pad=None) -> TileGeometry:
    """Decode tiles of tiled_decode (ed:276-287): core = sample_size//4, stride = core, pad = sample_size//scale*3
    (low_vram: stride = core//2, pad = core).  `core` / `pad` override the reference's sizes (stride = core): the opt-in
    de-duplicated decode uses larger cores so that each pixel is decoded fewer times."""
    if core is not None or pad is not None:
        ref_core = sample_size // 4
        core = int(core if core is not None else ref_core)
        pad = int(pad if pad is not None else sample_size // scale * 3)
        stride = core
    else:
        core = sample_size // 4
        stride, pad = core, sample_size // scale * 3
        if low_vram:
            stride, pad = core // 2, core
    wins, ntr, ntc = sliding_windows(H, W, core, core, stride)
    flat = [v for w in wins for v in w]
    trf, trc = cover_ranges([wins[r * ntc][0] for r in range(ntr)], core, H)
    tcf, tcc = cover_ranges([wins[c][2] for c in range(ntc)], core, W)
    return TileGeometry(wins, ntr, ntc, core, pad, stride,
                        dict(tiles=flat, trow_first=trf, trow_cnt=trc, tcol_first=tcf, tcol_cnt=tcc))


def cond_geometry(geo: WaveGeometry, scale, window, context):
    """Index tables of the ControlNet condition batch (reference elastic_diffusion_w_controlnet.py "cn:N"):
    nearest-upsample maps from the prepared condition (lh*scale, lw*scale) to full pixel size (cn:933) and the pixel
    origin of every view's context box computed with the window coordinates x scale and n = context*scale//2
    (cn:946-949; the reference hard-codes 8)."""
    H, W = geo.H, geo.W
    row_map = nearest_index(geo.lh * scale, H * scale)
    col_map = nearest_index(geo.lw * scale, W * scale)
    n = (context * scale) // 2
    origin = []
    for (h0, h1, w0, w1) in geo.views:
        n_t, n_b = context_extent(h0 * scale, h1 * scale, n, H * scale)
        n_l, n_r = context_extent(w0 * scale, w1 * scale, n, W * scale)
        if ((h1 - h0) * scale + n_t + n_b, (w1 - w0) * scale + n_l + n_r) != (geo.vh * scale, geo.vw * scale):
            raise ValueError("condition view and latent view sizes disagree (odd context size): the reference's ControlNet "
                             "call fails on this configuration too")
        origin += [h0 * scale - n_t, w0 * scale - n_l]
    return row_map, col_map, origin
