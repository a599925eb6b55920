"""Opt-in fused ops for the UNet forward: `a * gelu(g)` of a GEGLU projection and GroupNorm(+SiLU),
hand-written sm_100a kernels in csrc/unet_ops.cu behind the C ABI (`ed_geglu`, `ed_groupnorm_silu`).

The denoising loop spends > 99.9 % of its time inside the injected UNet; these are the two largest non-GEMM costs of an
SD/SDXL UNet forward (profiles/r2_probe_unet.json).  A UNet implementation opts in by routing the two patterns through an
`ops` object with this interface (the stand-in UNet of `standins/` does; for diffusers' `UNet2DConditionModel` the same two
call sites are `GEGLU.forward` and `ResnetBlock2D.forward` / `Transformer2DModel.norm`):

    ops.geglu(x)                      # x = proj(hidden) of shape (..., 2N)  ->  x[..., :N] * gelu(x[..., N:])
    ops.group_norm(gn_module, x)      # nn.GroupNorm forward
    ops.group_norm_silu(gn_module, x) # silu(group_norm(x))
    ops.layer_norm(ln_module, x)      # nn.LayerNorm forward over the last dimension
    ops.conv_add(conv, x, per_nc=None, residual=None)   # conv(x) [+ per_nc[:, :, None, None]] [+ residual]

`TorchOps` is the plain PyTorch formulation (what the reference's models run); `FusedOps` calls the kernels and falls back
to `TorchOps` per call for shapes outside the kernels' domain (non-contiguous, HW % 8 != 0, CPU tensors, autograd).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import native


class TorchOps:
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
        h = conv(x)
        if per_nc is not None:
            h = h + per_nc[:, :, None, None]
        if residual is not None:
            h = h + residual
        return h


class FusedOps:
    """Kernels of csrc/unet_ops.cu; `calls` counts how many went to the library / to the torch fall-back.
    One instance serves one stream at a time (the GroupNorm workspace is reused in stream order); use one instance per
    concurrently running UNet."""

    def __init__(self, channels_last_convs=False):
        self.calls = {"geglu": 0, "group_norm": 0, "layer_norm": 0, "conv_add": 0, "fallback": 0}
        self._ws = {}
        self._w_cl = {}
        # convs in cuDNN's native NHWC layout, see _conv_add_nhwc.  Measured on B200 (profiles/r2_probe_unet_fused_nhwc.json): the
        # SDXL-shaped forward gets faster only at batch <= 2 (15.1 -> 14.4 ms at batch 1) and slower from batch 3 on (163 -> 175 ms at
        # batch 20: torch's NCHW->NHWC input pass costs more than the per-call weight transform it removes) - off by default
        self.channels_last_convs = channels_last_convs

    def _ok(self, x):
        return x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.float16, torch.bfloat16) and \
            not (torch.is_grad_enabled() and x.requires_grad)

    def geglu(self, x):
        n2 = x.shape[-1]
        if not self._ok(x) or n2 % 16:
            self.calls["fallback"] += 1
            return TorchOps.geglu(x)
        out = torch.empty(x.shape[:-1] + (n2 // 2,), device=x.device, dtype=x.dtype)
        native.check(native.lib().ed_geglu(native.ptr(x), native.ptr(out), x.numel() // n2, n2 // 2, native.dtype_code(x.dtype),
                                           native.stream_handle()), "ed_geglu")
        self.calls["geglu"] += 1
        return out

    def _gn(self, gn, x, silu):
        if not self._ok(x) or x.dim() < 3:
            self.calls["fallback"] += 1
            return TorchOps.group_norm_silu(gn, x) if silu else gn(x)
        N, C = x.shape[0], x.shape[1]
        HW = x.numel() // (N * C)
        w, b = gn.weight, gn.bias
        if HW % 8 or C % gn.num_groups or (w is not None and w.dtype != x.dtype) or (b is not None and b.dtype != x.dtype):
            self.calls["fallback"] += 1
            return TorchOps.group_norm_silu(gn, x) if silu else gn(x)
        L = native.lib()
        S = L.ed_groupnorm_split(N, C, HW, gn.num_groups)
        key = (x.device, N * gn.num_groups * S)
        ws = self._ws.get(key)          # one workspace per size: stable addresses (CUDA-graph capture), stream-ordered reuse
        if ws is None:
            ws = self._ws[key] = torch.empty(N * gn.num_groups * S * 2, device=x.device, dtype=torch.float32)
        out = torch.empty_like(x)
        native.check(L.ed_groupnorm_silu(native.ptr(x), native.ptr(w), native.ptr(b), native.ptr(out), native.ptr(ws), N, C, HW,
                                         gn.num_groups, float(gn.eps), int(silu), native.dtype_code(x.dtype),
                                         native.stream_handle()), "ed_groupnorm_silu")
        self.calls["group_norm"] += 1
        return out

    def layer_norm(self, ln, x):
        D = x.shape[-1]
        w, b = ln.weight, ln.bias
        if (not self._ok(x) or len(ln.normalized_shape) != 1 or ln.normalized_shape[0] != D or D % 8 or D > 2048 or
                (w is not None and w.dtype != x.dtype) or (b is not None and b.dtype != x.dtype)):
            self.calls["fallback"] += 1
            return ln(x)
        out = torch.empty_like(x)
        native.check(native.lib().ed_layernorm(native.ptr(x), native.ptr(w), native.ptr(b), native.ptr(out), x.numel() // D, D,
                                               float(ln.eps), native.dtype_code(x.dtype), native.stream_handle()), "ed_layernorm")
        self.calls["layer_norm"] += 1
        return out

    def conv_add(self, conv, x, per_nc=None, residual=None):
        """cuDNN conv WITHOUT its bias, then one vectorised in-place pass for bias + time-embedding + residual (torch runs the
        bias and the time-embedding as two unvectorised broadcast adds); bit-identical to the separate ops."""
        if (not self._ok(x) or x.dim() != 4 or not isinstance(conv, torch.nn.Conv2d) or conv.padding_mode != "zeros" or
                (conv.bias is not None and conv.bias.dtype != x.dtype)):
            self.calls["fallback"] += 1
            return TorchOps.conv_add(conv, x, per_nc, residual)
        if self.channels_last_convs and conv.groups == 1 and conv.in_channels % 8 == 0 and conv.out_channels % 64 == 0:
            out = self._conv_add_nhwc(conv, x, per_nc, residual)
            if out is not None:
                return out
        y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        N, C = y.shape[0], y.shape[1]
        HW = y.shape[2] * y.shape[3]
        ok = y.is_contiguous() and HW % 8 == 0 and \
            (per_nc is None or (per_nc.is_contiguous() and per_nc.dtype == y.dtype and tuple(per_nc.shape) == (N, C))) and \
            (residual is None or (residual.is_contiguous() and residual.dtype == y.dtype and residual.shape == y.shape))
        if not ok:
            self.calls["fallback"] += 1
            if conv.bias is not None:
                y = y + conv.bias[None, :, None, None]
            if per_nc is not None:
                y = y + per_nc[:, :, None, None]
            return y if residual is None else y + residual
        if conv.bias is None and per_nc is None and residual is None:
            return y
        native.check(native.lib().ed_bias_add(native.ptr(y), native.ptr(conv.bias), native.ptr(per_nc), native.ptr(residual), N, C, HW,
                                              native.dtype_code(y.dtype), native.stream_handle()), "ed_bias_add")
        self.calls["conv_add"] += 1
        return y

    def _conv_add_nhwc(self, conv, x, per_nc, residual):
        """The conv in cuDNN's native layout: weights kept channels-last ONCE (no per-call weight transform: 1.8 ms of a batch-1
        SDXL forward), input converted by one torch pass, output left in NHWC and brought back to NCHW by the epilogue kernel
        itself (ed_bias_add_nhwc) instead of cuDNN's separate nhwcToNchw pass."""
        key = id(conv)
        ent = self._w_cl.get(key)
        if ent is None or ent[0] is not conv.weight or ent[1] != conv.weight._version:
            ent = self._w_cl[key] = (conv.weight, conv.weight._version, conv.weight.detach().contiguous(memory_format=torch.channels_last))
        y = F.conv2d(x.contiguous(memory_format=torch.channels_last), ent[2], None, conv.stride, conv.padding, conv.dilation, 1)
        N, C, H, W = y.shape
        HW = H * W
        if not y.is_contiguous(memory_format=torch.channels_last) or HW % 64 or \
                (per_nc is not None and not (per_nc.is_contiguous() and per_nc.dtype == y.dtype and tuple(per_nc.shape) == (N, C))) or \
                (residual is not None and not (residual.is_contiguous() and residual.dtype == y.dtype and residual.shape == y.shape)):
            return None
        out = torch.empty((N, C, H, W), device=y.device, dtype=y.dtype)
        native.check(native.lib().ed_bias_add_nhwc(native.ptr(y), native.ptr(out), native.ptr(conv.bias), native.ptr(per_nc),
                                                   native.ptr(residual), N, C, HW, native.dtype_code(y.dtype), native.stream_handle()),
                     "ed_bias_add_nhwc")
        self.calls["conv_add_nhwc"] = self.calls.get("conv_add_nhwc", 0) + 1
        return out

    def group_norm(self, gn, x):
        return self._gn(gn, x, False)

    def group_norm_silu(self, gn, x):
        return self._gn(gn, x, True)
