"""Opt-in fused ops for the UNet forward: `a * gelu(g)` of a GEGLU projection and GroupNorm(+SiLU),
hand-written sm_100a kernels in csrc/unet_ops.cu behind the C ABI (`ed_geglu`, `ed_groupnorm_silu`).

The denoising loop spends > 99.9 % of its time inside the injected UNet; these are the two largest non-GEMM costs of an
SD/SDXL UNet forward (profiles/r2_probe_unet.json).  A UNet implementation opts in by routing the two patterns through an
`ops` object with this interface (the stand-in UNet of `standins/` does; for diffusers' `UNet2DConditionModel` the same two
call sites are `GEGLU.forward` and `ResnetBlock2D.forward` / `Transformer2DModel.norm`):

    ops.geglu(x)                      # x = proj(hidden) of shape (..., 2N)  ->  x[..., :N] * gelu(x[..., N:])
    ops.group_norm(gn_module, x)      # nn.GroupNorm forward
    ops.group_norm_silu(gn_module, x) # silu(group_norm(x))

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


class FusedOps:
    """Kernels of csrc/unet_ops.cu; `calls` counts how many went to the library / to the torch fall-back."""

    def __init__(self):
        self.calls = {"geglu": 0, "group_norm": 0, "fallback": 0}
        self._ws = {}

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

    def group_norm(self, gn, x):
        return self._gn(gn, x, False)

    def group_norm_silu(self, gn, x):
        return self._gn(gn, x, True)
