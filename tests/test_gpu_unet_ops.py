"""GPU: the opt-in fused UNet ops (csrc/unet_ops.cu through `unet_ops.FusedOps`) against the plain torch formulation."""
import pytest
import torch
import torch.nn.functional as F

import standins
from conftest import PKG

pytestmark = pytest.mark.gpu
ops_mod = __import__("importlib").import_module(PKG.__name__ + ".unet_ops")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 1024, 10240), (3, 4096, 5120), (1, 77, 64), (5, 16)])
def test_geglu_is_bit_identical_to_the_two_torch_ops(dtype, shape):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(shape, device="cuda", generator=g) * 2.5).to(dtype)
    fo = ops_mod.FusedOps()
    got = fo.geglu(x)
    a, b = x.chunk(2, dim=-1)
    want = a * F.gelu(b)
    assert fo.calls["geglu"] == 1 and got.shape == want.shape
    assert torch.equal(got, want), (got.float() - want.float()).abs().max().item()


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2 ** -7), (torch.float16, 2 ** -10), (torch.float32, 2e-6)])
@pytest.mark.parametrize("N,C,H,W", [(1, 320, 128, 128), (3, 640, 64, 64), (2, 1280, 32, 32), (1, 1920, 32, 32), (4, 64, 8, 8)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm_silu_matches_torch(dtype, tol, N, C, H, W, silu):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(N, C, H, W, device="cuda", generator=g) * 1.7 + 0.6).to(dtype)
    gn = torch.nn.GroupNorm(32, C).to("cuda", dtype)
    with torch.no_grad():
        gn.weight.copy_(torch.randn(C, device="cuda", generator=g) * 0.3 + 1)
        gn.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.2)
        fo = ops_mod.FusedOps()
        got = fo.group_norm_silu(gn, x) if silu else fo.group_norm(gn, x)
        want64 = F.group_norm(x.double(), 32, gn.weight.double(), gn.bias.double(), gn.eps)
        want64 = F.silu(want64) if silu else want64
        torch_out = F.silu(gn(x)) if silu else gn(x)
    assert fo.calls["group_norm"] == 1
    err_mine = (got.double() - want64).abs().max().item()
    err_torch = (torch_out.double() - want64).abs().max().item()
    scale = want64.abs().max().item()
    assert err_mine <= max(2 * err_torch, tol * scale), (err_mine, err_torch, scale)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("N,Cin,Cout,H,k,stride", [(2, 320, 320, 64, 3, 1), (1, 640, 1280, 32, 3, 1), (3, 960, 640, 32, 1, 1),
                                                   (2, 320, 320, 64, 3, 2), (1, 4, 320, 128, 3, 1)])
def test_conv_add_is_bit_identical_to_the_separate_torch_ops(dtype, N, Cin, Cout, H, k, stride):
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(3)
    conv = torch.nn.Conv2d(Cin, Cout, k, stride=stride, padding=k // 2).to("cuda", dtype)
    x = torch.randn(N, Cin, H, H, device="cuda", generator=g).to(dtype)
    Ho = H // stride
    temb = torch.randn(N, Cout, device="cuda", generator=g).to(dtype)
    res = torch.randn(N, Cout, Ho, Ho, device="cuda", generator=g).to(dtype)
    fo = ops_mod.FusedOps()
    with torch.no_grad():
        for per_nc, residual in ((None, None), (temb, None), (None, res), (temb, res)):
            got = fo.conv_add(conv, x, per_nc=per_nc, residual=residual)
            want = ops_mod.TorchOps.conv_add(conv, x, per_nc=per_nc, residual=residual)
            assert torch.equal(got, want), (per_nc is not None, residual is not None, (got.float() - want.float()).abs().max().item())
    assert fo.calls["conv_add"] == 4 and fo.calls["fallback"] == 0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("N,Cin,Cout,H,k,stride", [(2, 320, 320, 64, 3, 1), (1, 640, 1280, 32, 3, 1), (3, 960, 640, 32, 1, 1),
                                                   (2, 320, 320, 64, 3, 2), (1, 1280, 1280, 16, 3, 1)])
def test_conv_add_in_cudnn_native_layout_matches_the_separate_torch_ops(dtype, N, Cin, Cout, H, k, stride):
    """channels_last_convs: weights kept channels-last once, conv output left in NHWC, ed_bias_add_nhwc transposes back while it
    adds.  cuDNN may pick another engine for the other layout, so the conv itself is compared within rounding noise, and
    the epilogue exactly: ed_bias_add_nhwc(conv_nhwc) == the torch adds applied to the SAME conv output."""
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(3)
    conv = torch.nn.Conv2d(Cin, Cout, k, stride=stride, padding=k // 2).to("cuda", dtype)
    x = torch.randn(N, Cin, H, H, device="cuda", generator=g).to(dtype)
    Ho = H // stride
    temb = torch.randn(N, Cout, device="cuda", generator=g).to(dtype)
    res = torch.randn(N, Cout, Ho, Ho, device="cuda", generator=g).to(dtype)
    fo = ops_mod.FusedOps(channels_last_convs=True)
    with torch.no_grad():
        y_cl = F.conv2d(x.contiguous(memory_format=torch.channels_last), conv.weight.contiguous(memory_format=torch.channels_last),
                        None, conv.stride, conv.padding)
        for per_nc, residual in ((None, None), (temb, None), (None, res), (temb, res)):
            got = fo.conv_add(conv, x, per_nc=per_nc, residual=residual)
            assert got.is_contiguous() and got.shape == (N, Cout, Ho, Ho)
            exact = y_cl.contiguous() + conv.bias[None, :, None, None]
            if per_nc is not None:
                exact = exact + per_nc[:, :, None, None]
            if residual is not None:
                exact = exact + residual
            assert torch.equal(got, exact), (got.float() - exact.float()).abs().max().item()
            want = ops_mod.TorchOps.conv_add(conv, x, per_nc=per_nc, residual=residual)
            tol = {torch.bfloat16: 2 ** -6, torch.float16: 2 ** -9, torch.float32: 1e-4}[dtype]
            assert (got.float() - want.float()).abs().max().item() <= tol * want.float().abs().max().item()
    assert fo.calls.get("conv_add_nhwc", 0) == (4 if (Ho * Ho) % 64 == 0 else 0), fo.calls


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2 ** -7), (torch.float16, 2 ** -10), (torch.float32, 2e-6)])
@pytest.mark.parametrize("shape", [(2, 4096, 640), (3, 1024, 1280), (5, 77, 2048), (7, 64), (1, 9, 1024)])
def test_layernorm_matches_torch(dtype, tol, shape):
    g = torch.Generator(device="cuda").manual_seed(4)
    D = shape[-1]
    x = (torch.randn(shape, device="cuda", generator=g) * 2.0 + 0.3).to(dtype)
    ln = torch.nn.LayerNorm(D).to("cuda", dtype)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(D, device="cuda", generator=g) * 0.3 + 1)
        ln.bias.copy_(torch.randn(D, device="cuda", generator=g) * 0.2)
        fo = ops_mod.FusedOps()
        got = fo.layer_norm(ln, x)
        want64 = F.layer_norm(x.double(), (D,), ln.weight.double(), ln.bias.double(), ln.eps)
        torch_out = ln(x)
    assert fo.calls["layer_norm"] == 1
    err_mine, err_torch = (got.double() - want64).abs().max().item(), (torch_out.double() - want64).abs().max().item()
    assert err_mine <= max(2 * err_torch, tol * want64.abs().max().item()), (err_mine, err_torch)


def test_large_offset_groups_do_not_cancel():
    """|mean| >> std: the shifted sums keep the variance (a naive E[x^2] - mean^2 in fp32 would lose it)."""
    x = (torch.randn(2, 64, 32, 32, device="cuda") * 0.01 + 300.0)
    gn = torch.nn.GroupNorm(32, 64).cuda()
    with torch.no_grad():
        got = ops_mod.FusedOps().group_norm(gn, x)
        want = F.group_norm(x.double(), 32, gn.weight.double(), gn.bias.double(), gn.eps).float()
    assert (got - want).abs().max().item() < 2e-2 and abs(got.std().item() - 1) < 0.05


def test_standin_unet_with_fused_ops_matches_plain_ops_and_captures_in_a_cuda_graph():
    torch.manual_seed(0)
    unet = standins.StandInUNet("tiny-xl", device="cuda", dtype=torch.bfloat16).eval()
    n = 3
    x = torch.randn(n, 4, 128, 128, device="cuda")
    ehs = torch.randn(n, 77, 64, device="cuda", dtype=torch.bfloat16)
    kw = {"added_cond_kwargs": {"text_embeds": torch.randn(n, 32, device="cuda", dtype=torch.bfloat16),
                                "time_ids": torch.tensor([[4096., 8192, 0, 0, 4096, 8192]], device="cuda").repeat(n, 1)}}
    t = torch.tensor(981, device="cuda")
    with torch.no_grad():
        plain = unet(x, t, encoder_hidden_states=ehs, **kw)["sample"].float()
        fo = ops_mod.FusedOps()
        unet.set_ops(fo)
        fused = unet(x, t, encoder_hidden_states=ehs, **kw)["sample"].float()
        assert min(fo.calls[k] for k in ("geglu", "group_norm", "layer_norm", "conv_add")) > 0 and fo.calls["fallback"] == 0, fo.calls
        rel = (fused - plain).pow(2).mean().sqrt().item() / plain.pow(2).mean().sqrt().item()
        assert rel < 2e-2, rel                       # bf16 network: GroupNorm statistics differ in the last bits
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            unet(x, t, encoder_hidden_states=ehs, **kw)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = unet(x, t, encoder_hidden_states=ehs, **kw)["sample"]
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out.float(), fused)
        unet.set_ops(ops_mod.TorchOps)
        assert torch.equal(unet(x, t, encoder_hidden_states=ehs, **kw)["sample"].float(), plain)
