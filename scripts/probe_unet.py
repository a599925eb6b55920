"""GPU probe of the stand-in UNet's small-batch cost (what caps the 1 -> 8 GPU curve: t(n) ~ 10.3 ms + 10.2 ms * n).

    python scripts/probe_unet.py [out.json]

Measures, all as CUDA-graph replays (what the pipeline runs):
  sweep          t(n) for n in 1..20                                      (NCHW weights, cudnn.benchmark off: round-1 setup)
  sweep_bench    same with torch.backends.cudnn.benchmark = True
  sweep_cl       same with channels_last weights / activations
  streams        k concurrent batch-1 graphs on k streams vs one batch-k graph (k = 2, 3)
  profile_b1     torch.profiler kernel table of ONE eager batch-1 forward: kernel count, time in kernels shorter than
                 5 us / 10 us (the launch-bound floor), top kernels by total time
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import standins as syn  # noqa: E402

dev = torch.device("cuda")
OUT = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out/probe_unet.json"
res = {}


def inputs(n, cl=False):
    x = torch.randn(n, 4, 128, 128, device=dev)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    ehs = torch.randn(n, 77, 2048, device=dev, dtype=torch.bfloat16)
    kw = {"added_cond_kwargs": {"text_embeds": torch.randn(n, 1280, device=dev, dtype=torch.bfloat16),
                                "time_ids": torch.tensor([[4096., 8192, 0, 0, 4096, 8192]], device=dev).repeat(n, 1)}}
    return x, ehs, kw


t981 = torch.tensor(981, device=dev)


def capture(unet, n, cl=False):
    x, ehs, kw = inputs(n, cl)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(2):
            unet(x, t981, encoder_hidden_states=ehs, **kw)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = unet(x, t981, encoder_hidden_states=ehs, **kw)["sample"]
    return g, out, (x, ehs, kw)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sweep(unet, ns, cl=False):
    out = {}
    for n in ns:
        g, _, keep = capture(unet, n, cl)
        out[n] = round(timed(g.replay), 3)
        del g, keep
    return out


def save():
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)


def kernel_table(unet, n, key):
    """torch.profiler kernel table of ONE eager batch-n forward: kernel count, time in short kernels, top kernels"""
    x, ehs, kw = inputs(n)
    with torch.no_grad():
        for _ in range(2):
            unet(x, t981, encoder_hidden_states=ehs, **kw)
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            unet(x, t981, encoder_hidden_states=ehs, **kw)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    dt = lambda e: getattr(e, "device_time", None) or getattr(e, "cuda_time", 0.0)
    durs = [dt(e) for e in evs]          # us
    by = {}
    for e in evs:
        a = by.setdefault(e.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += dt(e)
    top = sorted(by.items(), key=lambda kv: -kv[1][1])[:25]
    res[key] = {"kernels": len(durs), "sum_ms": round(sum(durs) / 1e3, 3),
                "n_under_5us": sum(d < 5 for d in durs), "ms_under_5us": round(sum(d for d in durs if d < 5) / 1e3, 3),
                "n_under_10us": sum(d < 10 for d in durs), "ms_under_10us": round(sum(d for d in durs if d < 10) / 1e3, 3),
                "n_under_20us": sum(d < 20 for d in durs), "ms_under_20us": round(sum(d for d in durs if d < 20) / 1e3, 3),
                "top": [{"name": k, "count": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2)} for k, v in top]}
    print(key, {k: v for k, v in res[key].items() if k != "top"}, flush=True)


unet = syn.StandInUNet("XL1.0", device=dev, dtype=torch.bfloat16).eval()
NS = (1, 2, 3, 4, 5, 6, 8, 10, 20)
if "--fused-sweep" in sys.argv:      # t(n) with the library's fused GEGLU / GroupNorm(+SiLU) kernels inside the UNet, then plain
    import importlib
    ops = importlib.import_module("elasticdiffusion-official_b200").unet_ops
    OUT = "gpurun_out/probe_unet_fused.json"
    fo = ops.FusedOps()
    unet.set_ops(fo)
    res["sweep_fused"] = sweep(unet, NS)
    res["fused_calls"] = dict(fo.calls)
    print("sweep_fused", res["sweep_fused"], flush=True)
    unet.set_ops(ops.TorchOps)
    res["sweep_torch"] = sweep(unet, (1, 3, 6, 20))
    print("sweep_torch", res["sweep_torch"], flush=True)
    fo_cl = ops.FusedOps(channels_last_convs=True)      # convs in cuDNN's native layout (weights channels-last once, fused back-transform)
    unet.set_ops(fo_cl)
    res["sweep_fused_nhwc_convs"] = sweep(unet, NS)
    res["fused_nhwc_calls"] = dict(fo_cl.calls)
    print("sweep_fused_nhwc_convs", res["sweep_fused_nhwc_convs"], flush=True)
    for n in (1, 20):
        kernel_table(unet, n, f"profile_fused_nhwc_b{n}")
    if "--quick" in sys.argv:
        save()
        sys.exit(0)
    unet.set_ops(fo)
    for n in (1, 20):
        kernel_table(unet, n, f"profile_fused_b{n}")
    save()
    sys.exit(0)
res["sweep"] = sweep(unet, NS)
print("sweep", res["sweep"], flush=True)
save()

# ---- concurrent batch-1 graphs on streams vs one batch-k graph --------------------------------------------------------
res["streams"] = {}
for k in (2, 3):
    g_batch, _, keep_b = capture(unet, k)
    singles = [capture(unet, 1) for _ in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]

    def concurrent():
        cur = torch.cuda.current_stream()
        for s, (g, _, _) in zip(streams, singles):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                g.replay()
        for s in streams:
            cur.wait_stream(s)
    res["streams"][k] = {"one_batch_k_graph_ms": round(timed(g_batch.replay), 3), "k_concurrent_b1_graphs_ms": round(timed(concurrent), 3),
                         "one_b1_graph_ms": round(timed(singles[0][0].replay), 3)}
    del g_batch, singles, keep_b
print("streams", res["streams"], flush=True)
save()

# ---- kernel table of one eager batch-1 / batch-3 forward -------------------------------------------------------------
for n in (1, 3):
    kernel_table(unet, n, f"profile_b{n}")
save()

# ---- cudnn.benchmark ---------------------------------------------------------------------------------------------------
torch.backends.cudnn.benchmark = True
res["sweep_bench"] = sweep(unet, (1, 3, 6, 20))
print("sweep_bench", res["sweep_bench"], flush=True)
save()
torch.backends.cudnn.benchmark = False

# ---- channels_last -------------------------------------------------------------------------------------------------------
unet = unet.to(memory_format=torch.channels_last)
res["sweep_cl"] = sweep(unet, (1, 3, 6, 20), cl=True)
print("sweep_cl", res["sweep_cl"], flush=True)
torch.backends.cudnn.benchmark = True
res["sweep_cl_bench"] = sweep(unet, (1, 3, 6, 20), cl=True)
print("sweep_cl_bench", res["sweep_cl_bench"], flush=True)
save()
