"""Next-round probe (not run yet - no GPU minutes were left in round 1): does running a rank's few wave samples as
CONCURRENT batch-1 CUDA graphs on separate streams beat one small-batch graph?

Why: at 8 GPUs a rank runs batch 3 then batch 1 per step and the stand-in UNet's replay time is affine in the batch,
t(n) ~ 10.3 ms + 10.2 ms * n (scripts/sweep_unet_batch.py), i.e. ~10 ms per forward are spent in kernels too small to fill
148 SMs.  If k concurrent batch-1 graphs overlap that under-utilisation, the 8-GPU step floor t(3) + t(1) = 61 ms shrinks and
the strong-scaling gap to the ideal 6.5x closes from the UNet side.  Results are unchanged (same samples, same kernels).

    gpurun --timeout 300 -- python scripts/probe_unet_streams.py
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import standins as syn  # noqa: E402
dev = torch.device("cuda")
unet = syn.StandInUNet("XL1.0", device=dev, dtype=torch.bfloat16).eval()
t = torch.tensor(981, device=dev)


def inputs(n):
    x = torch.randn(n, 4, 128, 128, device=dev, dtype=torch.bfloat16)
    ehs = torch.randn(n, 77, 2048, device=dev, dtype=torch.bfloat16)
    kw = {"added_cond_kwargs": {"text_embeds": torch.randn(n, 1280, device=dev, dtype=torch.bfloat16),
                                "time_ids": torch.tensor([[4096., 8192, 0, 0, 4096, 8192]], device=dev).repeat(n, 1)}}
    return x, ehs, kw


def capture(n):
    x, ehs, kw = inputs(n)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(2):
            unet(x, t, encoder_hidden_states=ehs, **kw)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = unet(x, t, encoder_hidden_states=ehs, **kw)["sample"]
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


for k in (2, 3):
    g_batch, _, keep_b = capture(k)
    singles = [capture(1) for _ in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]

    def concurrent():
        cur = torch.cuda.current_stream()
        for s, (g, _, _) in zip(streams, singles):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                g.replay()
        for s in streams:
            cur.wait_stream(s)

    tb, tc, t1 = timed(g_batch.replay), timed(concurrent), timed(singles[0][0].replay)
    print(f"k={k}: one batch-{k} graph {tb:.2f} ms | {k} concurrent batch-1 graphs {tc:.2f} ms | one batch-1 graph {t1:.2f} ms", flush=True)
