"""CUDA-graph replay time of the stand-in UNet vs batch size (explains the multi-GPU scaling of bench.py: with wave
samples sharded over N ranks a rank runs batch ceil(20/N) then ceil(6/N))."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import standins as syn  # noqa: E402
dev = torch.device("cuda")
unet = syn.StandInUNet("XL1.0", device=dev, dtype=torch.bfloat16).eval()
t = torch.tensor(981, device=dev)
res = {}
with torch.no_grad():
    for n in (1, 2, 3, 5, 6, 10, 20):
        x = torch.randn(n, 4, 128, 128, device=dev, dtype=torch.bfloat16)
        ehs = torch.randn(n, 77, 2048, device=dev, dtype=torch.bfloat16)
        kw = {"added_cond_kwargs": {"text_embeds": torch.randn(n, 1280, device=dev, dtype=torch.bfloat16),
                                    "time_ids": torch.tensor([[4096., 8192, 0, 0, 4096, 8192]], device=dev).repeat(n, 1)}}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                unet(x, t, encoder_hidden_states=ehs, **kw)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            unet(x, t, encoder_hidden_states=ehs, **kw)
        for _ in range(2):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[n] = e0.elapsed_time(e1) / 5
        print(f"batch {n:2d}: {res[n]:7.2f} ms  ({res[n] / n:6.2f} ms/sample)", flush=True)
        del g
for N in (1, 2, 4, 8):
    a, b = -(-20 // N), -(-6 // N)
    if a in res and b in res:
        print(f"N={N}: UNet floor per step = t({a}) + t({b}) = {res[a] + res[b]:.1f} ms")
