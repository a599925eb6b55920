"""Where does one wave's UNet forward spend its GPU time?  (torch.profiler, top CUDA kernels; run on the GPU box)"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import standins as syn  # noqa: E402
dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
unet = syn.StandInUNet("XL1.0", device=dev, dtype=torch.bfloat16).eval()
x = torch.randn(n, 4, 128, 128, device=dev, dtype=torch.bfloat16)
ehs = torch.randn(n, 77, 2048, device=dev, dtype=torch.bfloat16)
kw = {"added_cond_kwargs": {"text_embeds": torch.randn(n, 1280, device=dev, dtype=torch.bfloat16),
                            "time_ids": torch.tensor([[4096., 8192, 0, 0, 4096, 8192]], device=dev).repeat(n, 1)}}
t = torch.tensor(981, device=dev)
with torch.no_grad():
    for _ in range(3):
        unet(x, t, encoder_hidden_states=ehs, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        unet(x, t, encoder_hidden_states=ehs, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"batch {n}: {e0.elapsed_time(e1) / 3:.1f} ms per forward")
    from torch.utils.flop_counter import FlopCounterMode
    with FlopCounterMode(display=False) as fc:
        unet(x, t, encoder_hidden_states=ehs, **kw)
    fl = fc.get_total_flops()
    print(f"flops per forward {fl / 1e12:.2f} TFLOP ({fl / n / 1e12:.2f} per sample) -> {fl / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e12:.0f} TFLOP/s")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        unet(x, t, encoder_hidden_states=ehs, **kw)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
