"""Per-step host cost of the denoise loop (tiny UNet => the GPU is nearly idle, the step time is the host floor), with a
cProfile of the loop."""
import cProfile
import importlib
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

P = bench.pkg()
dev = torch.device("cuda")
unet, vae, txt = bench.build_modules("tiny", dev, torch.bfloat16)
ed = P.ElasticDiffusion.from_components(dev, unet, vae, None, txt, sd_version="XL1.0", view_batch_size=16, projection_dim=32)
ed.autocast, ed.unet_input_dtype, ed.use_cuda_graphs = False, torch.bfloat16, True
kw = dict(bench.GEN, height=1024, width=2048, num_inference_steps=50, resampling_steps=7, progress=lambda it: it)
ed.seed_everything(0)
ed.denoise(max_steps=4, **kw)
torch.cuda.synchronize()
ed.seed_everything(0)
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
ed.denoise(max_steps=12, **kw)
pr.disable()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / 12 * 1e3:.1f} ms per step (host floor, tiny UNet)")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
