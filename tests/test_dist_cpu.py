"""CPU, world_size 2 over gloo: the wave-sample sharding + all-gather of the product's `_unet` (the N>1 path of
DESIGN.md section 6) returns exactly what the unsharded call returns, for even and ragged sample counts."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_grad_enabled(False)      # the product calls _unet under @torch.no_grad (denoise)
        from conftest import make_ed
        torch.manual_seed(0)
        ed = make_ed("XL1.0", 16)
        ed.autocast = False
        ok = True
        for n in (6, 5, 20, 1):
            ed.last_run = dict(unet_calls=0, unet_samples=0, collectives=0)
            ed._shard_dtype = None
            g = torch.Generator().manual_seed(n)
            canvas = torch.randn(n, 4, 128, 128, generator=g)
            text = torch.randn(n, 77, 16, generator=g)
            pool = torch.randn(n, 8, generator=g)
            tid = torch.tensor([[4096., 8192., 0., 0., 4096., 8192.]]).repeat(n, 1)
            t = torch.tensor(981)
            ed.shard_waves = True
            got = ed._unet(canvas, t, text, pool, tid)
            mine = ed.last_run["unet_samples"]
            ed.shard_waves = False
            want = ed._unet(canvas, t, text, pool, tid)
            per = (n + world - 1) // world
            ok &= torch.allclose(got, want, atol=1e-6) and got.shape == want.shape
            ok &= mine == max(0, min((rank + 1) * per, n) - min(rank * per, n))
            ok &= ed.last_run["collectives"] == 1
        q.put((rank, bool(ok)))
    except Exception as e:  # report instead of leaving the parent to time out
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_wave_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _decode_worker(rank, world, port, q):
    """Tile-sharded decode on gloo: every rank decodes the contiguous block of tiles the product's
    `shard_range` gives it, the centre crops are all-gathered and blended with pad = 0 (the product's "nccl" exchange of
    `tiled_decode`, kernels replaced by the spec emulations) -> equals the oracle's unsharded decode_tiled."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_grad_enabled(False)
        from conftest import PKG, oracle_models
        from oracle import reference_port as rp
        from oracle import wave_spec as ws
        ok = True
        for (B, H, W, low_vram) in ((1, 64, 96, False), (2, 40, 72, True)):
            m = oracle_models("2.1", 4)
            z = torch.randn(B, 4, H, W, generator=torch.Generator().manual_seed(5))
            want = rp.decode_tiled(m, z, low_vram=low_vram)
            tg = PKG.geometry.build_tiles(H, W, 64, 8, low_vram)
            n = len(tg.tiles) * B
            per, lo, hi = PKG.pipeline.shard_range(n, world, rank)
            boxes = ws.spec_tile_gather(z, tg.tiles, tg.core, tg.pad)
            dec = m.vae.decode(boxes[lo:hi] / m.vae.config.scaling_factor).sample if hi > lo else boxes.new_zeros(0, 3, 1, 1)
            p0, c = tg.pad * 8, tg.core * 8
            send = torch.zeros(per, 3, c, c)
            send[:hi - lo] = dec[:, :, p0:p0 + c, p0:p0 + c]
            crops = torch.empty(world * per, 3, c, c)
            dist.all_gather_into_tensor(crops, send)
            got = ws.spec_tile_blend(crops[:n], tg.tiles, B, H, W, tg.core, 0, 8)
            ok &= (got - want).abs().max().item() <= 1e-5 and hi - lo > 0      # conv batching: last-bit noise
        q.put((rank, bool(ok)))
    except Exception as e:
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_tile_sharded_decode_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_decode_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
