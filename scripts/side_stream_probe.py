import torch, time
dev = torch.device("cuda")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
def busy(n=60):
    for _ in range(n): torch.mm(a, a)
g = torch.cuda.CUDAGraph()
s0 = torch.cuda.Stream(); s0.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s0): busy(3)
torch.cuda.current_stream().wait_stream(s0)
with torch.cuda.graph(g): busy(60)
torch.cuda.synchronize()
t0=time.perf_counter(); g.replay(); torch.cuda.synchronize(); print("graph ms", (time.perf_counter()-t0)*1e3)
pin = torch.empty(8192, dtype=torch.int64).pin_memory()
for name, side in [("default-prio", torch.cuda.Stream()), ("high-prio", torch.cuda.Stream(priority=-1))]:
    for mode in ("graph", "eager"):
        ev = torch.cuda.Event()
        torch.cuda.synchronize()
        if mode == "graph": g.replay()
        else: busy(60)
        t0 = time.perf_counter()
        with torch.cuda.stream(side):
            d = torch.randint(0, 101, (8192,), device=dev)
            pin.copy_(d, non_blocking=True)
            ev.record(side)
        t1 = time.perf_counter()
        ev.synchronize()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{name:12s} main={mode:5s}: issue {1e3*(t1-t0):6.2f} ms, wait {1e3*(t2-t1):6.2f} ms, main done after {1e3*(time.perf_counter()-t0):6.2f} ms")
