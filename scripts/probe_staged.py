"""GPU probe (debug): run single staged-epilogue cases in fresh processes; argv: case indices | 'san' to wrap in compute-sanitizer."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
CASES = [((2, 192, 192, 128, (128, 128), 64), "float32", 4), ((1, 192, 192, 64, (64, 64), 32), "float32", 3),
         ((1, 128, 256, 128, (64, 128), 64), "float32", 4), ((2, 192, 192, 128, (128, 128), 64), "bfloat16", 4),
         ((1, 192, 192, 128, (128, 128), 64), "float32", 1)]
if len(sys.argv) > 2 and sys.argv[1] == "one":
    import torch
    import test_gpu_kernels as t
    cfg, dt, R1 = CASES[int(sys.argv[2])]
    L = t.native.lib()
    t.native.check(L.ed_set_epilogue_mode(2))
    t.test_wave_epilogue_matches_spec(cfg, "plain", getattr(torch, dt), "staged", R1=R1)
    print("PASS", cfg, dt, R1)
    sys.exit(0)
CASES.append(((1, 72, 100, 64, (36, 50), 32), "float32", 3))      # g_lp = 7: unaligned start, no overhang
CASES.append(((1, 72, 100, 64, (36, 50), 32), "bfloat16", 3))
for i in (0, 1, 5, 6):
  for origin in (0, 1, 2, 3):
    env = dict(os.environ, ED_STAGED_ORIGIN=str(origin))
    for tag, pre in ((f"origin={origin}", []),):
        r = subprocess.run(pre + [sys.executable, __file__, "one", str(i)], capture_output=True, text=True, env=env, timeout=300)
        out = (r.stdout + r.stderr)
        keep = [l for l in out.splitlines() if any(k in l for k in ("PASS", "Error", "error", "=========", "trap", "Illegal", "at ", "diff"))]
        print(f"### case {i} {CASES[i]} [{tag}] rc={r.returncode}")
        print("\n".join(keep[:3]))
