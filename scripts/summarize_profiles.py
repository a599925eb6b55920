"""Turn what a `bash scripts/gpu_round.sh` call left in gpurun_out/ into the tracked evidence under profiles/:
  profiles/<tag>_bench_n1.json        the bench line
  profiles/<tag>_roofline.json        L2-exceeding kernel rooflines incl. the A/B against the direct epilogue kernel
  profiles/<tag>_epi_staged.ncu-rep   `ncu --set full` capture of the staged epilogue (4 launches: +renoise, +rrg wave 2,
                                      plain wave 2, +rrg R1=8)
  profiles/<tag>_ncu_summary.md       headline metrics of those launches
  profiles/<tag>_launches_summary.md  per-kernel shares of the ncu launch list of the bench command
  profiles/traffic.json               DRAM bytes per launch that bench.py reports as roofline.traffic
Usage: python scripts/summarize_profiles.py <tag>      (e.g. r1b)"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2d"
REP = sys.argv[2] if len(sys.argv) > 2 else "r2_epilogue.ncu-rep"     # name of the ncu report in gpurun_out/
CASES = ["ed_wave_epilogue+renoise", "ed_wave_epilogue+rrg(wave2:R1=1)", "ed_wave_epilogue(wave2:R1=1)", "ed_wave_epilogue+rrg"]
METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
           ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
           ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA"),
           ("smsp__inst_executed.sum", "warp instructions"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
           ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
           ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue")]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    for src, dst in (("bench_n1.json", f"{tag}_bench_n1.json"), ("roofline.json", f"{tag}_roofline.json"),
                     (REP, f"{tag}_epilogue.ncu-rep")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    roof = json.load(open(os.path.join(OUT, "roofline.json")))["roofline_all"]
    # ---- ncu --set full summary -------------------------------------------------------------------------------------
    rep = os.path.join(OUT, REP)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    traffic = {}
    md = [f"# ncu --set full summaries ({tag}): the fused wave epilogue at the L2-exceeding roofline sizes (B=96 SDXL latents)", "",
          "Captured with `ncu --set full --clock-control none --import-source on -k regex:wave_epilogue -c 6` around",
          "`python bench.py --roofline-only --roofline-iters 1 --roofline-warm 0 --roofline-cases <the four epilogue cases>`",
          f"(scripts/gpu_round.sh, stage `ncu`); raw report: `profiles/{tag}_epilogue.ncu-rep`.  AUTO kernel selection: the launch with a",
          "noise stream runs the TMA tile-staged kernel, the three without run the half kernels (template arguments <dtype, MULTI, PEER>).",
          "Durations under ncu are cold-cache single launches; the timed numbers are the CUDA-event ones in the roofline table.",
          "`instructions / element` = smsp__inst_executed.sum (warp level) x 32 lanes / (96 x 4 x 128 x 256 output elements).", ""]
    for case, d in zip(CASES, data):
        g = lambda m: (d[hdr.index(m)], units[hdr.index(m)]) if m in hdr else ("n/a", "")
        rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
        traffic[case] = {"bytes": rd + wr, "source": f"profiles/{tag}_epilogue.ncu-rep"}
        alg = roof[case]["algorithmic_MB"] * 1e6
        md += [f"## {case} -> `{d[hdr.index('Kernel Name')][:70]}`", "",
               f"algorithmic {alg / 1e6:.1f} MB; CUDA-event time {roof[case]['ms'] * 1e3:.1f} us = {roof[case]['GB/s']:.0f} GB/s = "
               f"{roof[case]['frac']:.2f} of the measured HBM peak", "", "| metric | value |", "|---|---|"]
        for m, label in METRICS:
            v, u = g(m)
            md.append(f"| {label} (`{m}`) | {v} {u} |")
        try:
            ipe = float(g("smsp__inst_executed.sum")[0].replace(",", "")) * 32 / (96 * 4 * 128 * 256)
            md.append(f"| instructions / element (thread level) | {ipe:.0f} |")
        except ValueError:
            pass
        md += [f"| **traffic (DRAM read+write)** | {(rd + wr) / 1e6:.1f} MB = {(rd + wr) / alg:.2f} x algorithmic bytes |", ""]
    open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w").write("\n".join(md))
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    # ---- launch list ----------------------------------------------------------------------------------------------------
    lpath = os.path.join(OUT, "launches.csv")
    if os.path.exists(lpath):
        lines = [l for l in open(lpath, errors="replace") if l.startswith('"')]
        rows = list(csv.reader(lines))
        hdr = rows[0]
        ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            if len(r) <= iv or r[im] != "gpu__time_duration.sum":
                continue
            a = agg.setdefault(r[ik], [0, 0.0])
            a[0] += 1
            a[1] += float(r[iv].replace(",", ""))        # ns
        tot = sum(a[1] for a in agg.values())
        n = sum(a[0] for a in agg.values())
        ours = {k: v for k, v in agg.items() if k.startswith("ed::") or "ed::" in k or "wave_epilogue" in k or "tma_box" in k
                or "pick_gather" in k or "owner_map" in k or "gather_views" in k or "geglu_kernel" in k or "gn_stats" in k
                or "gn_apply" in k}
        md = [f"# ncu launch list ({tag}) (`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 5000`)", "",
              "Command: `BENCH_GRAPHS=0 BENCH_CUPROF=1 python bench.py --steps 1 --warmup 3 --no-extras --no-parity` (cfg3, 1 GPU, CUDA graphs",
              "off so that every kernel is a separate launch; cudaProfilerStart/Stop bracket the timed step, so model initialisation and",
              "warm-up are not captured; fused UNet ops on).  The capture was cut by the stage's time limit after the launches listed here",
              "(steady-state: most of one denoise step).  Per-launch times under ncu are cold-cache and serialised: compare SHARES, not",
              "absolutes.",
              "", f"total captured: {n} launches, {tot / 1e6:.1f} ms", "", "## libelastic_b200 kernels", "",
              "| kernel | launches | total us | avg us | share of captured GPU time |", "|---|---|---|---|---|"]
        for k, (c, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| `{k[:90]}` | {c} | {t / 1e3:.1f} | {t / 1e3 / c:.2f} | {100 * t / tot:.3f} % |")
        ot = sum(v[1] for v in ours.values())
        md += ["", f"All libelastic_b200 kernels together: {ot / 1e3:.0f} us = {100 * ot / tot:.2f} % of the captured GPU time; the rest is the "
               "UNet (PyTorch: cuDNN conv / cuBLAS GEMM / cuDNN attention / elementwise) and torch's RNG kernels - the step is UNet-bound.  "
               "`geglu_kernel` / `gn_stats_kernel` / `gn_apply_kernel` are the library's opt-in fused ops INSIDE the UNet forward.",
               "", "## top 15 kernels overall", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
            md.append(f"| `{k[:90]}` | {c} | {t / 1e6:.2f} | {100 * t / tot:.1f} % |")
        open(os.path.join(PROF, f"{tag}_launches_summary.md"), "w").write("\n".join(md) + "\n")
    print("wrote", [f for f in sorted(os.listdir(PROF)) if f.startswith(tag)] + ["traffic.json"])


if __name__ == "__main__":
    main()
