#!/bin/bash
# One gpurun call: kernel parity, roofline A/B, ncu capture of the staged epilogue, e2e parity, bench line.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [stage ...]'   (default: all stages)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
STAGES=${@:-"kernels roofline ncu e2e bench"}
CASES='ed_wave_epilogue+renoise,ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue(wave2:R1=1),ed_wave_epilogue+rrg'
for s in $STAGES; do
  t0=$(date +%s)
  case $s in
    kernels) timeout 420 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/t_kernels.log 2>&1; rc=$? ;;
    roofline) timeout 240 python bench.py --roofline-only > gpurun_out/roofline.json 2> gpurun_out/roofline.err; rc=$? ;;
    ncu) timeout 300 ncu --set full --clock-control none --import-source on -k regex:wave_epilogue_staged -c 4 -f \
           -o gpurun_out/r1_epi_staged python bench.py --roofline-only --roofline-iters 1 --roofline-warm 0 \
           --roofline-cases "$CASES" > gpurun_out/ncu.log 2>&1; rc=$? ;;
    e2e) timeout 480 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/t_e2e.log 2>&1; rc=$? ;;
    bench) timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; rc=$? ;;
    launches) BENCH_GRAPHS=0 BENCH_CUPROF=1 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 5000 --csv \
           --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/launches.log 2>&1; rc=$? ;;
    smoke) timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; rc=$? ;;
    *) echo "unknown stage $s"; rc=99 ;;
  esac
  echo "[stage $s] rc=$rc $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/stages.log
done
tail -3 gpurun_out/t_kernels.log 2>/dev/null
tail -3 gpurun_out/t_e2e.log 2>/dev/null
head -c 3000 gpurun_out/roofline.json 2>/dev/null
