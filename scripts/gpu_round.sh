#!/bin/bash
# One gpurun call: a list of stages, each under its own timeout, logs into gpurun_out/.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [stage ...]'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
STAGES=${@:-"kernels e2e bench"}
CASES='ed_wave_epilogue+renoise,ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue(wave2:R1=1),ed_wave_epilogue+rrg'
# the snapshot may carry a library older than the sources (edits while the call was queued): rebuild first (seconds)
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo "[build] FAILED"; tail -20 gpurun_out/build.log; }
for s in $STAGES; do
  t0=$(date +%s)
  case $s in
    gputests) timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; rc=$? ;;
    kernels) timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/t_kernels.log 2>&1; rc=$? ;;
    peer) timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k peer > gpurun_out/t_peer.log 2>&1; rc=$? ;;
    roofline) timeout 240 python bench.py --roofline-only > gpurun_out/roofline.json 2> gpurun_out/roofline.err; rc=$? ;;
    ncu) timeout 300 ncu --set full --clock-control none --import-source on -k regex:wave_epilogue -c 6 -f \
           -o gpurun_out/r2_epilogue python bench.py --roofline-only --roofline-iters 1 --roofline-warm 0 \
           --roofline-cases "$CASES" > gpurun_out/ncu.log 2>&1; rc=$? ;;
    e2e) timeout 480 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_zz_staged_in_pipeline.py -x -q > gpurun_out/t_e2e.log 2>&1; rc=$? ;;
    bench) timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; rc=$? ;;
    benchdrv) timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; rc=$? ;;
    bench5) timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5.err; rc=$? ;;
    bench4) timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4.err; rc=$? ;;
    bench2) timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_cfg2_n1.json 2> gpurun_out/bench_cfg2.err; rc=$? ;;
    minbx) for m in 5 7; do ED_NVCC_FLAGS="-DED_HALF_MINB=$m" python -c "import importlib; importlib.import_module('elasticdiffusion-official_b200').native.build(force=True)" && \
           timeout 240 python bench.py --roofline-only --roofline-iters 30 --roofline-cases 'ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue(wave2:R1=1)' > gpurun_out/roofline_minb$m.json 2> gpurun_out/roofline_minb$m.err; cat gpurun_out/roofline_minb$m.json; done; rc=$?; \
           python -c "import importlib; importlib.import_module('elasticdiffusion-official_b200').native.build(force=True)"; \
           timeout 240 python bench.py --roofline-only --roofline-iters 30 --roofline-cases 'ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue(wave2:R1=1)' > gpurun_out/roofline_minb6.json; cat gpurun_out/roofline_minb6.json ;;
    minb8) ED_NVCC_FLAGS="-DED_HALF_MINB=8" python -c "import importlib; importlib.import_module('elasticdiffusion-official_b200').native.build(force=True)" && \
           timeout 240 python bench.py --roofline-only --roofline-cases 'ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue(wave2:R1=1)' > gpurun_out/roofline_minb8.json 2> gpurun_out/roofline_minb8.err; rc=$?; \
           python -c "import importlib; importlib.import_module('elasticdiffusion-official_b200').native.build(force=True)" ;;
    minbm) for m in 5 6; do ED_NVCC_FLAGS="-DED_HALF_MINB_MULTI=$m" python -c "import importlib; importlib.import_module('elasticdiffusion-official_b200').native.build(force=True)" && \
           timeout 240 python bench.py --roofline-only --roofline-cases 'ed_wave_epilogue+rrg' > gpurun_out/roofline_minbm$m.json 2> gpurun_out/roofline_minbm$m.err; cat gpurun_out/roofline_minbm$m.json; done; rc=$?; \
           python -c "import importlib; importlib.import_module('elasticdiffusion-official_b200').native.build(force=True)" ;;
    benchref) timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; rc=$? ;;
    probefusedq) timeout 600 python scripts/probe_unet.py --fused-sweep --quick > gpurun_out/probe_unet_fused.log 2>&1; rc=$?; grep -v Warn gpurun_out/probe_unet_fused.log | tail -6 | cut -c1-400 ;;
    probefused) timeout 600 python scripts/probe_unet.py --fused-sweep > gpurun_out/probe_unet_fused.log 2>&1; rc=$?; tail -3 gpurun_out/probe_unet_fused.log ;;
    unetops) timeout 600 python -m pytest tests/test_gpu_unet_ops.py -x -q > gpurun_out/t_unetops.log 2>&1; rc=$?; tail -15 gpurun_out/t_unetops.log ;;
    probe) timeout 600 python scripts/probe_unet.py gpurun_out/probe_unet.json > gpurun_out/probe_unet.log 2>&1; rc=$? ;;
    launches) BENCH_GRAPHS=0 BENCH_CUPROF=1 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 5000 --csv \
           --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-extras --no-parity > gpurun_out/launches.log 2>&1; rc=$? ;;
    multi) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29533 \
           scripts/check_multi_gpu.py > gpurun_out/multi_gpu_n${NGPU:-2}.log 2>&1; rc=$? ;;
    scale) for n in ${SCALE_NS:-"2 4 8"}; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
             --master-port 29544 bench.py --gpus $n --steps 8 --warmup 3 --no-extras > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; done; rc=$? ;;
    scale4) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29555 \
             bench.py --gpus ${NGPU:-8} --workload cfg4 --steps 8 --warmup 3 --no-extras > gpurun_out/bench_cfg4_n${NGPU:-8}.json 2> gpurun_out/bench_cfg4_n${NGPU:-8}.err; rc=$? ;;
    scale5) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29566 \
             bench.py --gpus 4 --workload cfg5 --steps 8 --warmup 3 --no-extras > gpurun_out/bench_cfg5_n4.json 2> gpurun_out/bench_cfg5_n4.err; rc=$? ;;
    smoke) timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; rc=$? ;;
    *) echo "unknown stage $s"; rc=99 ;;
  esac
  echo "[stage $s] rc=$rc $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/stages.log
done
for f in t_gpu t_kernels t_peer t_e2e; do [ -f gpurun_out/$f.log ] && tail -4 gpurun_out/$f.log; done
[ -f gpurun_out/probe_unet.log ] && tail -12 gpurun_out/probe_unet.log
[ -f gpurun_out/bench_n1.json ] && head -c 1500 gpurun_out/bench_n1.json
[ -f gpurun_out/bench.err ] && tail -5 gpurun_out/bench.err
[ -f gpurun_out/bench_ref.json ] && head -c 600 gpurun_out/bench_ref.json
for f in gpurun_out/bench_n*.json gpurun_out/bench_cfg*_n*.json; do [ -f $f ] && { echo "== $f"; tail -1 $f | head -c 400; echo; }; done
for f in gpurun_out/multi_gpu_n*.log; do [ -f $f ] && tail -12 $f; done
for f in bench_cfg5 bench_cfg4 bench_cfg2; do [ -f gpurun_out/$f.err ] && tail -3 gpurun_out/$f.err; done
[ -f gpurun_out/roofline_minb8.json ] && cat gpurun_out/roofline_minb8.json
true
