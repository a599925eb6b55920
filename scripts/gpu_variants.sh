#!/bin/bash
# debug: A/B of staged-epilogue build variants (ED_STAGED_MINB) + channels_last UNet on the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES='ed_wave_epilogue+renoise,ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue+rrg'
for mb in 2 3 4; do
  cp build/variants/libelastic_b200.mb$mb.so elasticdiffusion-official_b200/libelastic_b200.so
  python bench.py --roofline-only --roofline-cases "$CASES" > gpurun_out/roof_mb$mb.json 2> gpurun_out/roof_mb$mb.err
  echo "mb$mb: $(cat gpurun_out/roof_mb$mb.json)"
done
cp build/variants/libelastic_b200.mb2.so elasticdiffusion-official_b200/libelastic_b200.so
BENCH_CL=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/bench_cl.json 2> gpurun_out/bench_cl.err
echo "CL: $(cut -c1-220 gpurun_out/bench_cl.json)"
timeout 300 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/bench_nocl.json 2> gpurun_out/bench_nocl.err
echo "noCL: $(cut -c1-220 gpurun_out/bench_nocl.json)"
python -c "
import json
for f in ('gpurun_out/bench_cl.json','gpurun_out/bench_nocl.json'):
    d=json.load(open(f)); print(f, d['value'], d['kernels_in_step'])
"
