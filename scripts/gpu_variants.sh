#!/bin/bash
# debug: A/B of staged-epilogue launch geometry (env overrides read by the launcher)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES='ed_wave_epilogue+renoise,ed_wave_epilogue+rrg(wave2:R1=1),ed_wave_epilogue(wave2:R1=1),ed_wave_epilogue+rrg'
for th in 256 128 64; do for cpt in 2 4; do
  echo "threads=$th cpt=$cpt: $(ED_STAGED_THREADS=$th ED_STAGED_CPT=$cpt python bench.py --roofline-only --roofline-cases "$CASES" 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['roofline_all']
print({k.replace('ed_wave_epilogue',''):v['ms'] for k,v in d.items()})")"
done; done
