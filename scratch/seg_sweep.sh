#!/bin/bash
# Segment-length x start sweep for the headline recipe: SER gate and Msamples/s per point.
# Run on the GPU box:  bash scratch/seg_sweep.sh > gpurun_out/seg_sweep.jsonl
for seg in 8454 32768 131072 262144; do
  for start in cold stream acquire; do
    python bench.py --seg $seg --start $start --no-cpu-baseline --no-c5 --no-dropin --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
w=d.get('withheld') or {}
print(json.dumps({'seg':$seg,'start':'$start','value':d['value'],'withheld_value':w.get('value'),'ms_per_step':d['ms_per_step'],'ser':d['sanity']['ser'],'ser_max_segment':d['sanity']['ser_max_segment'],'segments':d['sanity']['segments'],'acquisition_ms':(d.get('acquisition') or {}).get('ms'),'stage_ms':d['roofline']['stage_ms_per_step']}))"
  done
done
