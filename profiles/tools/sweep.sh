#!/bin/bash
# Device-resident throughput of every preset / BASELINE chain / single block on one GPU.
# usage: profiles/tools/sweep.sh OUTDIR      (one bench line per chain in OUTDIR/<name>.json)
out=${1:-gpurun_out/sweep}; mkdir -p "$out"
run() { name=$(echo "$1" | tr ' ' '_'); shift; python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | tail -1 > "$out/$name.json"
        python -c "
import json,sys
d=json.load(open('$out/$name.json')); print('%-28s %10.0f Ms/s  frac %.3f  %s' % ('$name', d['value'], d['roofline']['frac'], d['roofline']['kernel'][:60]))"; }
for p in "Rain Delay" "Slapback Echo" "Robot Voice" "Cathedral" "Guitar Filter" c3-dist-octaver-delay c2-biquad-cascade \
         block-delay block-reverb block-filter block-gate block-octaver block-distortion; do run "$p" --steps 10 --preset "$p"; done
run "Clean Noise Removal" --steps 5 --preset "Clean Noise Removal" --clips 296
run c2-one-60s-clip --steps 20 --preset c2-biquad-cascade --clips 1 --seconds 60
run c4-convreverb --steps 5 --preset c4-convreverb --clips 256 --seconds 30
