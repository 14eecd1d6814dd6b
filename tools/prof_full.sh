#!/bin/bash
# ncu --set full capture (with source counters) of the big kernels of ONE full-size bench job, exported as text / csv on the GPU box
# (the .ncu-rep files are too large to bring back).  usage: tools/prof_full.sh <tag> <kernel-regex> <skip> <count>
# The launch filter counts only kernels that match the regex; synth staging launches encode_tables / encode_emit once per input image.
set -u
TAG=${1:-r02}; RE=${2:-'block_decode_fused|merge_tiles|encode_emit|encode_tables'}; SKIP=${3:-20}; CNT=${4:-4}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/rep \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu.log 2>&1
ncu -i $OUT/rep.ncu-rep --page details > $OUT/details.txt 2>/dev/null
ncu -i $OUT/rep.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/rep.ncu-rep --page source --csv --print-source cuda,sass > $OUT/source.csv 2>/dev/null || ncu -i $OUT/rep.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
gzip -f $OUT/source.csv $OUT/raw.csv
rm -f $OUT/rep.ncu-rep
ls -la $OUT
