#!/bin/bash
# One ncu --set full capture of the dominant kernel per workload (run under gpurun, one GPU):
#   bash profiles/capture.sh <tag> cfg2 cfg3 cfg4
# Each workload first runs plain (must exit 0), then under ncu; reports land in gpurun_out/<tag>_<workload>.ncu-rep.
# Read them here with profiles/analyze.py <rep> <reads in the launch> <summary.json> <workload> and profiles/lines.py.
tag=$1; shift
for w in "$@"; do
  cmd="python bench.py --workload $w --steps 1 --warmup 3 --batches-per-step 1 --no-cpu-baseline --no-north-star"
  skip=3; [ "$w" = cfg4 ] && skip=6   # bulk launches of the 3 warm-up steps (cfg4: two bulk passes per step)
  $cmd > gpurun_out/${tag}_${w}_plain.json 2> gpurun_out/${tag}_${w}_plain.err &&
  ncu --set full --clock-control none --import-source on -k regex:k_place_tally_bulk -s $skip -c 1 -f \
      -o gpurun_out/${tag}_${w} $cmd > gpurun_out/${tag}_${w}_ncu.log 2>&1
  echo "$w rc=$?"
done
