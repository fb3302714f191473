#!/bin/bash
# ncu --set full of the point kernels at the bench workloads (C3: 32 x 12 views of 2048 points, C5: 8 x 20 views of 16 k points) +
# the launch list of a short bench run; summaries are written under profiles/ by scripts/ncu_summary.py on the CPU box.
TAG=${1:-r3z}; OUT=gpurun_out; mkdir -p $OUT
for c in c3 c5; do
  ncu --set full --import-source on --clock-control none -k regex:"points_bin_kernel|points_tile_kernel|points_backward_kernel" -s 4 -c 4 -f -o $OUT/${TAG}_points_$c python scripts/one_step_points.py $c 3 > $OUT/${TAG}_ncu_points_$c.log 2>&1; tail -1 $OUT/${TAG}_ncu_points_$c.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $OUT/${TAG}_b.log 2>&1; tail -1 $OUT/${TAG}_b.log | cut -c1-100
