#!/bin/bash
# Quick A/B on one box: mesh parity tests, then per-kernel device times of the resident C2 mesh step under each
# setting of the profiling knobs given as arguments ("VAR=val VAR2=val" per quoted argument), then the default bench.
#   gpurun --timeout 900 -- 'bash scripts/ab_session.sh <tag> "MVR_SHADE_PPT=1" "MVR_SCATTER_FPC=512 MVR_SHADE_PPT=2"'
TAG=${1:-ab}; shift; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log | cut -c1-300
ab() { echo "== $*"; env MVR_X=0 $* python scripts/kernel_times.py 2>/dev/null | grep "mesh_scatter_kernel\|mesh_shade_kernel \|mesh_backward_kernel \|sum of kernels" | awk '{printf "%s  ", $0} END {print ""}' | sed 's/ per step ([0-9]* launches\/step)//g; s/  */ /g'; }
{
ab ""
for v in "$@"; do ab "$v"; done
ab ""
} 2>&1 | tee $OUT/${TAG}_ab.txt
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_mesh.json 2> $OUT/${TAG}_bench.err; cut -c1-1400 $OUT/${TAG}_bench_mesh.json
