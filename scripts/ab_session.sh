#!/bin/bash
# A/B of the profiling knobs on one box: per-kernel device times of the resident C2 mesh step under each setting.
#   gpurun --timeout 900 -- 'bash scripts/ab_session.sh <tag>'
TAG=${1:-ab}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or mvrenderer or smoke" > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log | cut -c1-200
ab() { label=$1; shift; echo "== $label"; env "$@" python scripts/kernel_times.py 2>/dev/null | grep "mesh_scatter_kernel\|mesh_shade_kernel \|mesh_backward_kernel \|sum of kernels" | awk '{printf "%s  ", $0} END {print ""}' | sed 's/ per step ([0-9]* launches\/step)//g; s/  */ /g'; }
{
ab "default (PPT=4 MINB=4)" MVR_X=0
ab "shade PPT=1" MVR_SHADE_PPT=1
ab "shade PPT=2" MVR_SHADE_PPT=2
ab "shade PPT=4 MINB=3" MVR_SHADE_MINB=3
ab "shade PPT=2 MINB=3" MVR_SHADE_PPT=2 MVR_SHADE_MINB=3
ab "shade PPT=1 MINB=3" MVR_SHADE_PPT=1 MVR_SHADE_MINB=3
ab "scatter FPC=256" MVR_SCATTER_FPC=256
ab "scatter FPC=512" MVR_SCATTER_FPC=512
ab "scatter FPC=2048" MVR_SCATTER_FPC=2048
ab "scatter RUN=4" MVR_SCATTER_RUN=4
ab "scatter RUN=16" MVR_SCATTER_RUN=16
ab "scatter RUN=16 FPC=512" MVR_SCATTER_RUN=16 MVR_SCATTER_FPC=512
ab "bwd UNCOND" MVR_BWD_UNCOND=1
ab "default again" MVR_X=0
} 2>&1 | tee $OUT/${TAG}_ab.txt
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_mesh.json 2> $OUT/${TAG}_bench.err; cut -c1-1500 $OUT/${TAG}_bench_mesh.json
