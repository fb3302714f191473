#!/bin/bash
# mesh parity tests + per-kernel times of the scatter path (C2, C5) under the given env settings ("VAR=val ..." per argument)
TAG=${1:-r3x}; shift; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mesh or mvrenderer or golden or non_square or normalized" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log | cut -c1-200
for cfg in c2 c5; do for v in "" "$@"; do echo "== $cfg $v"; env MVR_X=0 $v python scripts/kernel_times.py $cfg 2>&1 | grep "scatter\|shade_kernel \|tile\|bin_kernel\|backward_kernel \|sum of" ; done; done 2>&1 | tee $OUT/${TAG}_ab.txt
