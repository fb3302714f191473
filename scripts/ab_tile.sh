#!/bin/bash
# A/B of the tile-binned mesh forward against the scatter path on one box: mesh parity tests under both, per-kernel times at C2 and C5.
TAG=${1:-r3b}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mesh or mvrenderer or golden or non_square or normalized" > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log | cut -c1-300
for cfg in c2 c5; do for t in 1 0; do echo "== $cfg MVR_MESH_TILED=$t"; MVR_MESH_TILED=$t python scripts/kernel_times.py $cfg 2>&1 | grep -v " 0.0 us" ; done; done 2>&1 | tee $OUT/${TAG}_ab.txt
