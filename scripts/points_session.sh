#!/bin/bash
# Quick point-path session: all GPU tests, then the point benches (C3 K=4 alpha, K=1, C1, C5).
TAG=${1:-pt}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log | cut -c1-300
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload points"
run() { name=$1; shift; $B "$@" > $OUT/${TAG}_$name.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "fwd", d.get("forward_only", {}).get("value"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
          d["e2e"].get("pipelined", {}).get("value"), d["roofline"]["kernel_ms_all"], d["roofline"]["frac"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run points
run points_k1 --points-per-pixel 1
run c1 --batch 1 --points-per-pixel 1
run c5_points --batch 8 --views 20 --image-size 400 --points 16384
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -5
run points_graph --cuda-graph
run c1_graph --batch 1 --points-per-pixel 1 --cuda-graph
