#!/bin/bash
# Quick GPU-box session: parity tests + the bench lines quoted in README / DESIGN (no CPU baseline, no ncu).
#   gpurun --timeout 900 -- 'bash scripts/gpu_session.sh <tag>'        (scripts/gpu_session_full.sh adds the CPU
#   baseline, the reference arm, the ncu launch list and the --set full captures summarised under profiles/)
TAG=${1:-rX}; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log | cut -c1-200
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > $OUT/${TAG}_$name.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "fwd", d.get("forward_only", {}).get("value"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
          d["e2e"].get("list_api", {}).get("value"), d["roofline"]["kernel_ms_all"], d["roofline"]["frac"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run mesh
run points --workload points
run points_graph --workload points --cuda-graph
run points_k1 --workload points --points-per-pixel 1
run c1 --workload points --batch 1 --points-per-pixel 1
run c1_graph --workload points --batch 1 --points-per-pixel 1 --cuda-graph
run c5_mesh --batch 8 --views 20 --image-size 400 --faces 100000
run c5_points --workload points --batch 8 --views 20 --image-size 400 --points 16384
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -5
