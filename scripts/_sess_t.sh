TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log | cut -c1-300
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > $OUT/${TAG}_$name.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("list_api",{}).get("value"), d["roofline"]["kernel_ms_all"], d["roofline"]["frac"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run mesh
run points --workload points
run points_graph --workload points --cuda-graph
run points_k1 --workload points --points-per-pixel 1
run c1 --workload points --batch 1 --points-per-pixel 1
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -8
