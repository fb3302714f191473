TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -15 $OUT/${TAG}_pytest.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
$B --cuda-graph > $OUT/${TAG}_bench_graph.json 2>> $OUT/${TAG}_bench.err
$B --workload points > $OUT/${TAG}_bench_points.json 2>> $OUT/${TAG}_bench.err
$B --workload points --cuda-graph > $OUT/${TAG}_bench_points_graph.json 2>> $OUT/${TAG}_bench.err
$B --workload points --points-per-pixel 1 --cuda-graph > $OUT/${TAG}_bench_points_k1_graph.json 2>> $OUT/${TAG}_bench.err
for f in bench bench_graph bench_points bench_points_graph bench_points_k1_graph; do echo "== $f"; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$f.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_all"], d["gpu_launches"])
PY
done
tail -8 $OUT/${TAG}_bench.err
