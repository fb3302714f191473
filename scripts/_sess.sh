TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log | cut -c1-300
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
MVR_BWD_MINB=4 $B > $OUT/${TAG}_bench_bwd4.json 2>> $OUT/${TAG}_bench.err
$B --workload points > $OUT/${TAG}_bench_points.json 2>> $OUT/${TAG}_bench.err
$B --workload points --cuda-graph > $OUT/${TAG}_bench_points_graph.json 2>> $OUT/${TAG}_bench.err
$B --workload points --points-per-pixel 1 > $OUT/${TAG}_bench_points_k1.json 2>> $OUT/${TAG}_bench.err
$B --workload points --batch 8 --views 20 --image-size 400 --points 16384 > $OUT/${TAG}_bench_c5_points.json 2>> $OUT/${TAG}_bench.err
MVR_POINTS_TILED=0 $B --workload points --batch 8 --views 20 --image-size 400 --points 16384 > $OUT/${TAG}_bench_c5_points_untiled.json 2>> $OUT/${TAG}_bench.err
for f in bench bench_bwd4 bench_points bench_points_graph bench_points_k1 bench_c5_points bench_c5_points_untiled; do echo "== $f"; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$f.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_all"], d["gpu_launches"])
PY
done
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -8
