# full session: smoke, tests, benches (with cpu baseline on the default), ncu launch list + captures
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log | cut -c1-200
python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench_mesh.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B --workload points > $OUT/${TAG}_bench_points.json 2>> $OUT/${TAG}_bench.err
$B --batch 8 --views 20 --image-size 400 --faces 100000 > $OUT/${TAG}_bench_c5_mesh.json 2>> $OUT/${TAG}_bench.err
$B --workload points --batch 8 --views 20 --image-size 400 --points 16384 > $OUT/${TAG}_bench_c5_points.json 2>> $OUT/${TAG}_bench.err
for f in bench_mesh bench_points bench_c5_mesh bench_c5_points; do python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$f.json"))
print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "pipe", d["e2e"].get("pipelined",{}).get("value"), d["e2e"].get("list_api",{}).get("value"), d["roofline"]["kernel_ms_all"], d["roofline"]["frac"], d["gpu_launches"], d.get("cpu_baseline",{}).get("value"))
PY
done
cut -c1-400 $OUT/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mesh_ -s 40 -c 7 -o $OUT/${TAG}_mesh -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:points_ -s 40 -c 7 -o $OUT/${TAG}_points -f python bench.py --workload points --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_p.log 2>&1
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -5
