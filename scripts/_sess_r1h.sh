TAG=r1h; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
$B --workload points > $OUT/${TAG}_bench_points.json 2>> $OUT/${TAG}_bench.err
$B --workload points --points-per-pixel 1 > $OUT/${TAG}_bench_points_k1.json 2>> $OUT/${TAG}_bench.err
$B --batch 8 --views 20 --image-size 400 --faces 100000 > $OUT/${TAG}_bench_c5_mesh.json 2>> $OUT/${TAG}_bench.err
$B --workload points --batch 8 --views 20 --image-size 400 --points 16384 > $OUT/${TAG}_bench_c5_points.json 2>> $OUT/${TAG}_bench.err
python scripts/e2e_profile.py > $OUT/${TAG}_e2e_profile.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:points_scatter -s 6 -c 1 -o $OUT/${TAG}_c5_points_scatter -f python bench.py --workload points --batch 8 --views 20 --image-size 400 --points 16384 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mesh_scatter -s 6 -c 1 -o $OUT/${TAG}_mesh_scatter -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_m.log 2>&1
for f in bench bench_points bench_points_k1 bench_c5_mesh bench_c5_points; do echo "== $f"; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$f.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_all"])
PY
done
tail -5 $OUT/${TAG}_bench.err; head -3 $OUT/${TAG}_e2e_profile.txt
