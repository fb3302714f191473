TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
MVR_SHADE_MINB=5 $B > $OUT/${TAG}_bench_s5.json 2>> $OUT/${TAG}_bench.err
MVR_SHADE_MINB=6 $B > $OUT/${TAG}_bench_s6.json 2>> $OUT/${TAG}_bench.err
$B --workload points > $OUT/${TAG}_bench_points.json 2>> $OUT/${TAG}_bench.err
python scripts/e2e_timeline.py > $OUT/${TAG}_e2e_timeline.txt 2>&1
for f in bench bench_s5 bench_s6 bench_points; do echo "== $f"; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$f.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_all"])
PY
done
tail -5 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_e2e_timeline.txt
