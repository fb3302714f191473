# session: parity tests + bench (mesh, points, C1, C5 points) + ncu of the point kernels.  usage: scripts/_sess_r.sh <tag>
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log | cut -c1-300
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
run c1_graph --workload points --batch 1 --points-per-pixel 1 --cuda-graph
run c5_points --workload points --batch 8 --views 20 --image-size 400 --points 16384
python scripts/e2e_host_profile.py > $OUT/${TAG}_e2e_host.txt 2>&1; tail -10 $OUT/${TAG}_e2e_host.txt
ncu --set full --clock-control none --import-source on -k regex:points_ -s 20 -c 6 -o $OUT/${TAG}_points -f python bench.py --workload points --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_p.log 2>&1
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -8
