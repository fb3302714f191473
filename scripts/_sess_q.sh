# quick A/B session: parity tests + bench variants selected by env knobs.  usage: scripts/_sess_q.sh <tag>
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log | cut -c1-300
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > $OUT/${TAG}_$name.json 2>> $OUT/${TAG}_bench.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_all"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
EXTRA=""
run base MVR_SHADE_PIX=2
run pix1 MVR_SHADE_PIX=1
run pix2_m2 MVR_SHADE_PIX=2 MVR_SHADE_MINB=2
run pix2_m4 MVR_SHADE_PIX=2 MVR_SHADE_MINB=4
run bwd2 MVR_BWD_MINB=2
EXTRA="--workload points"
run points X=1
EXTRA="--workload points --cuda-graph"
run points_graph X=1
EXTRA="--cuda-graph"
run mesh_graph X=1
grep -v "UserWarning\|run_backward" $OUT/${TAG}_bench.err | tail -8
