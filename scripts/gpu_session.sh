#!/bin/bash
# One GPU-box session: parity tests, bench (mesh + points + C5 shapes), e2e host breakdown, ncu launch list + captures.
# usage: scripts/gpu_session.sh <tag> [quick]
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
B="python bench.py --steps 30 --warmup 5"
$B > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
MVR_BWD_MINB=2 $B --no-cpu-baseline > $OUT/${TAG}_bench_bwd2.json 2>> $OUT/${TAG}_bench.err
$B --workload points --no-cpu-baseline > $OUT/${TAG}_bench_points.json 2>> $OUT/${TAG}_bench.err
$B --workload points --points-per-pixel 1 --no-cpu-baseline > $OUT/${TAG}_bench_points_k1.json 2>> $OUT/${TAG}_bench.err
$B --batch 8 --views 20 --image-size 400 --faces 100000 --no-cpu-baseline > $OUT/${TAG}_bench_c5_mesh.json 2>> $OUT/${TAG}_bench.err
$B --workload points --batch 8 --views 20 --image-size 400 --points 16384 --no-cpu-baseline > $OUT/${TAG}_bench_c5_points.json 2>> $OUT/${TAG}_bench.err
python scripts/e2e_breakdown.py > $OUT/${TAG}_e2e_breakdown.txt 2>&1
if [ "$2" != "quick" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_b.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:mesh_ -s 30 -c 5 -o $OUT/${TAG}_mesh -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_m.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:points_ -s 20 -c 4 -o $OUT/${TAG}_points -f python bench.py --workload points --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_p.log 2>&1
fi
for f in bench bench_bwd2 bench_points bench_points_k1 bench_c5_mesh bench_c5_points; do echo "== $f"; cut -c1-1500 $OUT/${TAG}_$f.json; done
tail -5 $OUT/${TAG}_bench.err
