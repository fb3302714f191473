TAG=$1; N=$2; OUT=gpurun_out; mkdir -p $OUT
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 30 --warmup 5 > $OUT/${TAG}_scale_n$n.json 2>> $OUT/${TAG}_scale.err; fi
    python - <<PY
import json
d=json.load(open("$OUT/${TAG}_scale_n$n.json"))
print("N=$n", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("list_api",{}).get("value"), d["clocks"])
PY
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 30 --warmup 5 --workload points > $OUT/${TAG}_scale_points_n$N.json 2>> $OUT/${TAG}_scale.err
cut -c1-400 $OUT/${TAG}_scale_points_n$N.json
grep -v "^\*\|Setting OMP\|^$\|NCCL version" $OUT/${TAG}_scale.err | tail -5
