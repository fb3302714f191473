OUT=gpurun_out; mkdir -p $OUT
B="--steps 30 --warmup 5 --no-cpu-baseline"
for i in 1 2 3; do
  for v in old new; do
    if [ $v = old ]; then d=_ab_old; else d=.; fi
    (cd $d && python bench.py $B 2>/dev/null) > $OUT/ab_${v}_$i.json
    python - <<PY
import json
d=json.load(open("$OUT/ab_${v}_$i.json"))
print("$v $i", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_all"])
PY
  done
done
