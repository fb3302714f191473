TAG=$1; N=$2; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_smi.txt 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2>> $OUT/${TAG}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --workload points > $OUT/${TAG}_bench_points_n$N.json 2>> $OUT/${TAG}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref_n$N.json 2>> $OUT/${TAG}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 examples/train_step.py > $OUT/${TAG}_train_step_n$N.log 2>&1
for f in bench_n1 bench_n$N bench_points_n$N bench_ref_n$N; do echo "== $f"; cut -c1-900 $OUT/${TAG}_$f.json; done
tail -5 $OUT/${TAG}_bench.err; tail -5 $OUT/${TAG}_train_step_n$N.log
