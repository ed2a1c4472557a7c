# N GPUs of one box: the C3 weak-scaling point (12.5M rows per GPU, batch-1) and, with "all", the strong-scaling default line,
# C4 (50M rows, 16k queries, top-100) and C5 (20M leaves, 4 levels, top-200).
N=${1:-8}
run() { name=$1; shift
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" > gpurun_out/bench_${name}_n$N.json 2> gpurun_out/bench_${name}_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/bench_${name}_n$N.json 2> gpurun_out/bench_${name}_n$N.err; fi
  echo "$name rc=$? stdout lines: $(grep -c . gpurun_out/bench_${name}_n$N.json)"; tail -2 gpurun_out/bench_${name}_n$N.err | cut -c1-200; cut -c1-420 gpurun_out/bench_${name}_n$N.json; }
run c3weak --rows $((12500000 * N)) --tag C3-weak --scaling weak --skip batch64,wide,cpu --steps 100 --warmup 5
if [ "$2" = "all" ]; then
run strong --steps 200 --warmup 5
run c4 --rows 50000000 --tag C4 --skip batch64,cpu --steps 100 --warmup 5
run c5 --rows 20000000 --levels 4 --k 200 --tag C5 --skip batch64,wide,cpu --steps 100 --warmup 5
fi
