# 8 GPUs of one box: strong-scaling default line, C3 (100M rows), C4 (50M rows, 16k queries, top-100), C5 (20M leaves, 4 levels, top-200)
N=${1:-8}
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/bench_${name}_n$N.json 2> gpurun_out/bench_${name}_n$N.err; echo "$name rc=$?"; tail -3 gpurun_out/bench_${name}_n$N.err | cut -c1-300; cut -c1-600 gpurun_out/bench_${name}_n$N.json; }
run strong --steps 200 --warmup 5
run c4 --rows 50000000 --tag C4 --skip batch64,cpu --steps 100 --warmup 5
run c5 --rows 20000000 --levels 4 --k 200 --tag C5 --skip batch64,wide,cpu --steps 100 --warmup 5
run c3 --rows 100000000 --tag C3 --skip batch64,wide,cpu --steps 100 --warmup 5
