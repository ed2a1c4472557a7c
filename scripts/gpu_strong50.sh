# Strong scaling at a corpus that still fits one GPU: 50M x 1024 bf16 = 102.4 GB, batch-1, N GPUs.
N=${1:-1}
if [ "$N" = "1" ]; then python bench.py --gpus 1 --rows 50000000 --tag strong-50M --skip batch64,wide,cpu --steps 60 --warmup 5 > gpurun_out/bench_strong50_n$N.json 2> gpurun_out/bench_strong50_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --rows 50000000 --tag strong-50M --skip batch64,wide,cpu --steps 100 --warmup 5 > gpurun_out/bench_strong50_n$N.json 2> gpurun_out/bench_strong50_n$N.err; fi
echo rc=$?; tail -2 gpurun_out/bench_strong50_n$N.err | cut -c1-200; cut -c1-330 gpurun_out/bench_strong50_n$N.json
