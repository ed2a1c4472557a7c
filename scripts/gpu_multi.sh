# N GPUs of one box: the default (strong-scaling, 10M rows) bench line; with "c3" also the C3 point (12.5M rows per GPU).
N=$1
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?; tail -8 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
if [ "$2" = "nccl" ] || [ "$3" = "nccl" ]; then
TT_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo rc=$?; tail -3 gpurun_out/bench_n${N}_nccl.err; cat gpurun_out/bench_n${N}_nccl.json
fi
if [ "$2" = "c3" ] || [ "$3" = "c3" ]; then
ROWS=$((12500000 * N))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --rows $ROWS > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo rc=$?; tail -5 gpurun_out/bench_c3_n$N.err; cat gpurun_out/bench_c3_n$N.json
fi
