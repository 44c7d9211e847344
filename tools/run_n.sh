# usage (on a GPU box): bash tools/run_n.sh N tag
# the multi-GPU parity check (tests/test_multigpu.py) and the bench under torchrun on N GPUs of one node
N=$1; TAG=$2
python -m pytest tests/test_multigpu.py -x -q > gpurun_out/${TAG}_mg${N}.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_mg${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
tail -n 2 gpurun_out/${TAG}_mg${N}.log; head -c 300 gpurun_out/${TAG}_n${N}.json
