N=$1; TAG=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
head -c 250 gpurun_out/${TAG}_n${N}.json
