set -x
ncu --set full --clock-control none --import-source on -k regex:cheb_eval_wt -s 9 -c 3 -f -o gpurun_out/wt3 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s9_ncu_wt.log 2>&1
ls -la gpurun_out/wt3.ncu-rep
