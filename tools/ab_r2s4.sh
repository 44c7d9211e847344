set -x
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
TBSLAS_TENSOR_CTAS=12 $B > gpurun_out/r2s4_ctas12.json 2>/dev/null
TBSLAS_TENSOR_CTAS=24 $B > gpurun_out/r2s4_ctas24.json 2>/dev/null
TBSLAS_TENSOR_CTAS=600 $B > gpurun_out/r2s4_ctas600.json 2>/dev/null
TBSLAS_TENSOR_CTAS=12 ncu --set full --clock-control none --import-source on -k regex:tensor_grid_dmma2 -s 2 -c 1 -f -o gpurun_out/tensor2 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s4_ncu_tensor.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_kernel -s 9 -c 3 -f -o gpurun_out/locate2 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s4_ncu_locate.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_perm -s 6 -c 1 -f -o gpurun_out/scatter2 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s4_ncu_scatter.log 2>&1
ls -la gpurun_out/*.ncu-rep
