TBSLAS_TENSOR_DMMA=3 python -m pytest tests -m gpu -x -q -k "tensor or insitu or config or ns_call or dropin" > gpurun_out/r2s15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s15_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
$B > gpurun_out/r2s15_m2.json 2>/dev/null
TBSLAS_TENSOR_DMMA=3 $B > gpurun_out/r2s15_m3.json 2>/dev/null
TBSLAS_TENSOR_DMMA=3 TBSLAS_TENSOR_CTAS=8 $B > gpurun_out/r2s15_m3c8.json 2>/dev/null
TBSLAS_TENSOR_DMMA=3 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s15_m3_c5.json 2>/dev/null
tail -n 3 gpurun_out/r2s15_pytest.log
