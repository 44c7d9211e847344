set -x
python -m pytest tests -m gpu -x -q -k "tensor or insitu or config or ns_call or dropin" > gpurun_out/r2s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s7_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
$B > gpurun_out/r2s7_default.json 2>/dev/null
TBSLAS_TENSOR_CTAS=24 $B > gpurun_out/r2s7_ctas24.json 2>/dev/null
TBSLAS_TENSOR_CTAS=6 $B > gpurun_out/r2s7_ctas6.json 2>/dev/null
tail -n 3 gpurun_out/r2s7_pytest.log
