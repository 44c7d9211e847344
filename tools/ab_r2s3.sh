set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s3_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
$B > gpurun_out/r2s3_default.json 2> gpurun_out/r2s3_default.err
TBSLAS_TENSOR_CTAS=6 $B > gpurun_out/r2s3_ctas6.json 2>/dev/null
TBSLAS_TENSOR_CTAS=12 $B > gpurun_out/r2s3_ctas12.json 2>/dev/null
TBSLAS_TENSOR_DMMA=1 $B > gpurun_out/r2s3_dmma1.json 2>/dev/null
TBSLAS_B200_LIB=$PWD/tbslas_b200/variants/libtbslas_b200_loc1.so $B > gpurun_out/r2s3_loc1.json 2>/dev/null
TBSLAS_B200_LIB=$PWD/tbslas_b200/variants/libtbslas_b200_xcs.so $B > gpurun_out/r2s3_xcs.json 2>/dev/null
tail -n 3 gpurun_out/r2s3_pytest.log
