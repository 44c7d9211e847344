python -m pytest tests -m gpu -x -q > gpurun_out/r2s18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s18_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
TBSLAS_B200_LIB=$PWD/tbslas_b200/variants/libtbslas_b200_O1.so $B > gpurun_out/r2s18_O1.json 2>/dev/null
TBSLAS_B200_LIB=$PWD/tbslas_b200/variants/libtbslas_b200_O1.so python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu --no-parity > gpurun_out/r2s18_O1_c1.json 2>/dev/null
tail -n 2 gpurun_out/r2s18_pytest.log
