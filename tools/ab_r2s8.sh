set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s8_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu"
$B > gpurun_out/r2s8_default.json 2>/dev/null
TBSLAS_B200_LIB=$PWD/tbslas_b200/variants/libtbslas_b200_p3.so $B --no-parity > gpurun_out/r2s8_p3.json 2>/dev/null
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2s8_c1.json 2>/dev/null
python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2s8_c5.json 2>/dev/null
./tools/evalbench 2>&1 | head -14 > gpurun_out/r2s8_evalbench.txt
tail -n 3 gpurun_out/r2s8_pytest.log
