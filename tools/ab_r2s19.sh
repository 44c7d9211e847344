python -m pytest tests -m gpu -x -q -k "virtual or insitu or tensor" > gpurun_out/r2s19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s19_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
$B > gpurun_out/r2s19_real.json 2>/dev/null
TBSLAS_VIRTUAL_X=1 $B > gpurun_out/r2s19_virt.json 2>/dev/null
TBSLAS_VIRTUAL_X=1 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s19_virt_c5.json 2>/dev/null
tail -n 2 gpurun_out/r2s19_pytest.log
