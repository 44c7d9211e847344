python -m pytest tests -m gpu -x -q > gpurun_out/r2s17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s17_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
$B > gpurun_out/r2s17_c2.json 2>/dev/null
python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s17_c5.json 2>/dev/null
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r2s17_c3.json 2>/dev/null
python bench.py --workload c1 --tensor-grid always --steps 20 --warmup 5 --no-cpu --no-parity > gpurun_out/r2s17_c1_always.json 2>/dev/null
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_grid_velocity" > gpurun_out/r2s17_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2s17_racecheck.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor or insitu" > gpurun_out/r2s17_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2s17_memcheck.log
tail -n 3 gpurun_out/r2s17_pytest.log gpurun_out/r2s17_racecheck.log gpurun_out/r2s17_memcheck.log
