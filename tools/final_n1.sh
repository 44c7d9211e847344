# usage (on a GPU box): bash tools/final_n1.sh
# the single-GPU evidence of a round: GPU suite, default bench, reference arm, the other four workloads,
# the ncu launch list of the library's kernels, smoke()
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
python bench.py > gpurun_out/r2f_c2.json 2> gpurun_out/r2f_c2.err
python bench.py --impl reference > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
for w in c1 c3 c4 c5; do python bench.py --workload $w --no-cpu > gpurun_out/r2f_$w.json 2> gpurun_out/r2f_$w.err; done
bash tools/launchlist.sh
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
tail -n 2 gpurun_out/r2f_pytest.log gpurun_out/r2f_smoke.log
