B="python bench.py --steps 5 --warmup 3 --no-cpu --no-parity"
for c in 2 4 6 12 16 32; do
TBSLAS_TENSOR_DMMA=3 TBSLAS_TENSOR_CTAS=$c $B > gpurun_out/r2s16_c$c.json 2>/dev/null
done
