B="python bench.py --steps 4 --warmup 3 --no-cpu --no-parity"
for v in pair f1p0 f1p1 f2p0 f2p1; do
TBSLAS_B200_LIB=$PWD/tbslas_b200/variants/libtbslas_b200_$v.so $B > gpurun_out/r2s14_$v.json 2>/dev/null
done
