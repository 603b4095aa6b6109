export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev.so
for t in 128 256 384 512 0; do
if [ $t != 0 ]; then export GLMMA_FAST_THREADS=$t; else unset GLMMA_FAST_THREADS; fi
echo "threads $t"
bash scripts/dev_bench.sh 2>&1 | tail -1
done
