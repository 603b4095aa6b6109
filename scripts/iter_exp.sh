for lib in libglmma.so libglmma_exp.so; do
export GLMMA_LIB=$PWD/glmmadaptive_b200/$lib
echo $lib
python bench.py --workload C4 --clusters 1000000 --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:50], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']))"
done
python scripts/family_sweep.py 200000 2>&1 | grep "^zi.poisson"
