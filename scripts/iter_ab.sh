for v in A B; do
export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev$v.so
echo "variant $v"
bash scripts/dev_bench.sh 2>&1 | tail -1
python bench.py --ni 16 --clusters 250000 --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']))"
done
