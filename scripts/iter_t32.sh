export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fit.py -x -q -k "(C3 or negbin or negative or large_cluster or ragged or huge or size_classes or larger_than) and not zi and not censored and not C1 and not C2 and not C4 and not C5" 2>&1 | tail -12
for ni in 16 17 24 32 64; do
  python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'check', j['check'])"
done
GLMMA_NO_TILE32=1 python bench.py --ni 24 --clusters 166666 --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('generic:', j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'check', j['check'])"
