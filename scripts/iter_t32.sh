export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fit.py -x -q -k "(two_pass and negative) or (size_classes and negative) or huge or (larger_than and C3) or (large_cluster) or (em_estep and C3)" 2>&1 | tail -12
for ni in 17 32 33 48 64 128; do
  python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'check', j['check'])"
done
GLMMA_NO_TILE32=1 python bench.py --ni 64 --clusters 62500 --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('generic:', j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'check', j['check'])"
python bench.py --ragged --ni 40 --clusters 100000 --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('ragged40:', j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'check', j['check'])"
GLMMA_NO_TILE32=1 python bench.py --ragged --ni 40 --clusters 100000 --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('ragged40 generic:', j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'check', j['check'])"
