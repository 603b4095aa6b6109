export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fit.py -x -q -k "C3 or negbin or negative or large_cluster or ragged or em_" 2>&1 | tail -3
bash scripts/dev_bench.sh
for ni in 12 16; do
  python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:60], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']))"
done
