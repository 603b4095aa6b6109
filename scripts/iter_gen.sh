for w in C5a C5b C4 C1; do
for ni in 16 24; do
python bench.py --workload $w --ni $ni --clusters $((2400000 / ni)) --steps 3 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:70], '| %.3f ms'%(j['ms_per_step']))"
done
done
