mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "censored or C5a or famil or golden or edges" 2>&1 | tail -5
for w in C5a C4; do
  python bench.py --workload $w --clusters 1000000 --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['config']['workload'][:50], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']))"
done
for w in C4 C5a; do
python bench.py --workload $w --steps 3 --warmup 3 --clusters 100000 --no-cpu-baseline --no-fit > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_fast -s 3 -c 1 -f -o gpurun_out/prof_$w python bench.py --workload $w --steps 3 --warmup 3 --clusters 100000 --no-cpu-baseline --no-fit > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out/prof_C4.ncu-rep gpurun_out/prof_C5a.ncu-rep
