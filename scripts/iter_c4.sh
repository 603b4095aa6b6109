mkdir -p gpurun_out
for w in C4; do
python bench.py --workload $w --steps 3 --warmup 3 --clusters 100000 --no-cpu-baseline --no-fit > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_fast -s 3 -c 1 -f -o gpurun_out/prof_$w python bench.py --workload $w --steps 3 --warmup 3 --clusters 100000 --no-cpu-baseline --no-fit > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out/prof_C4.ncu-rep
