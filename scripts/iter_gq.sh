# one ncu --set full capture per call (a report is 15-30 MB; gpurun copies back 64 MB): usage  bash scripts/iter_gq.sh C4|C5a|C1
mkdir -p gpurun_out
w=${1:-C4}
python bench.py --workload $w --steps 3 --warmup 3 --clusters 100000 --no-cpu-baseline --no-fit > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_fast -s 3 -c 1 -f -o gpurun_out/prof_$w python bench.py --workload $w --steps 3 --warmup 3 --clusters 100000 --no-cpu-baseline --no-fit > gpurun_out/ncu_$w.log 2>&1
ls -la gpurun_out/prof_$w.ncu-rep
