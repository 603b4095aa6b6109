mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fit > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_dev.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fit > gpurun_out/ncu3.log 2>&1
python bench.py --steps 3 --warmup 3 --clusters 200000 --no-cpu-baseline --no-fit > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_fast -s 3 -c 1 -f -o gpurun_out/prof_dev python bench.py --steps 3 --warmup 3 --clusters 200000 --no-cpu-baseline --no-fit > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/prof_dev.ncu-rep
