# round-2 measurement set (one B200): other BASELINE configurations at 1e6 clusters, cluster-size sweep of
# the config-3 model, ragged clusters, family sweep; launch list + ncu --set full of the dominant kernel.
mkdir -p gpurun_out
: > gpurun_out/r2_configs_1gpu_1e6.jsonl
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r2_configs_1gpu_1e6.jsonl
for w in C2 C4 C5a C5b C1; do
  python bench.py --workload $w --clusters 1000000 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_configs_1gpu_1e6.jsonl
done
: > gpurun_out/r2_cluster_sizes_C3.jsonl
for ni in 8 12 16 17 24 32 33 48 64 128; do
  python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 >> gpurun_out/r2_cluster_sizes_C3.jsonl
done
python bench.py --ragged --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_ragged_C3.jsonl
python scripts/family_sweep.py 200000 > gpurun_out/r2_family_sweep.txt 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fit > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_dev.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fit > gpurun_out/ncu3.log 2>&1
python bench.py --steps 3 --warmup 3 --clusters 200000 --no-cpu-baseline --no-fit > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_fast -s 3 -c 1 -f -o gpurun_out/prof_dev python bench.py --steps 3 --warmup 3 --clusters 200000 --no-cpu-baseline --no-fit > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/r2_family_sweep.txt
# (the ncu --set full captures of the config-4 / config-5a kernels are taken one per call: scripts/iter_gq.sh -- three
# reports in one call exceed what gpurun copies back)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
# two-pass tile against the generic kernel on the same data
{
echo "# two-pass tile of the register variant vs the generic kernel (GLMMA_NO_TILE32=1); config-3 model, 4e6 observations"
for ni in 17 32 33 64 128; do
  for sw in 0 1; do
    if [ $sw = 1 ]; then export GLMMA_NO_TILE32=1; else unset GLMMA_NO_TILE32; fi
    python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json,os
j=json.loads(sys.stdin.read()); print(('generic ' if os.environ.get('GLMMA_NO_TILE32') else 'two-pass'), j['config']['workload'][:62], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'loglik %.10f'%j['check']['loglik'])"
  done
done
for sw in 0 1; do
  if [ $sw = 1 ]; then export GLMMA_NO_TILE32=1; else unset GLMMA_NO_TILE32; fi
  python bench.py --ragged --ni 40 --clusters 100000 --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json,os
j=json.loads(sys.stdin.read()); print(('generic ' if os.environ.get('GLMMA_NO_TILE32') else 'two-pass'), 'ragged 1 + Poisson(39)', j['config']['workload'][:50], '| %.1f evals/s %.3f ms'%(j['value'], j['ms_per_step']), 'loglik %.10f'%j['check']['loglik'])"
done
unset GLMMA_NO_TILE32
} > gpurun_out/r2_two_pass_tile.txt
cat gpurun_out/r2_two_pass_tile.txt
