# quick check of a build on one B200: GPU tests, then the configurations whose occupancy changed
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/quick_pytest.log
: > gpurun_out/quick.jsonl
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 >> gpurun_out/quick.jsonl
for w in C4 C5a C2 C1; do
  python bench.py --workload $w --clusters 1000000 --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 >> gpurun_out/quick.jsonl
done
for ni in 12 16; do
  python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 >> gpurun_out/quick.jsonl
done
cat gpurun_out/quick_pytest.log
python - <<'PY'
import json
for l in open("gpurun_out/quick.jsonl"):
    if l.startswith("{"):
        j = json.loads(l); print(j["config"]["workload"][:60], "| %.1f evals/s  %.3f ms" % (j["value"], j["ms_per_step"]))
PY
python scripts/family_sweep.py 200000 2>&1 | tail -18
