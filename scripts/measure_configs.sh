# every BASELINE configuration at 1e6 clusters on one GPU (bench.py lines with whole fits)
mkdir -p gpurun_out
: > gpurun_out/r2_configs_1gpu_1e6.jsonl
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r2_configs_1gpu_1e6.jsonl
for w in C2 C4 C5a C5b C1; do
  python bench.py --workload $w --clusters 1000000 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_configs_1gpu_1e6.jsonl
done
python bench.py --ragged --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_ragged_C3.jsonl
python - <<'PY'
import json
for f in ("gpurun_out/r2_configs_1gpu_1e6.jsonl", "gpurun_out/r2_ragged_C3.jsonl"):
    for l in open(f):
        if l.startswith("{"):
            j = json.loads(l); m = j["mixed_model"]
            print(j["config"]["workload"][:60], "| %.1f evals/s e2e %.1f | fit %.3f s (first %.3f)" % (j["value"], j["e2e"]["value"], m["seconds"], m["first_fit_seconds"]))
PY
