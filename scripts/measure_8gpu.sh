# 8 GPUs of one box (one process per GPU): config 3 scaling point, config 5 at 5e6 clusters and config 4, with whole fits
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
: > gpurun_out/r2_configs_8gpu.jsonl
$TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_configs_8gpu.jsonl
$TR bench.py --gpus 8 --workload C5a --clusters 5000000 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_configs_8gpu.jsonl
$TR bench.py --gpus 8 --workload C5b --clusters 5000000 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_configs_8gpu.jsonl
$TR bench.py --gpus 8 --workload C4 --clusters 1000000 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_configs_8gpu.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2_configs_8gpu.jsonl"):
    if l.startswith("{"):
        j = json.loads(l); mm = j.get("mixed_model") or {}
        print(j["config"]["workload"][:70], "| %.1f evals/s e2e %.1f fit %s s" % (j["value"], j["e2e"]["value"], mm.get("seconds")), j.get("multi_handle_check", {}) and j["multi_handle_check"].get("ok"))
PY
