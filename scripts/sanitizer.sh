# compute-sanitizer passes over the small-shape GPU tests (one family per kernel variant: register
# variant NJ = 8 / 12 / 16, generic staged / tile-outer / streaming, EM modes 1 / 2, mode search G = 1 / 8 / 32,
# log_post_b, prep on lists): memcheck, racecheck (shared-memory hazards), synccheck.  Summaries -> gpurun_out/.
set -u
mkdir -p gpurun_out
SEL='test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fit.py'
K='core_families or complete_clusters or weights_offset or large_cluster or modes_vs_oracle or em_estep or larger_than_the_staged or ragged or single_observation or log_post_b or nagq'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/$SEL -x -q -k "$K" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -4
done
