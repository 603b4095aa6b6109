# development loop on the GPU box: negative-binomial-only build (make -C glmmadaptive_b200/csrc DEV=1),
# the parity tests that family covers, then the short bench -- fused path and, for comparison, the plain one
export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fit.py -x -q -k "C3 or negbin or negative or weights_offset or large_cluster or ragged" 2>&1 | tail -5
bash scripts/dev_bench.sh
GLMMA_NO_FUSED=1 bash scripts/dev_bench.sh
