# development loop: `make -C glmmadaptive_b200/csrc DEV=1` builds libglmma_dev.so (negative binomial only)
export GLMMA_LIB=${GLMMA_LIB:-$PWD/glmmadaptive_b200/libglmma_dev.so}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fit 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('evals/s', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'ms_step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'modes cold/warm Mcl/s', round(d['find_modes']['cold_clusters_per_s']/1e6,1), round(d['find_modes']['warm_clusters_per_s']/1e6,1))
    else: print(ln.rstrip()[:300])
"
