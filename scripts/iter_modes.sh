export GLMMA_LIB=$PWD/glmmadaptive_b200/libglmma_dev.so
for g in 1 8 32; do
export GLMMA_MODES_G=$g
echo "G=$g"
for ni in 4 8 12 24; do
python bench.py --ni $ni --clusters $((4000000 / ni)) --steps 3 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); f=j['find_modes']; print('  ni', j['config']['workload'].split(' x ')[1][:6], 'modes cold %.1f warm %.1f Mcl/s'%(f['cold_clusters_per_s']/1e6, f['warm_clusters_per_s']/1e6), 'iters %.2f'%f['mean_iterations_cold'])"
done
done
