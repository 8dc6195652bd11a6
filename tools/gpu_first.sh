cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -8
python -m pytest tests/test_gpu_fixtures.py -q -m gpu -k "golden or bott or golovin or sstp2" 2>&1 | tail -4
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g ms/step %.2f e2e %.4g frac %.3f launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac_of_hbm_roofline'],d['gpu_launches']))
for k,v in list(d['kernels'].items())[:14]: print('  %-40s %3d %8.3f ms %.3f'%(k,v['launches'],v['ms'],v['share']))
"
