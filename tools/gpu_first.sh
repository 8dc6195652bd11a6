cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g ms/step %.2f e2e %.4g frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac_of_hbm_roofline']))
for k,v in list(d['kernels'].items())[:8]: print('  %-40s %3d %8.3f ms %.3f'%(k,v['launches'],v['ms'],v['share']))
"; }
echo "=== fmad=false"
python -m pytest tests/test_gpu_parity.py -q -m gpu -s 2>&1 | grep -E "^E  |parcel sstp|full step|passed|failed" | head -40
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_nofma.log 2>&1; show gpurun_out/bench_nofma.log
echo "=== fmad=true for lcx_cond.cu"
touch libcloudphxx_b200/csrc/lcx_cond.cu
LCX_COND_FMAD=1 python -c "
import sys; sys.path.insert(0,'.')
from libcloudphxx_b200 import build
build.build_all(verbose=False)"
python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "parcel or full_step" 2>&1 | grep -E "^E  |parcel sstp|full step|passed|failed" | head -40
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fma.log 2>&1; show gpurun_out/bench_fma.log
