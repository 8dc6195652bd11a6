cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
def g(s):
    v=[v for n,v in ks.items() if s in n]
    return sum(x['ms']/x['launches'] for x in v) if v else 0
print('$1', 'ms/step %.3f e2e %.3g (%.3f ms) transport %.3f'%(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], g('k_transport')))
"
}
for defs in "-DLCX_TRC_MINB=3" "-DLCX_TRC_MINB=4" "-DLCX_TRC_MINB=5" "-DLCX_TRC_MINB=6"; do
  touch libcloudphxx_b200/csrc/lcx_transport.cu
  LCX_DEFS_LCX_TRANSPORT="$defs" python -c "
import sys; sys.path.insert(0,'.')
from libcloudphxx_b200 import build
build.build_all(verbose=False)"
  run "$defs"
done
