cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
def g(s):
    v=[v for n,v in ks.items() if s in n]
    return sum(x['ms']/x['launches'] for x in v) if v else 0
print('$1', 'ms/step %.3f e2e %.3g cond %.3f vterm %.3f transport %.3f coal %.3f gather %.3f'%(d['ms_per_step'], d['e2e']['value'], g('k_cond_cells'), g('k_vterm'), g('k_transport'), g('k_coal_small'), g('k_gather')))
"
}
for defs in "-DLCX_NU_UNROLL" "-DLCX_EXP_ESTRIN" "-DLCX_NU_UNROLL -DLCX_EXP_ESTRIN" "-DLCX_NU_UNROLL -DLCX_EXP_ESTRIN -DLCX_COND_MINB=6"; do
  touch libcloudphxx_b200/csrc/lcx_cond.cu
  LCX_COND_DEFS="$defs" python -c "
import sys; sys.path.insert(0,'.')
from libcloudphxx_b200 import build
build.build_all(verbose=False)"
  run "$defs"
done
