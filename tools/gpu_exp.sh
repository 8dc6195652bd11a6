cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
def g(s):
    v=[v for n,v in ks.items() if s in n]
    return sum(x['ms']/x['launches'] for x in v) if v else 0
print('$1', 'ms/step %.3f  cond %.3f vterm %.3f transport %.3f coal %.3f gather %.3f'%(d['ms_per_step'], g('k_cond_cells'), g('k_vterm'), g('k_transport'), g('k_coal_small'), g('k_gather')))
"
}
for defs in "-DLCX_TR_MINB=5" "-DLCX_TR_MINB=6" ; do
  touch libcloudphxx_b200/csrc/lcx_transport.cu
  LCX_DEFS_LCX_TRANSPORT="$defs" python -c "
import sys; sys.path.insert(0,'.')
from libcloudphxx_b200 import build
build.build_all(verbose=False)"
  for ctas in 32 48 64; do
    LCX_TR_CTAS=$ctas run "$defs ctas=$ctas"
  done
done
