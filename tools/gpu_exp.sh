cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  python bench.py --nx 32 --ny 128 --nz 128 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
def g(s):
    v=[v for n,v in ks.items() if s in n]
    return sum(x['ms'] for x in v)/d['steps'] if v else 0
print('$1', 'ms/step %.3f  cond %.3f vterm %.3f transport %.3f coal %.3f'%(d['ms_per_step'], g('k_cond_cells'), g('k_vterm'), g('k_transport'), g('k_coal_small')))
"
}
for defs in "-DLCX_COND_MINB=8" "-DLCX_COND_MINB=8 -DLCX_NO_FAST_DIV" "-DLCX_COND_MINB=6" "-DLCX_COND_MINB=5" ; do
  touch libcloudphxx_b200/csrc/lcx_cond.cu
  LCX_COND_DEFS="$defs" python -c "
import sys; sys.path.insert(0,'.')
from libcloudphxx_b200 import build
build.build_all(verbose=False)"
  run "$defs fused"
  LCX_NO_REFRESH_FUSE=1 run "$defs unfused"
done
