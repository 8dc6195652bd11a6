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
print('$1', 'ms/step %.3f e2e %.3g (%.3f ms) coal %.3f vterm %.3f'%(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], g('k_coal_small'), g('k_vterm')))
"
}
LCX_COAL_FUSE_VT=0 run "separate vterm"
run "fused vterm"
