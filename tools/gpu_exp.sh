cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks=d['kernels']
print('$1', 'ms/step %.3f e2e %.3g (%.3f ms)'%(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
for k,v in list(ks.items())[:3]: print('   %-40s %3d %8.3f'%(k[:40], v['launches'], v['ms']/2))
"
}
run "secant"
