cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cond_classed.py tests/test_gpu_cond_staged.py tests/test_gpu_sync_chunks.py tests/test_gpu_lazy_gather.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-alt --profile-steps 0 > /dev/null 2>&1
for v in "cfg4 -1" "cfg4 0" "cfg5 -1" "cfg5 0" "cfg5 1"; do
  set -- $v
  LCX_COND_CLASSED=$2 python bench.py --config $1 --steps 12 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/r02_cls_$1_$2.json 2> gpurun_out/r02_cls_$1_$2.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_cls_$1_$2.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$1 classed $2: ms/step %.3f  e2e %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']), {n: round(v['ms']/v['launches'],3) for n,v in list(k.items())[:4]})
P
done
