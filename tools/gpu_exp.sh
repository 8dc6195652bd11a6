cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-steps 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step %.3f e2e %.3g (%.3f ms)'%(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
"
