# chunked step_sync: tests, then the bench with chunks off / graded off / default (e2e is what moves)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sync_chunks.py -x -q > gpurun_out/r02_chunks_tests.log 2>&1; tail -3 gpurun_out/r02_chunks_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-alt --profile-steps 0 > /dev/null 2>&1      # the first process on a fresh box runs slower
for v in "1 1" "8 0" "8 1" "12 1" "8 0" "8 1"; do
  set -- $v
  LCX_SYNC_CHUNKS=$1 LCX_SYNC_GRADED=$2 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-alt --profile-steps 0 > gpurun_out/r02_chunks_$1_$2.json 2> gpurun_out/r02_chunks_$1_$2.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_chunks_$1_$2.json').read().strip().splitlines()[-1])
print('chunks $1 graded $2: ms/step %.3f  e2e %.4g (%.3f ms)' % (d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
P
done
