# chunked step_sync: tests, then the bench with chunks off / 4 / 8 / 16 (e2e is what moves)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sync_chunks.py -x -q > gpurun_out/r02_chunks_tests.log 2>&1; tail -3 gpurun_out/r02_chunks_tests.log
for k in 1 4 8 12; do
  LCX_SYNC_CHUNKS=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-alt --profile-steps 0 > gpurun_out/r02_chunks_$k.json 2> gpurun_out/r02_chunks_$k.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_chunks_$k.json').read().strip().splitlines()[-1])
print('chunks $k: ms/step %.3f  e2e %.4g (%.3f ms)' % (d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
P
done
