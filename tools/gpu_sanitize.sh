# compute-sanitizer over the code paths added in round 2 (small cases; the tools slow kernels down 10-50x)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 9 --target-processes all"
( timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_sync_chunks.py tests/test_gpu_device_init.py -q -x -k "not f32" 2>&1 | tail -4 ) > gpurun_out/r02_san_memcheck_a.log 2>&1
( timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_multi.py -q -x -k "pred_corr or roundtrip_on_one_device or torchrun_ranks_roundtrip" 2>&1 | tail -4 ) > gpurun_out/r02_san_memcheck_b.log 2>&1
( timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_lazy_gather.py tests/test_gpu_f32.py tests/test_gpu_philox.py -q -x -k "lazy_equals_eager or golovin_box_exact or full_step or resident" 2>&1 | tail -4 ) > gpurun_out/r02_san_memcheck_c.log 2>&1
( timeout 600 $S --tool racecheck python -m pytest tests/test_gpu_sync_chunks.py tests/test_gpu_lazy_gather.py -q -x -k "chunked_equals_whole or lazy_equals_eager" 2>&1 | tail -4 ) > gpurun_out/r02_san_racecheck.log 2>&1
tail -n 3 gpurun_out/r02_san_*.log
