cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "parcel" 2>&1 | grep -E "^E  |parcel sstp|passed|failed" | head -60
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not parcel" 2>&1 | tail -5
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1; tail -3 gpurun_out/bench_full.log
