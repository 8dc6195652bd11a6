# Round-end validation: the whole GPU parity suite, the smoke check, the fast-math accuracy check and the default bench line.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r02i}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_suite.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 tools/check_fastmath.bin
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -c 300 gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
