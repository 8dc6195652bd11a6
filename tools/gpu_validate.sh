# Round-end validation: the whole GPU parity suite, the smoke check, the fast-math accuracy check and the default bench line.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 tools/check_fastmath.bin
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
