# Round-end evidence (one B200): the GPU suite, smoke, fast-math accuracy, the default bench line, cfg5 and f32 lines, the ncu
# launch list of the bench command and ncu --set full of the hot kernels at the bench size.  Usage: gpu_profile_final.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_suite.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 tools/check_fastmath.bin
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -c 400 gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err
timeout 300 python bench.py --config cfg5 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err
timeout 300 python bench.py --real f32 --no-cpu-baseline --no-alt > gpurun_out/${TAG}_bench_f32.json 2> gpurun_out/${TAG}_bench_f32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-alt --profile-steps 0 > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_launches.log | cut -c1-200
# resident steps only (--profile-steps 0, e2e loop follows): skip the launches of init + the first API step, take one of each hot kernel in steady state
LCX_SYNC_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"k_cond_range|k_coal_small|k_transport|k_gather|k_vterm|k_radix_scatter|k_mv_place|k_mv_list|k_mv_count" -s 12 -c 12 -o gpurun_out/${TAG}_hot \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --profile-steps 0 > gpurun_out/${TAG}_ncu_hot.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_hot.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}
