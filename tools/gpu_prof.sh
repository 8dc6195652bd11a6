cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of the bench command (cold-cache, serialised: shares matter)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --nx 32 --ny 128 --nz 128 --steps 2 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:k_cond_cells -s 1 -c 1 -o gpurun_out/cond_r01 \
    python bench.py --nx 32 --ny 128 --nz 128 --steps 1 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_cond.log 2>&1
tail -2 gpurun_out/ncu_cond.log | cut -c1-200
ls -la gpurun_out
