# Round-end evidence: (1) ncu launch list of the bench command, (2) ncu --set full of the hot kernels at the bench size.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"k_cond_range|k_cond_cells|k_coal_small|k_transport|k_gather|k_vterm|k_radix_scatter" -s 8 -c 8 -o gpurun_out/hot_full_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_hot_full.log 2>&1
tail -2 gpurun_out/ncu_hot_full.log | cut -c1-300
ls -la gpurun_out | tail -5
