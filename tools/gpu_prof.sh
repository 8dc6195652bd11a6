cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r01c}
ncu --set full --clock-control none --import-source on -k regex:"k_cond_cells|k_coal_small|k_transport|k_gather|k_vterm|k_radix_scatter" -s 8 -c 8 -o gpurun_out/hot_$TAG \
    python bench.py --nx 32 --ny 128 --nz 128 --steps 1 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_hot.log 2>&1
tail -2 gpurun_out/ncu_hot.log | cut -c1-200
ls -la gpurun_out
