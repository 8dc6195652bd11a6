# Parity of the balanced-range condensation kernel, its timing against the 8-lanes-per-cell kernel at bench size
# (per-lane shared-memory accumulators = the product; lib_shfl = segmented shuffle reduction, -DLCX_COND_RANGE_SHFL),
# and one ncu --set full capture of the range kernel at bench size.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cond_layout or layouts_agree" 2>&1 | tail -5
timeout 200 python tools/exp_cond_layout.py 64 256 128 -1 4 8 > gpurun_out/exp_layout_acc.log 2>&1; tail -6 gpurun_out/exp_layout_acc.log
LCX_B200_LIBDIR=$GRAFT_REPO_ROOT/libcloudphxx_b200/lib_shfl timeout 200 python tools/exp_cond_layout.py 64 256 128 8 16 > gpurun_out/exp_layout_shfl.log 2>&1; tail -4 gpurun_out/exp_layout_shfl.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cond_range" -s 2 -c 1 -o gpurun_out/cond_range_r01 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_cond_range.log 2>&1
tail -3 gpurun_out/ncu_cond_range.log | cut -c1-200
