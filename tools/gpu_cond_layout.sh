# Parity of the condensation kernel under every work distribution, its timing per layout at bench size (equal step numbers),
# and one ncu --set full capture of the range kernel.  Variant builds (LCX_BUILD_TAG=<tag> python -m libcloudphxx_b200.build, then
# LCX_B200_LIBDIR=.../lib_<tag>) can be timed with the same tools/exp_cond_layout.py call.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cond_layout or layouts_agree" 2>&1 | tail -5
timeout 200 python tools/exp_cond_layout.py 64 256 128 -1 4 8 16 > gpurun_out/exp_layout.log 2>&1; tail -8 gpurun_out/exp_layout.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cond_range" -s 2 -c 1 -o gpurun_out/cond_range \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_cond_range.log 2>&1
tail -3 gpurun_out/ncu_cond_range.log | cut -c1-200
