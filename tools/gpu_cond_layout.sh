# Parity of the balanced-range condensation kernel, then its timing against the 8-lanes-per-cell kernel at bench size,
# with and without the constant-bank coefficient table (lib_nokc = -DLCX_NO_KC build), then the default bench line.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cond_layout or layouts_agree" 2>&1 | tail -5
timeout 200 python tools/exp_cond_layout.py 64 256 128 -1 4 8 16 > gpurun_out/exp_layout_kc.log 2>&1; tail -8 gpurun_out/exp_layout_kc.log
LCX_B200_LIBDIR=$GRAFT_REPO_ROOT/libcloudphxx_b200/lib_nokc timeout 200 python tools/exp_cond_layout.py 64 256 128 -1 16 > gpurun_out/exp_layout_nokc.log 2>&1; tail -4 gpurun_out/exp_layout_nokc.log
timeout 300 python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -c 3000 gpurun_out/bench_r01e.json
