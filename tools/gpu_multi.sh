cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -40
