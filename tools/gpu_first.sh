cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fixtures.py -q -m gpu 2>&1 | tail -30
python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -5
