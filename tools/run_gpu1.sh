set -x
python -m pytest tests/test_gpu_rebuild.py -x -q -m gpu 2>&1 | tail -30
