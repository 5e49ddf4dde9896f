set -x
timeout 900 python -m pytest tests/test_gpu_mobi.py -m gpu -x -q -k "one_model_year" 2>&1 | tail -15
