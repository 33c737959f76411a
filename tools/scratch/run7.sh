python tools/shape_n8.py 17 10 2>&1 | tail -1
python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
