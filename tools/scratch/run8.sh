python -m pytest tests/test_gpu_kzg.py tests/test_gpu_protocols.py tests/test_gpu_lagrange.py -x -q -m gpu 2>&1 | tail -3
python tools/pass_bench.py 1 2>&1 | grep pass
python tools/shape_n8.py 17 10 2>&1 | tail -1
