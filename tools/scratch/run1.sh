python -m pytest tests/test_gpu_kzg.py tests/test_gpu_protocols.py -x -q -m gpu 2>&1 | tail -5
python tools/pass_bench.py 0 1 0 1 2>&1 | tail -20
