python -m pytest tests/test_gpu_kzg.py tests/test_gpu_protocols.py -x -q -m gpu 2>&1 | tail -3
for cfg in "32 32768" "64 32768" "64 16384" "16 32768" "32 65536"; do
  set -- $cfg
  echo "== RED_SPAN=$1 MIN_SPANS=$2"
  TSGPU_RED_SPAN=$1 TSGPU_RED_MIN_SPANS=$2 python tools/pass_bench.py 1 2>&1 | grep pass
  TSGPU_RED_SPAN=$1 TSGPU_RED_MIN_SPANS=$2 python tools/shape_n8.py 17 5 2>&1 | tail -1
done
