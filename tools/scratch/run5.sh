for v in "" _c128 _c256; do
  echo "== lib$v"
  TSGPU_LIB=$PWD/multilinear-map-cryptography_b200/libtsgpu$v.so python tools/pass_bench.py 1 2>&1 | grep pass
  TSGPU_LIB=$PWD/multilinear-map-cryptography_b200/libtsgpu$v.so python tools/shape_n8.py 17 5 2>&1 | tail -1
done
