for L in "" c32 c16; do
  if [ -n "$L" ]; then export TSGPU_LIB=$PWD/multilinear-map-cryptography_b200/libtsgpu_$L.so; fi
  echo "== chunk variant: ${L:-64}"
  python tools/shape_n8.py 17 10 2>&1 | tail -1
  python tools/shape_n8.py 18 10 2>&1 | tail -1
  python tools/shape_n8.py 20 5 2>&1 | tail -1
done
