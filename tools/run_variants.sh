for v in "" "$@"; do
  if [ -z "$v" ]; then unset KFB_LIB; else export KFB_LIB=$PWD/build/variants/libkfb200_$v.so; fi
  echo "== variant: ${v:-default}"
  python tools/quick_bench.py 65536 1000 2>&1 | grep "generic_adjoint=False" | cut -c1-150
  python tools/quick_bench.py 1048576 1000 2>&1 | grep "generic_adjoint=False" | cut -c1-150
done
