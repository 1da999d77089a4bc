
for v in "" loop0 loop1 loop2s8 loop2w1 loop2w4; do
  if [ -z "$v" ]; then unset KFB_LIB; else export KFB_LIB=$PWD/build/variants/libkfb200_$v.so; fi
  echo "== variant: ${v:-default}"
  python tools/quick_bench.py 65536 1000 2>&1 | grep "generic_adjoint=False"
  python tools/quick_bench.py 1048576 1000 2>&1 | grep "generic_adjoint=False"
done
