for i in 1 2 3; do
  ZEDO_B200_LIB=$PWD/zedo_release_b200/libzedo_b200_base.so python tools/layer_bench.py 262144 30 fp8lo 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', d['fp8lo/exp0'])"
  python tools/layer_bench.py 262144 30 fp8lo 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new ', d['fp8lo/exp0'])"
done
