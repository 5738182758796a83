for cons in 8 16 24 31; do for stg in 8 16; do
  DRGNN_NVCC_EXTRA="-DDRGNN_CONSUMERS=$cons -DDRGNN_MAX_STAGES=$stg" python -m deeprank_gnn_b200.build --force > /dev/null 2>&1
  echo "== consumers $cons stages $stg"
  python tools/agg_stream.py 6553600 32 10 | grep tiled
  python tools/agg_stream.py 6553600 16 10 | grep tiled
  python tools/agg_stream.py 6553600 64 10 | grep tiled
done; done
