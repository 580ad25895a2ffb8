timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_search or newton_configs or survivor" 2>&1 | tail -4
echo "FUSE_BLOCK=8 (default build)"; timeout 300 python tools/time_small_batch.py 1 4096 2>/dev/null | tee gpurun_out/small_batch_fb8.json
for v in 1 16 32; do echo "FUSE_BLOCK=$v"; ACOC_LIB=$PWD/variants/libacoc_fb$v.so timeout 300 python tools/time_small_batch.py 1 4096 2>/dev/null | tee gpurun_out/small_batch_fb$v.json; done
