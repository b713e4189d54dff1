for v in 0 1 2; do
  echo "K1 variant $v"; AES_K1_VARIANT=$v CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
done
AES_K1_VARIANT=2 timeout 300 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k spectral 2>&1 | tail -1
