# r2ak: rows10 with the next item's rows fetched by bulk copies (variants 8, 9, 10) against the shipped <1,4,1>
timeout 600 python -m pytest tests -m gpu -x -q -k "spectral or Clean or noise or golden or preset" 2>&1 | tail -3
for v in 0 8 9 10; do
  echo "variant $v"; AES_ROWS10_VARIANT=$v CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
  AES_ROWS10_VARIANT=$v timeout 600 python -m pytest tests -m gpu -x -q -k "spectral or Clean or noise" 2>&1 | tail -1
done
