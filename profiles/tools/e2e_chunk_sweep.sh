#!/bin/bash
# End-to-end rate (host buffers in/out) of the default bench workload against the number of
# sub-batches the host pipeline cuts a batch into (AES_HOST_CHUNKS; aes_chain_process_host).
for c in "$@"; do
  AES_HOST_CHUNKS=$c python bench.py --no-cpu --steps 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); e = d['e2e']
print('chunks', $c, 'e2e', round(e['value']), 'Ms/s  matches', e.get('matches_device_path'), ' pcm16', round(e['pcm16_file_route']['value']), 'Ms/s')"
done
