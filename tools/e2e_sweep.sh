#!/bin/bash
# e2e (host buffers, PCIe inside the timed region) for several chunk / stream settings of roundtrip_host
for cfg in "16 4" "32 4" "8 4" "32 8" "64 8"; do
  set -- $cfg
  SNFFT_E2E_CHUNKS=$1 SNFFT_E2E_STREAMS=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $1 streams $2 e2e', d['e2e']['value']/1e9, 'G el/s')"
done
