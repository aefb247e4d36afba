#!/bin/bash
# A/B: default library vs variants
for v in default "$@"; do
  if [ "$v" = default ]; then unset SNFFT_B200_LIB; else export SNFFT_B200_LIB=$PWD/snfft_b200/_lib/variants/$v.so; fi
  for rep in 1 2; do
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['per_kind_ms'], d['config'].get('round_trip_rel_err'))"
  done
done
