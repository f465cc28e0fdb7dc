#!/bin/bash
# A/B of library builds with the same ABI: tools/ab_bench.sh libA.so libB.so ...  (prints pyramid / base-level / direct search times)
for L in "$@"; do
  VVENC_B200_LIB=$PWD/$L python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); e=d['extra']
print('$L', 'step', round(d['ms_per_step'],3), 'pyr', round(e['pyramid_ms'],3), 'base', round(e['base_level_direct_ms'],3), 'direct', round(e['direct_search']['search_ms'],3), {k:round(v['sad_search_ms'],3) for k,v in e['kernel_ms'].items()})"
done
