#!/bin/bash
# A/B of library builds on one box: bench.py (headline only) per library given on the command line.
for lib in "$@"; do
  for rep in 1 2; do
    HSL_B200_LIB=$lib python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', 'value %.4e kernel_ms %.4f frac %.4f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  done
done
