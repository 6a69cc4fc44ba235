#!/bin/bash
# kernel time of the bench configs as a function of the L2 prefetch distance (CTAs ahead; 0 = off)
for cfg in c2 c3 c5; do
  for pf in 0 148 296 444 592 888; do
    NTHASH_B200_PREFETCH_CTAS=$pf python bench.py --config $cfg --steps 10 --no-cpu-baseline --e2e-steps 1 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg prefetch_ctas=$pf kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
  done
  python bench.py --config $cfg --steps 10 --no-cpu-baseline --e2e-steps 1 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg default kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
