#!/bin/bash
for n in 128 256; do
python bench.py --no-cpu --no-extra --steps 5 --warmup 3 --e2e-pairs $n 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print($n,'e2e',d['e2e']['value'],d['e2e']['link']['frac_of_ceiling'],'i32',d['e2e']['with_i32_web']['value'],d['e2e']['with_i32_web']['frac_of_ceiling'],'ceil',d['e2e']['link']['ceiling_MDE_per_s'])"
done
