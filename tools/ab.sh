#!/bin/bash
# A/B on the same box: tools/_base.so vs the in-tree library.  usage: tools/ab.sh <what>   (e.g. c2,c4)
L=stereomatching_b200/libstereo_b200.so
cp $L /tmp/new.so
fmt='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["config"],d["D"],d["sw"],d["variant"],"main",d["main_kernel_us"],"pack",d["pack_kernel_us"],"batch_us",d.get("batch_us_per_pair"),d["equal_direct_kernel"],d["equal_oracle_slab"],d.get("batch_equal"))'
for rep in 1 2; do
  cp tools/_base.so $L; echo "== base"; python tests/sweep_configs.py --what $1 2>&1 | python -c "$fmt"
  cp /tmp/new.so $L; echo "== new";  python tests/sweep_configs.py --what $1 2>&1 | python -c "$fmt"
done
