#!/bin/bash
# usage: tools/envs.sh <what> "<env settings A>" "<env settings B>" ...   (same library, different env hooks)
what=$1; shift
fmt='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["config"],d["w"],d["D"],d["sw"],d["variant"],"main",d["main_kernel_us"],"pack",d["pack_kernel_us"],"batch_us",d.get("batch_us_per_pair"),d["equal_direct_kernel"],d["equal_oracle_slab"],d.get("batch_equal"))
    elif "rror" in l: print(l.strip())'
for rep in 1 2; do
  for e in "$@"; do
    echo "== [$e]"; env $e python tests/sweep_configs.py --what $what 2>&1 | python -c "$fmt"
  done
done
