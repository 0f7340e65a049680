#!/bin/bash
# usage: tools/scale.sh N   -> gpurun_out/sc_bench_nN.json, sc_bands_nN.json
N=$1; O=gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --no-cpu > $O/sc_bench_n1.json 2> $O/sc_bench_n1.err
  python bench_bands.py > $O/sc_bands_n1.json 2> $O/sc_bands_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/sc_bench_n$N.json 2> $O/sc_bench_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench_bands.py --gpus $N > $O/sc_bands_n$N.json 2> $O/sc_bands_n$N.err
fi
python - <<PY
import json
for f in ("$O/sc_bench_n$N.json", "$O/sc_bands_n$N.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["e2e"].get("with_u8_web",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
