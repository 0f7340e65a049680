#!/bin/bash
# N GPUs: the driver's own SCALE command (bench.py under torchrun) + the reference arm
N=${1:-2}
O=gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N) > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -n 6 $O/bench_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/bench_n%s.json'%n) if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'n',d['n_gpus'])
print('e2e',d['e2e']['value'],d['e2e'].get('link'))
print('c4',d.get('config4_pairs'))
print('c3',d.get('config3_bands'))
print('parity',d['parity'],'clocks',d['clocks'])
PY
