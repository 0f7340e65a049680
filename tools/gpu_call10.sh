#!/bin/bash
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/c10_pytest.log 2>&1; tail -4 $O/c10_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/c10_smoke.log 2>&1; cat $O/c10_smoke.log
python tools/gpu_dbg.py
(time python bench.py) > $O/c10_bench.json 2> $O/c10_bench.err; tail -n 5 $O/c10_bench.err
python tools/latency.py > $O/c10_latency.log 2>&1; cat $O/c10_latency.log
python tests/sweep_configs.py --what c1,c2,c3,c4 --md $O/c10_sweep.md > $O/c10_sweep.jsonl 2> $O/c10_sweep.err; cat $O/c10_sweep.md
