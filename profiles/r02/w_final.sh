#!/bin/bash
# round 2, call W: final state - full GPU suite, smoke(), default bench line, reference arm
set -x
O=gpurun_out/r02w; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; tail -10 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py > $O/bench_npt_256.json 2> $O/bench_npt_256.err; cat $O/bench_npt_256.json | cut -c1-400
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
