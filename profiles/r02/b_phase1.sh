#!/bin/bash
# round 2, call B: phase-1 checks - new GPU tests (256^3 / 64^3 parity, restart), smoke(), bench line with the new check block
set -x
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py --steps 100 --warmup 10 > $O/bench_npt.json 2> $O/bench_npt.err; cat $O/bench_npt.json; tail -2 $O/bench_npt.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
nproc; free -g | head -2
