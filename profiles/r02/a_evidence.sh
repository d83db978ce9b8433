#!/bin/bash
# round 2, call A: state of HEAD before the kernel work - GPU tests, sanitizer, ncu --set full of the three marching
# kernels of one NPT step (FORCE rotate+write, STEP, FORCE rotate)
set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python profiles/r02/sanitize_run.py > $O/racecheck.log 2>&1; tail -5 $O/racecheck.log
timeout 900 compute-sanitizer --tool memcheck python profiles/r02/sanitize_run.py > $O/memcheck.log 2>&1; tail -5 $O/memcheck.log
timeout 600 compute-sanitizer --tool synccheck python profiles/r02/sanitize_run.py > $O/synccheck.log 2>&1; tail -3 $O/synccheck.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_march -s 10 -c 3 -o $O/march_npt \
    python bench.py --ensemble npt --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_npt.log 2>&1; tail -3 $O/ncu_npt.log
python bench.py --ensemble npt --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_npt.json 2> $O/bench_npt.err; cat $O/bench_npt.json
ls -la $O
