#!/bin/bash
# round 2, call L: evidence for the product build - GPU tests, sanitizer on the new kernels, ncu launch list + full capture of
# the NPT step, bench lines (NPT / NVE / NVT 256^3, NVT 64^3, default model, reference arm)
set -x
O=gpurun_out/r02l; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python profiles/r02/sanitize_run.py > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
timeout 900 compute-sanitizer --tool memcheck python profiles/r02/sanitize_run.py > $O/memcheck.log 2>&1; tail -3 $O/memcheck.log
MICMEC_B200_WRAP_ON_LOAD=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python profiles/r02/sanitize_run.py > $O/racecheck_images.log 2>&1; tail -3 $O/racecheck_images.log
MICMEC_B200_WRAP_ON_LOAD=1 MICMEC_B200_TAIL_IN_KERNEL=1 timeout 900 compute-sanitizer --tool memcheck python profiles/r02/sanitize_run.py > $O/memcheck_images_tik.log 2>&1; tail -3 $O/memcheck_images_tik.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $O/launches_npt.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > $O/launches_npt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_march2 -s 10 -c 3 -o /tmp/npt_final \
    python bench.py --ensemble npt --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_npt.log 2>&1; tail -2 $O/ncu_npt.log | cut -c1-200
ncu -i /tmp/npt_final.ncu-rep --page raw --csv > $O/npt_final_raw.csv 2>/dev/null
python bench.py > $O/bench_npt_256.json 2> $O/bench_npt_256.err; cat $O/bench_npt_256.json
python bench.py --ensemble nve --no-cpu-baseline > $O/bench_nve_256.json 2> $O/bench_nve_256.err
python bench.py --ensemble nvt --no-cpu-baseline > $O/bench_nvt_256.json 2> $O/bench_nvt_256.err
python bench.py --ensemble nvt --grid 64 --no-cpu-baseline > $O/bench_nvt_64.json 2> $O/bench_nvt_64.err
python bench.py --model default --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_npt_256_default_model.json 2> $O/bench_npt_256_default_model.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
