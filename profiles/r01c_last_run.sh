mkdir -p gpurun_out
timeout 45 python profiles/replica_opt.py --nrep 1280 --cpu-sample 0 > gpurun_out/replica_opt_1280c.json 2>gpurun_out/replica_opt_1280c.err; tail -c 700 gpurun_out/replica_opt_1280c.json; tail -2 gpurun_out/replica_opt_1280c.err
timeout 70 python bench.py > gpurun_out/bench_npt_final.json 2>gpurun_out/bench_npt_final.err; tail -c 2500 gpurun_out/bench_npt_final.json; tail -2 gpurun_out/bench_npt_final.err
