mkdir -p gpurun_out
timeout 120 python profiles/replica_batch.py > gpurun_out/replica_batch.log 2>&1; tail -2 gpurun_out/replica_batch.log
timeout 150 python profiles/replica_opt.py --nrep 1280 --cpu-sample 6 > gpurun_out/replica_opt_1280.json 2>gpurun_out/replica_opt_1280.err; tail -c 1500 gpurun_out/replica_opt_1280.json; tail -3 gpurun_out/replica_opt_1280.err
timeout 400 python profiles/replica_opt.py --nrep 10240 --cpu-sample 0 > gpurun_out/replica_opt_10240.json 2>gpurun_out/replica_opt_10240.err; tail -c 1500 gpurun_out/replica_opt_10240.json; tail -3 gpurun_out/replica_opt_10240.err
