mkdir -p gpurun_out
( time timeout 150 python -m pytest tests -q -m gpu -x 2>&1 ) > gpurun_out/pytest_gpu_final.log 2>&1; tail -6 gpurun_out/pytest_gpu_final.log
timeout 70 python profiles/replica_opt.py --nrep 1280 --cpu-sample 6 > gpurun_out/replica_opt_1280b.json 2>gpurun_out/replica_opt_1280b.err; tail -c 1200 gpurun_out/replica_opt_1280b.json; tail -3 gpurun_out/replica_opt_1280b.err
