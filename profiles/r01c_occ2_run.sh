mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline"
$B > gpurun_out/occ_base.json 2>gpurun_out/occ_base.err
MICMEC_B200_LIB=$PWD/profiles/ablate/lib_occ2.so $B > gpurun_out/occ2_v14.json 2>gpurun_out/occ2_v14.err
MICMEC_B200_LIB=$PWD/profiles/ablate/lib_occ2.so $B --variant 10 > gpurun_out/occ2_v10.json 2>gpurun_out/occ2_v10.err
MICMEC_B200_LIB=$PWD/profiles/ablate/lib_occ2.so $B --variant 2 > gpurun_out/occ2_v2.json 2>gpurun_out/occ2_v2.err
for f in gpurun_out/occ*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print('$f', '%.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'STEP %.4f'%r['kernel_ms'], 'FORCE %.4f'%r['force_only_kernel']['kernel_ms'], 'cons', d['check']['epot'])
" || tail -3 ${f%.json}.err; done
