# MM_PIN experiment: constants of the SINGLE path pinned into vector registers (main lib = MM_PIN 3)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_md_gpu.py tests/test_force_gpu.py -q -m gpu -x 2>&1 | tail -3
B="python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline"
for e in nve npt; do
  $B --ensemble $e > gpurun_out/pin3_$e.json 2>gpurun_out/pin3_$e.err
  for p in 0 4 6; do MICMEC_B200_LIB=$PWD/profiles/ablate/lib_pin$p.so $B --ensemble $e > gpurun_out/pin${p}_$e.json 2>gpurun_out/pin${p}_$e.err; done
done
for f in gpurun_out/pin*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print('$f', '%.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'STEP %.4f'%r['kernel_ms'], 'FORCE %.4f'%(r.get('force_only_kernel') or {}).get('kernel_ms',0), 'frac %.3f'%r['frac'], 'epot', d['check']['epot'])
" || tail -3 ${f%.json}.err; done
