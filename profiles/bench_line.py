import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): print(line[:300]); continue
    b=json.loads(line)
    print("%s  value=%.3e  ms/step=%.3f  kernel_ms=%.3f launches=%s temp=%.2f epot=%.6f" % (b['config']['workload'][:40], b['value'], b['ms_per_step'], b['roofline']['kernel_ms'], b['gpu_launches'], b['check']['temp_K'], b['check']['epot']))
