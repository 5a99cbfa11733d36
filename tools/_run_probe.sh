cd /root/repo
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 > gpurun_out/r02_bench_n2_v3.json 2> gpurun_out/r02_bench_n2_v3.err
tail -3 gpurun_out/r02_bench_n2_v3.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_v3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['recall_at_10'], d.get('e2e'))
c4=d.get('config4'); 
print(json.dumps({k:v for k,v in (c4 or {}).items() if k not in ('workload','mode','roofline','parity')},indent=1)[:2500])
PY
