cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --blocks none --cpu-build "" --builds 2 > gpurun_out/r02_bench_n2_v2.json 2> gpurun_out/r02_bench_n2_v2.err
tail -3 gpurun_out/r02_bench_n2_v2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_v2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['recall_at_10'])
print(json.dumps(d['sharded_step'],indent=1)[:2500])
PY
