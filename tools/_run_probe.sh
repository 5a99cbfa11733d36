cd /root/repo
python bench.py --blocks none --cpu-build "" --builds 2 > gpurun_out/r02_bench_n1_quick.json 2> gpurun_out/r02_bench_n1_quick.err; tail -c 300 gpurun_out/r02_bench_n1_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['sequential_order'])
PY
python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests29.log 2>&1; tail -2 gpurun_out/r02_tests29.log
