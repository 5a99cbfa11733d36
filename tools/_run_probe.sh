cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests21.log 2>&1; tail -3 gpurun_out/r02_tests21.log
for v in base MAIN; do
  if [ $v = MAIN ]; then L=/root/repo/parallel_hnsw_b200/libphnsw.so; else L=/root/repo/parallel_hnsw_b200/build/libphnsw_$v.so; fi
  echo "== $v"
  PHNSW_LIB=$L timeout 600 python tools/probe_adc_q8.py 1000000 10000 2>&1 | tail -4
  PHNSW_LIB=$L timeout 300 python tools/probe_k1.py --order 1 --overlap --tag ${v}_overlap 2>&1 | grep PROBE
done | tee gpurun_out/r02_probe24.log
