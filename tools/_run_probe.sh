cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests25.log 2>&1; tail -2 gpurun_out/r02_tests25.log
for v in base MAIN; do
  if [ $v = MAIN ]; then L=/root/repo/parallel_hnsw_b200/libphnsw.so; else L=/root/repo/parallel_hnsw_b200/build/libphnsw_$v.so; fi
  PHNSW_LIB=$L timeout 300 python tools/probe_cos.py 400000 10000 $v 2>&1 | grep PROBE
  PHNSW_LIB=$L timeout 300 python tools/probe_k1.py --order 1 0 --tag $v 2>&1 | grep PROBE
  PHNSW_LIB=$L timeout 300 python tools/probe_build.py 5 2>&1 | grep BUILD | tr '\n' ' '; echo
done | tee gpurun_out/r02_probe35.log
