cd /root/repo
for v in base MAIN rp nls; do
  if [ $v = MAIN ]; then L=/root/repo/parallel_hnsw_b200/libphnsw.so; else L=/root/repo/parallel_hnsw_b200/build/libphnsw_$v.so; fi
  PHNSW_LIB=$L timeout 300 python tools/probe_cos.py 400000 10000 $v 2>&1 | grep PROBE
  PHNSW_LIB=$L timeout 300 python tools/probe_k1.py --order 1 --dim 1536 --n 300000 --tag $v 2>&1 | grep PROBE
done | tee gpurun_out/r02_probe29.log
