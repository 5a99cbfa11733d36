cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests27.log 2>&1; tail -2 gpurun_out/r02_tests27.log
for cw in auto 3 2 auto; do
  if [ $cw = auto ]; then unset PHNSW_CTA_WARPS; else export PHNSW_CTA_WARPS=$cw; fi
  timeout 300 python tools/probe_k1.py --order 1 --overlap --tag "cta${cw}_overlap" 2>&1 | grep PROBE
done | tee gpurun_out/r02_probe37.log
unset PHNSW_CTA_WARPS
timeout 300 python tools/probe_k1.py --order 1 0 --tag "plain" 2>&1 | grep PROBE | tee -a gpurun_out/r02_probe37.log
