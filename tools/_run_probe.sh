cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests26.log 2>&1; tail -2 gpurun_out/r02_tests26.log
timeout 300 python tools/probe_k1.py --order 1 0 --overlap --tag "auto_overlap" 2>&1 | grep PROBE
PHNSW_CTA_WARPS=4 timeout 300 python tools/probe_k1.py --order 0 --overlap --tag "cta4_overlap" 2>&1 | grep PROBE
PHNSW_CTA_WARPS=5 timeout 300 python tools/probe_k1.py --order 0 --overlap --tag "cta5_overlap" 2>&1 | grep PROBE
