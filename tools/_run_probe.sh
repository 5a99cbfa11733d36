cd /root/repo
python bench.py > gpurun_out/r02_bench_n1_v3.json 2> gpurun_out/r02_bench_n1_v3.err; tail -c 300 gpurun_out/r02_bench_n1_v3.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
