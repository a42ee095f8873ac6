timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 1200 python tools/run_configs.py > gpurun_out/run_configs.log 2>&1
timeout 200 python tools/probe_gpu.py > gpurun_out/probe.log 2>&1
timeout 200 python tools/vector_probe.py 25 > gpurun_out/vector_probe.txt 2>&1
timeout 100 python examples/quickstart.py > gpurun_out/quickstart.log 2>&1; echo quickstart rc=$?
tail -7 gpurun_out/run_configs.log | cut -c1-160
