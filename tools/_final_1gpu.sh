set -x
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> /dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 2 -o gpurun_out/r1_gram_v6 python tools/ncu_target.py gram > gpurun_out/ncu_gram6.log 2>&1
timeout 1200 python tools/run_configs.py > gpurun_out/run_configs.log 2>&1
timeout 200 python tools/probe_gpu.py > gpurun_out/probe.log 2>&1
timeout 300 python tools/bootstrap_probe.py 10000000 100 50 > gpurun_out/bootstrap_probe.log 2>&1
timeout 200 python tools/fp64_probe.py > gpurun_out/fp64_probe.txt 2>&1
timeout 200 python tools/vector_probe.py 25 > gpurun_out/vector_probe.txt 2>&1
tail -6 gpurun_out/run_configs.log | cut -c1-200
