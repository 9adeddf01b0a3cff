set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r1.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1_final.log 2>&1
timeout 240 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
timeout 240 python bench.py --workload gen512 --skip-cpu-baseline > gpurun_out/bench512_r1_final.json 2> gpurun_out/bench512_r1_final.err
timeout 200 python tools/bench_conv.py --out gpurun_out/bench_conv_r1_final.jsonl > /dev/null 2>&1
timeout 200 python tools/microbench.py --out gpurun_out/microbench_r1_final.jsonl > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench_r1_final.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --no-graph > gpurun_out/ncu_bench_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pg::|bias_act|upfirdn2d|torgb|conv_igemm|instance_stats" -o gpurun_out/ops_r1_final python tools/prof_ops.py > gpurun_out/ncu_ops_final.log 2>&1
