# Round-2 measurement pass (run under gpurun; every step has its own hard timeout; outputs land in gpurun_out/ and are copied to profiles/ by hand).
set -x
R=gpurun_out
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $R/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $R/r2_smoke.log
timeout -s KILL 700 python -m pytest tests -m gpu -q > $R/r2_pytest_gpu.log 2>&1
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > $R/r2_bench_gen256.json 2> $R/r2_bench_gen256.err
timeout -s KILL 300 python bench.py --steps 200 --warmup 5 --no-extras --skip-cpu-baseline > $R/r2_bench_gen256_long.json 2> $R/r2_bench_gen256_long.err
timeout -s KILL 300 python bench.py --workload gen512 --skip-cpu-baseline --no-extras > $R/r2_bench_gen512.json 2> $R/r2_bench_gen512.err
timeout -s KILL 300 python tools/bench_conv.py --out $R/r2_conv_microbench.jsonl > /dev/null 2>&1
timeout -s KILL 300 python tools/microbench.py --out $R/r2_op_microbench.jsonl > /dev/null 2>&1
timeout -s KILL 300 python baseline/run_reference.py --mode ops > $R/r2_reference_cuda_ops.json 2> $R/r2_reference_cuda_ops.err
timeout -s KILL 200 python tools/step_breakdown.py > $R/r2_step_breakdown_gen256.txt 2>&1
timeout -s KILL 200 python tools/step_breakdown.py --workload gen512 > $R/r2_step_breakdown_gen512.txt 2>&1
