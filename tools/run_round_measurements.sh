# Round-2 measurement pass (run under gpurun; every step has its own hard timeout; outputs land in gpurun_out/ and are copied to profiles/ by hand).
set -x
R=gpurun_out
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > $R/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $R/r2_smoke.log
timeout -s KILL 700 python -m pytest tests -m gpu -q > $R/r2_pytest_gpu.log 2>&1
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > $R/r2_bench_gen256.json 2> $R/r2_bench_gen256.err
timeout -s KILL 300 python bench.py --steps 200 --warmup 5 --no-extras --skip-cpu-baseline > $R/r2_bench_gen256_long.json 2> $R/r2_bench_gen256_long.err
timeout -s KILL 300 python bench.py --workload gen512 --steps 20 --warmup 5 --skip-cpu-baseline --no-extras > $R/r2_bench_gen512.json 2> $R/r2_bench_gen512.err
timeout -s KILL 300 python tools/bench_conv.py --out $R/r2_conv_microbench.jsonl > /dev/null 2>&1
timeout -s KILL 300 python tools/microbench.py --out $R/r2_op_microbench.jsonl > /dev/null 2>&1
timeout -s KILL 300 python baseline/run_reference.py --mode ops > $R/r2_reference_cuda_ops.json 2> $R/r2_reference_cuda_ops.err
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > $R/r2_bench_reference_arm.json 2> $R/r2_bench_reference_arm.err
timeout -s KILL 300 python baseline/run_reference.py --mode ops_cpu > $R/r2_reference_cpu_ops.json 2> $R/r2_reference_cpu_ops.err
timeout -s KILL 240 python tools/bench_train.py --steps 16 --warmup 17 --out $R/r2_train_1gpu_fp32lib.json > /dev/null 2>&1
timeout -s KILL 240 python tools/bench_train.py --steps 16 --warmup 17 --allow-tf32 1 --out $R/r2_train_1gpu_tf32lib.json > /dev/null 2>&1
timeout -s KILL 300 python tools/prof_train.py > $R/r2_prof_train_fp32lib.log 2>&1
timeout -s KILL 300 python -m pytest tests/test_training_parity.py -m gpu -q -s > $R/r2_training_parity_pytest.log 2>&1     # prints the per-phase worst gradient errors
timeout -s KILL 200 python tools/step_breakdown.py > $R/r2_step_breakdown_gen256.txt 2>&1
timeout -s KILL 200 python tools/step_breakdown.py --workload gen512 > $R/r2_step_breakdown_gen512.txt 2>&1
# ncu passes (each only after the same command has exited 0 without ncu above): launch list of a 2-step eager bench run, and --set full on the op driver
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $R/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-extras --skip-cpu-baseline > $R/r2_ncu_bench.log 2>&1
timeout -s KILL 200 python tools/prof_ops.py > $R/r2_prof_ops_plain.log 2>&1 && \
timeout -s KILL 700 ncu --set full --clock-control none -k regex:"bias_act|upfirdn2d|conv_igemm|conv_rowfold_kernel|conv_wgrad|torgb|instance_stats|warp_perspective|patch_denorm" -c 60 -o /tmp/r2_ops_full -f python tools/prof_ops.py > $R/r2_ncu_ops.log 2>&1
python tools/ncu_summary.py /tmp/r2_ops_full.ncu-rep > $R/r2_kernels_ncu_full.csv 2>> $R/r2_ncu_ops.log
python tools/summarize_launches.py $R/r2_bench_launches.csv 40 > $R/r2_bench_launches_summary.txt 2>&1
