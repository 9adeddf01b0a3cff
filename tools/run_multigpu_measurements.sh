# Multi-GPU measurement pass (run under `gpurun --gpus 8`): configs[3] training step and configs[2] Generator_512 at 4 and 8 GPUs of one box.
# Each launch has its own hard timeout; outputs land in gpurun_out/ and are copied to profiles/ by hand.
set -x
R=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $R/r2_mg_gpus.txt 2>&1; nproc >> $R/r2_mg_gpus.txt
for n in 8 4; do
  timeout -s KILL 240 $TR --nproc-per-node $n --master-port $((29600+n)) tools/bench_train.py --steps 16 --warmup 17 --out $R/r2_train_${n}gpu.json > $R/r2_train_${n}gpu.log 2>&1
  timeout -s KILL 200 $TR --nproc-per-node $n --master-port $((29700+n)) bench.py --gpus $n --workload gen512 --steps 20 --warmup 5 --skip-cpu-baseline --no-extras > $R/r2_bench512_${n}gpu.json 2> $R/r2_bench512_${n}gpu.err
done
# 1- and 2-GPU points of the same commit on the same box (GPUs 0 | 1 alone would share nothing, but keep them serial so the host is quiet)
for n in 2 1; do
  timeout -s KILL 200 $TR --nproc-per-node $n --master-port $((29600+n)) tools/bench_train.py --steps 16 --warmup 17 --out $R/r2_train_${n}gpu_8box.json > $R/r2_train_${n}gpu_8box.log 2>&1
done
