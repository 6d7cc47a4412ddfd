# 2-GPU check of the data-parallel step: pipelined e2e, bucketed vs single all-reduce
set -x
N=${1:-2}
timeout 300 python -m pytest tests/test_unet_gpu.py -x -q -k "pipelined or full_size_step_properties" 2>&1 | tail -3
timeout 300 python bench.py --config 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/dp_n1_c2.json 2> gpurun_out/dp_n1_c2.err; cut -c1-400 gpurun_out/dp_n1_c2.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for b in 1 0; do
  B200_DP_BUCKETS=$b timeout 400 $TR bench.py --gpus $N --config 2 --steps 30 --warmup 5 > gpurun_out/dp_n${N}_c2_b$b.json 2> gpurun_out/dp_n${N}_c2_b$b.err; cut -c1-300 gpurun_out/dp_n${N}_c2_b$b.json
done
timeout 400 $TR bench.py --gpus $N --config 3 --steps 20 --warmup 5 > gpurun_out/dp_n${N}_c3.json 2> gpurun_out/dp_n${N}_c3.err; cut -c1-300 gpurun_out/dp_n${N}_c3.json
