set -x
mkdir -p gpurun_out/r2d
nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q -x --timeout 300 > gpurun_out/r2d/pytest_comm.log 2>&1; tail -30 gpurun_out/r2d/pytest_comm.log
for mode in own nccl; do
  if [ $mode = nccl ]; then export SN_COMM=nccl; else unset SN_COMM; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d/bench_n2_$mode.json 2> gpurun_out/r2d/bench_n2_$mode.err; tail -3 gpurun_out/r2d/bench_n2_$mode.err; cat gpurun_out/r2d/bench_n2_$mode.json | cut -c1-900
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sync-bn > gpurun_out/r2d/bench_n2_${mode}_nosync.json 2> gpurun_out/r2d/bench_n2_${mode}_nosync.err; cat gpurun_out/r2d/bench_n2_${mode}_nosync.json | cut -c1-400
done
unset SN_COMM
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d/bench_n1.json 2> gpurun_out/r2d/bench_n1.err; cat gpurun_out/r2d/bench_n1.json | cut -c1-400
