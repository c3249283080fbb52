mkdir -p gpurun_out/r2f
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --kernel-table gpurun_out/r2f/kt_sync.json > gpurun_out/r2f/b_sync.json 2> gpurun_out/r2f/b_sync.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sync-bn --kernel-table gpurun_out/r2f/kt_nosync.json > gpurun_out/r2f/b_nosync.json 2> gpurun_out/r2f/b_nosync.err
