# usage: bash tools/multi_gpu_run.sh N TAG   (on a box with N GPUs)
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q > gpurun_out/${TAG}_sharded_tests_${N}gpu.txt 2>&1; tail -2 gpurun_out/${TAG}_sharded_tests_${N}gpu.txt
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; tail -c 400 gpurun_out/${TAG}_bench_${N}gpu.err
timeout 600 $TR --master-port 29512 tools/bench_configs.py --only 5 > gpurun_out/${TAG}_config5_${N}gpu.json 2> gpurun_out/${TAG}_config5_${N}gpu.err
timeout 600 $TR --master-port 29513 examples/pointnet2_ae_step.py --steps 20 --warmup 3 > gpurun_out/${TAG}_config4_${N}gpu.json 2> gpurun_out/${TAG}_config4_${N}gpu.err
tail -c 300 gpurun_out/${TAG}_config4_${N}gpu.json
