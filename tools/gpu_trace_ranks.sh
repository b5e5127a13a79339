# usage: tools/gpu_trace_ranks.sh N   -- one traced config-2 solve at N GPUs (per-rank wave timelines: DDO_WAVE_TRACE)
N=$1
DDO_WAVE_TRACE=gpurun_out/r02_trace_${N}gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --no-config3 > gpurun_out/r02_trace_${N}gpu.json 2> gpurun_out/r02_trace_${N}gpu.err
tail -c 400 gpurun_out/r02_trace_${N}gpu.json
