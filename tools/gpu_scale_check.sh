# usage: tools/gpu_scale_check.sh N   -- bench line at N GPUs + one traced solve (per-rank wave timelines)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -c 300 gpurun_out/r02_bench_${N}gpu.err
bash tools/gpu_trace_ranks.sh $N
head -c 900 gpurun_out/r02_bench_${N}gpu.json
