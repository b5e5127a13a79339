python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-config3 > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err
echo "stdout lines: $(wc -l < gpurun_out/r02_bench_2gpu_final.json)"; head -c 150 gpurun_out/r02_bench_2gpu_final.json; echo
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 | head -c 150; echo
python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 5 2>/dev/null | head -c 200
