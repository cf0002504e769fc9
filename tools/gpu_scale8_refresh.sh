TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
COMMON="--gpus 8 --steps 200 --warmup 10 --no-cpu-baseline --no-cuda-eager"
$TR bench.py $COMMON --timeline gpurun_out/r2_scale8_timeline_cfg2_p2p.txt > gpurun_out/r2_bench_n8_cfg2_p2p.json 2> gpurun_out/n8_cfg2_p2p.err; cut -c1-160 gpurun_out/r2_bench_n8_cfg2_p2p.json
$TR bench.py $COMMON --batch 4096 --z-dim 100 --obs-dim 17 --timeline gpurun_out/r2_scale8_timeline_cfg5_p2p.txt > gpurun_out/r2_bench_n8_cfg5_p2p.json 2> gpurun_out/n8_cfg5_p2p.err; cut -c1-160 gpurun_out/r2_bench_n8_cfg5_p2p.json
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-cuda-eager --batch 4096 --z-dim 100 --obs-dim 17 --timeline gpurun_out/r2_timeline_n1_cfg5.txt > gpurun_out/r2_bench_n1_cfg5.json 2>/dev/null; cut -c1-160 gpurun_out/r2_bench_n1_cfg5.json
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-cuda-eager --batch 128 --timeline gpurun_out/r2_timeline_n1_batch128.txt > gpurun_out/r2_bench_n1_batch128.json 2>/dev/null; cut -c1-160 gpurun_out/r2_bench_n1_batch128.json
