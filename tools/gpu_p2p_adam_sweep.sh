# 8-GPU sweep of the CTA count of k_p2p_adam (gpurun --gpus 8)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
COMMON="--gpus 8 --steps 200 --warmup 10 --no-cpu-baseline --no-cuda-eager --no-parity-check"
for n in 296 148 74; do
  FB_P2P_ADAM_CTAS=$n $TR bench.py $COMMON --timeline gpurun_out/p2p_adam_$n.txt > gpurun_out/p2p_adam_$n.json 2>/dev/null
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/p2p_adam_$n.json') if l.startswith('{')][-1]); print($n, round(d['value'],1), d['ms_per_step'])"
  grep adam gpurun_out/p2p_adam_$n.txt
done
