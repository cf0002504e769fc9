python -m pytest tests -m gpu -q > gpurun_out/u17_pytest.log 2>&1; tail -3 gpurun_out/u17_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-cuda-eager > gpurun_out/u17_n2.json 2> gpurun_out/u17_n2.err; cut -c1-150 gpurun_out/u17_n2.json; python -c "
import json; d=json.loads([l for l in open('gpurun_out/u17_n2.json') if l.startswith('{')][-1]); print(d['value'], d['parity_check'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-cuda-eager --batch 4096 --z-dim 100 --obs-dim 17 > gpurun_out/u17_n2_cfg5.json 2> gpurun_out/u17_n2_cfg5.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/u17_n2_cfg5.json') if l.startswith('{')][-1]); print(d['value'], d['parity_check'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/u17_ref_n2.json 2> gpurun_out/u17_ref_n2.err; cut -c1-300 gpurun_out/u17_ref_n2.json
tail -2 gpurun_out/u17_n2.err | cut -c1-200
