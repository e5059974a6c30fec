timeout 300 python -m pytest tests/test_gpu_post.py tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2l_bench2.err | tail -1 > gpurun_out/r2l_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_2gpu.json')); print('2 GPUs: fps %.1f e2e %.1f replica_check %s' % (d['value'], d['e2e']['value'], d.get('replica_check')))"
grep -c "NCCL INFO" gpurun_out/r2l_bench2.err; grep -E "nranks|Init COMPLETE" gpurun_out/r2l_bench2.err | head -4
tools/e2e_cli.sh 2 720 32
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
