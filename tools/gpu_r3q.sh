mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_check.py > gpurun_out/r3q_dist_check_2gpu.log 2>&1; echo "dist_check rc=$?"; grep "dist_check" gpurun_out/r3q_dist_check_2gpu.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r3q_bench_n2.json 2> gpurun_out/r3q_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --config 3 --no-cpu-baseline > gpurun_out/r3q_bench_n2_config3.json 2> gpurun_out/r3q_bench_n2_config3.err; echo "bench n2 config3 rc=$?"
python - <<PY
import json
for f in ('r3q_bench_n2', 'r3q_bench_n2_config3'):
    txt = [l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1]
    d = json.loads(txt)
    print(f, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 1), d['n_gpus'], d['scaling'], d['config'].get('sharding'))
PY
