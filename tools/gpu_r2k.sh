set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r2k_dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check" gpurun_out/r2k_dist_check.log; tail -3 gpurun_out/r2k_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --songs 64 --no-cpu-baseline > gpurun_out/r2k_bench_n2_64songs.json 2> gpurun_out/r2k_bench_n2.err; echo "bench n2 rc=$?"
tail -c 500 gpurun_out/r2k_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 3 --no-cpu-baseline > gpurun_out/r2k_bench_n2_config3.json 2> gpurun_out/r2k_bench_n2_config3.err; echo "bench n2 c3 rc=$?"
tail -c 500 gpurun_out/r2k_bench_n2_config3.err
python - <<'PY'
import json
for f in ('gpurun_out/r2k_bench_n2_64songs.json', 'gpurun_out/r2k_bench_n2_config3.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['api'][:80], d['clocks'])
    except Exception as e: print('no json', f, e)
PY
