set -x
timeout 1200 python -m pytest tests -m gpu -x -q -k "stream or walkmap or strips or golden or baseline" > gpurun_out/r2_v9_gputest_2gpu.log 2>&1; tail -5 gpurun_out/r2_v9_gputest_2gpu.log
python bench.py --workload C3 --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/r2_v9_bench_C3.json 2> gpurun_out/r2_v9_bench_C3.err; tail -3 gpurun_out/r2_v9_bench_C3.err
python bench.py --workload C5 --steps 20 --warmup 5 --no-sub --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_v9_bench_C5_n1.json 2> gpurun_out/r2_v9_bench_C5_n1.err; tail -3 gpurun_out/r2_v9_bench_C5_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload C5 --steps 20 --warmup 5 --no-sub --e2e-steps 0 > gpurun_out/r2_v9_bench_C5_n2.json 2> gpurun_out/r2_v9_bench_C5_n2.err; tail -5 gpurun_out/r2_v9_bench_C5_n2.err
python - <<'PY'
import json
for f in ('r2_v9_bench_C3','r2_v9_bench_C5_n1','r2_v9_bench_C5_n2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "median %.4f"%d["ms_per_step_median"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, d.get("verify",{}).get("status"))
PY
