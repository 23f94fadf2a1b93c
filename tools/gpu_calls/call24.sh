set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_v4_gputest_2gpu.log 2>&1; tail -5 gpurun_out/r2_v4_gputest_2gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r2_v4_bench_C2.json 2> gpurun_out/r2_v4_bench_C2.err; tail -3 gpurun_out/r2_v4_bench_C2.err
ABM_EARLY_BIRTHS=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r2_v4_bench_C2_late.json 2> gpurun_out/r2_v4_bench_C2_late.err; tail -3 gpurun_out/r2_v4_bench_C2_late.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub > gpurun_out/r2_v4_bench_C2_n2.json 2> gpurun_out/r2_v4_bench_C2_n2.err; tail -5 gpurun_out/r2_v4_bench_C2_n2.err
ABM_EARLY_BIRTHS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub > gpurun_out/r2_v4_bench_C2_n2_late.json 2> gpurun_out/r2_v4_bench_C2_n2_late.err; tail -5 gpurun_out/r2_v4_bench_C2_n2_late.err
python - <<'PY'
import json
for f in ('r2_v4_bench_C2','r2_v4_bench_C2_late','r2_v4_bench_C2_n2','r2_v4_bench_C2_n2_late'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "median %.4f"%d["ms_per_step_median"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e ms %.3f"%d["e2e"]["ms_per_step"], d.get("verify",{}).get("status"))
PY
