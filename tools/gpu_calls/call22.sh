set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_v3_gputest_2gpu.log 2>&1; tail -5 gpurun_out/r2_v3_gputest_2gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub --scheduler randomly > gpurun_out/r2_v3_bench_C2_randomly.json 2> gpurun_out/r2_v3_bench_C2_randomly.err; tail -3 gpurun_out/r2_v3_bench_C2_randomly.err
python bench.py --workload C3 --steps 20 --warmup 5 --no-cpu-baseline --no-sub --scheduler randomly > gpurun_out/r2_v3_bench_C3_randomly.json 2> gpurun_out/r2_v3_bench_C3_randomly.err; tail -3 gpurun_out/r2_v3_bench_C3_randomly.err
python bench.py --workload C3 --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r2_v3_bench_C3.json 2> gpurun_out/r2_v3_bench_C3.err; tail -3 gpurun_out/r2_v3_bench_C3.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r2_v3_bench_C2.json 2> gpurun_out/r2_v3_bench_C2.err; tail -3 gpurun_out/r2_v3_bench_C2.err
python - <<'PY'
import json
for f in ('r2_v3_bench_C2','r2_v3_bench_C2_randomly','r2_v3_bench_C3','r2_v3_bench_C3_randomly'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e ms %.3f"%d["e2e"]["ms_per_step"])
PY
