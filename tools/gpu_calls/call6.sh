set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r2_s5_C2.json 2> gpurun_out/r2_s5_C2.err; tail -5 gpurun_out/r2_s5_C2.err
python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/r2_s5_C3.json 2> gpurun_out/r2_s5_C3.err; tail -5 gpurun_out/r2_s5_C3.err
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or tile_path or stream or packed or group or baseline_configs" > gpurun_out/r2_gputest_5.log 2>&1; tail -3 gpurun_out/r2_gputest_5.log
python - <<'PY'
import json
for f in ('gpurun_out/r2_s5_C2.json','gpurun_out/r2_s5_C3.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e ms", d["e2e"]["ms_per_step"])
PY
