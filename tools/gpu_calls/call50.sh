set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_end_gputest_1gpu.log 2>&1; tail -3 gpurun_out/r2_end_gputest_1gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_end_bench_C2.json 2> gpurun_out/r2_end_bench_C2.err; tail -3 gpurun_out/r2_end_bench_C2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_end_bench_C2.json').read().strip().splitlines()[-1])
print("value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e ms", d.get("e2e",{}).get("ms_per_step"), "roof", d["roofline"]["frac"], d["roofline_tick"]["frac"], d.get("clocks"))
for k in ("configs","c5"):
    if k in d:
        sub = d[k]["C3"] if k=="configs" else d[k]
        print("  ",k, sub if "error" in sub else ("value %.4g"%sub["value"], "ms %.4f"%sub["ms_per_step"], sub["roofline_tick"]["frac"]))
print("  ", d.get("cpu_baseline"))
PY
