set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_C2.json 2> gpurun_out/r2_final_bench_C2.err; tail -3 gpurun_out/r2_final_bench_C2.err
python bench.py --workload C3 --steps 20 --warmup 5 --no-sub > gpurun_out/r2_final_bench_C3.json 2> gpurun_out/r2_final_bench_C3.err; tail -3 gpurun_out/r2_final_bench_C3.err
python bench.py --workload V6 --steps 5 --warmup 2 --no-sub --e2e-steps 0 --cpu-ticks 1 > gpurun_out/r2_final_bench_V6.json 2> gpurun_out/r2_final_bench_V6.err; tail -3 gpurun_out/r2_final_bench_V6.err
python bench.py --stationary --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/r2_final_bench_C2_stationary.json 2> gpurun_out/r2_final_bench_C2_stationary.err; tail -3 gpurun_out/r2_final_bench_C2_stationary.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches_C2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub --e2e-steps 0 > gpurun_out/ncu_l_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches_C3.csv python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-sub --e2e-steps 0 > gpurun_out/ncu_l_c3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cta --launch-skip 12 --launch-count 4 -o gpurun_out/r2_final_C2_full python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 0 --no-sub > gpurun_out/ncu_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cta_lb --launch-skip 3 --launch-count 1 -o gpurun_out/r2_final_C3_stream python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 0 --no-sub > gpurun_out/ncu_c3.log 2>&1
python - <<'PY'
import json
for f in ('r2_final_bench_C2','r2_final_bench_C3','r2_final_bench_V6','r2_final_bench_C2_stationary'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e ms", d.get("e2e",{}).get("ms_per_step"), "roof", d["roofline"]["frac"] if d.get("roofline") else None, d["roofline_tick"]["frac"])
    for k in ("configs","c5"):
        if k in d:
            sub = d[k]["C3"] if k=="configs" else d[k]
            print("  ",k, sub if "error" in sub else ("value %.4g"%sub["value"], "ms %.4f"%sub["ms_per_step"], {kk: round(v["ms_per_tick"],4) for kk,v in sub["kernels"].items()}, sub["roofline_tick"]["frac"]))
    print("  ", d.get("cpu_baseline"))
PY
