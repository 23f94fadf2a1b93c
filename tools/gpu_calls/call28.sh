set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_C2_n8.json 2> gpurun_out/r2_final_bench_C2_n8.err; tail -5 gpurun_out/r2_final_bench_C2_n8.err
python - <<'PY'
import json
for f in ('r2_final_bench_C2_n8',):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "median %.4f"%d["ms_per_step_median"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("verify",{}).get("status"))
    c=d.get("c5"); print("   c5", c if "error" in c else ("value %.4g"%c["value"], "ms %.4f"%c["ms_per_step"], "median %.4f"%c["ms_per_step_median"], {k: round(v["ms_per_tick"],4) for k,v in c["kernels"].items()}))
PY
