set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_v2_bench_C2_n8.json 2> gpurun_out/r2_v2_bench_C2_n8.err; tail -5 gpurun_out/r2_v2_bench_C2_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_v2_bench_C2_n8.json').read().strip().splitlines()[-1])
print("value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["ms_per_step"], d.get("verify",{}).get("status"))
c=d.get("c5"); print("c5", c if "error" in c else ("value %.4g"%c["value"], "ms %.4f"%c["ms_per_step"], {k: round(v["ms_per_tick"],4) for k,v in c["kernels"].items()}, c["config"]["world_build_s"]))
PY
