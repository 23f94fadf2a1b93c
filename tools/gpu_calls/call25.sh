set -x
timeout 1200 python -m pytest tests -m gpu -x -q -k "strips or group or two_strips or rings" > gpurun_out/r2_v6_gputest_2gpu.log 2>&1; tail -5 gpurun_out/r2_v6_gputest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub > gpurun_out/r2_v6_bench_C2_n2.json 2> gpurun_out/r2_v6_bench_C2_n2.err; tail -5 gpurun_out/r2_v6_bench_C2_n2.err
ABM_OVERLAP_DECISIONS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub > gpurun_out/r2_v6_bench_C2_n2_nodec.json 2> gpurun_out/r2_v6_bench_C2_n2_nodec.err; tail -5 gpurun_out/r2_v6_bench_C2_n2_nodec.err
python - <<'PY'
import json
for f in ('r2_v6_bench_C2_n2','r2_v6_bench_C2_n2_nodec'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "median %.4f"%d["ms_per_step_median"], {k: round(v["ms_per_tick"],4) for k,v in d["kernels"].items()}, "e2e ms %.3f"%d["e2e"]["ms_per_step"], d.get("verify",{}).get("status"))
PY
