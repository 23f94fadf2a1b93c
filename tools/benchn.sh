# usage: benchn.sh TAG N TRANSPORT [extra]
TAG=$1; N=$2; T=$3; shift 3
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --transport $T "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/${TAG}.json").read().strip().splitlines()[-1])
print("${TAG}", d["n_gpus"], "%.4g" % d["value"], "%.4f ms" % d["ms_per_step"])
print({k: round(v["ms_per_tick"],4) for k,v in d.get("kernels",{}).items()})
EOF
