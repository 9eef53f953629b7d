# one 8-GPU box: config-4 scene sharded over N = 8, 4, 2 GPUs (bench.py under torchrun), default variant only
T=${1:-r2b}
run() {  # N tag extra-args
  N=$1; TAG=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/${T}_scale_${TAG}.json 2> gpurun_out/${T}_scale_${TAG}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_scale_${TAG}.json"))
    print("${TAG}", d["n_gpus"], round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), d["collectives"]["feature_gather"]["gbs_received_per_gpu"], d.get("parity_checked"))
except Exception as e:
    print("${TAG} failed", e)
PY
}
run 8 n8
run 4 n4
run 2 n2
tail -3 gpurun_out/${T}_scale_n8.err
