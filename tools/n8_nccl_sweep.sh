run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 12 --warmup 3 --no-inference --no-stress > gpurun_out/bench_n8_$tag.json 2> gpurun_out/bench_n8_$tag.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_n8_$tag.json")); print("$tag", d["ms_per_step"], d["value"])
except Exception as e:
    print("$tag", "failed", e)
PY
}
run default X=1
run ll128 NCCL_PROTO=LL128
run nvls NCCL_ALGO=NVLS,Ring
