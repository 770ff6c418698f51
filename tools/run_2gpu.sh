# two GPUs: sharded library tests on real devices, then bench.py under torchrun (weak + strong scaling in one line)
cd $GRAFT_REPO_ROOT
T=${1:-g2}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q -k "sharded" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err
tail -3 gpurun_out/${T}_bench_n${N}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_n${N}.json"))
print("weak value %.0f e2e %.0f ms/step %.2f pair %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]))
print(d["breakdown_ms"])
print("strong", {k: d["strong"][k] for k in ("value","ms_per_step","e2e","cells_per_gpu","rank0_ms")})
PY
