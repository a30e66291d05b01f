#!/bin/bash
# multi-GPU pass: bash tools/gpu_multi.sh <tag> <N> [extra bench args]
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo_${n}gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" \
  > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/${tag}_bench_${n}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_${n}gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","value_gathered","ms_per_step","n_gpus")})
print("e2e", d["e2e"]["value"], "lazy", d["e2e"]["lazy_mueller"]["value"])
g=d.get("gather") or {}
print({k:g.get(k) for k in ("ms_per_step","transport","transport_error","GBps_received_per_rank","ok","chunks")})
print("weak", (d.get("weak") or {}).get("value"))
PY
if [ $n -eq 2 ]; then timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "owns_the_buffers" 2>&1 | tail -2; fi
if [ -n "$HO_D2H_PROBE" ]; then timeout 200 python tools/d2h_probe.py d2h > gpurun_out/${tag}_d2h_probe_${n}gpu.txt 2>&1; tail -15 gpurun_out/${tag}_d2h_probe_${n}gpu.txt; fi
