#!/bin/bash
# quick experiment pass: parity suite + the kernel timings that matter
tag=${1:-exp}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -x ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -4 gpurun_out/${tag}_pytest.log | head -1)"
timeout 300 python tools/config_bench.py --only ${2:-c1,c2,c3,c4,c4_aniso,c4_tilted,c4_tilted_biaxial} --json gpurun_out/${tag}_config_bench.jsonl > gpurun_out/${tag}_config_bench.log 2>&1; echo "cfg rc=$?"
cut -c1-260 gpurun_out/${tag}_config_bench.log | sed 's/"shape_FABT.*"power"/"power"/'
if [ -n "$3" ]; then timeout 600 bash tools/ncu_round.sh ${tag} $3 > gpurun_out/${tag}_ncu_round.log 2>&1; echo "ncu rc=$?"; fi
