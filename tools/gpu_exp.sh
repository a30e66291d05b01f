#!/bin/bash
# Quick experiment pass under gpurun (1 GPU):  bash tools/gpu_exp.sh <tag> "<configs>" "<lib variants>" "<ncu captures>"
#   parity suite, tools/config_bench.py on the named configs with the in-tree library and with every
#   hyperbolic_optics_b200/libho_b200_<variant>.so, then the named captures of tools/ncu_round.sh.
tag=${1:-exp}
cfgs=${2:-c1,c2,c3,c4,c4_aniso,c4_tilted,c4_tilted_biaxial}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=8 ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -5 gpurun_out/${tag}_pytest.log | head -1)"
show() { cut -c1-300 $1 | sed 's/"shape_FABT.*"power"/"power"/; s/"finite_fraction.*//; s/"dtype": "c128", //'; }
timeout 300 python tools/config_bench.py --only $cfgs --json gpurun_out/${tag}_config_bench.jsonl > gpurun_out/${tag}_config_bench.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench.log
for v in $3; do
  echo "== variant $v"
  HO_B200_LIB=hyperbolic_optics_b200/libho_b200_$v.so timeout 300 python tools/config_bench.py --only $cfgs > gpurun_out/${tag}_config_bench_$v.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench_$v.log
done
if [ -n "$4" ]; then timeout 600 bash tools/ncu_round.sh ${tag} $4 > gpurun_out/${tag}_ncu_round.log 2>&1; echo "ncu rc=$?"; fi
