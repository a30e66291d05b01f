#!/bin/bash
tag=${1:-exp}
mkdir -p gpurun_out
show() { cut -c1-260 $1 | sed 's/"shape_FABT.*"power"/"power"/; s/"finite_fraction.*//'; }
for v in $2; do
  echo "== variant $v"
  HO_B200_LIB=hyperbolic_optics_b200/libho_b200_$v.so timeout 300 python tools/config_bench.py --only c1,c4,c4_tilted > gpurun_out/${tag}_config_bench_$v.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench_$v.log
done
timeout 300 python tools/config_bench.py --only c1,c4,c4_tilted > gpurun_out/${tag}_config_bench.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench.log
if [ -n "$3" ]; then timeout 600 bash tools/ncu_round.sh ${tag} $3 > gpurun_out/${tag}_ncu_round.log 2>&1; echo "ncu rc=$?"; fi
