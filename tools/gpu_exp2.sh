#!/bin/bash
tag=${1:-exp}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -x ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -5 gpurun_out/${tag}_pytest.log | head -1)"
show() { cut -c1-260 $1 | sed 's/"shape_FABT.*"power"/"power"/; s/"finite_fraction.*//'; }
timeout 300 python tools/config_bench.py --only c1,c2,c3,c4,c4_aniso,c4_tilted,c4_tilted_biaxial > gpurun_out/${tag}_config_bench.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench.log
for v in $2; do
  echo "== variant $v"
  HO_B200_LIB=hyperbolic_optics_b200/libho_b200_$v.so timeout 300 python tools/config_bench.py --only c1,c3,c4,c4_aniso,c4_tilted > gpurun_out/${tag}_config_bench_$v.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench_$v.log
done
