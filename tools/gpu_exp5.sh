#!/bin/bash
tag=${1:-exp}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=8 ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -5 gpurun_out/${tag}_pytest.log | head -1)"
show() { cut -c1-300 $1 | sed 's/"shape_FABT.*"power"/"power"/; s/"finite_fraction.*//; s/"dtype": "c128", //'; }
timeout 300 python tools/config_bench.py --only c5,c5_scaled_nopower,c5_scaled > gpurun_out/${tag}_config_bench.log 2>&1; echo "cfg rc=$?"; show gpurun_out/${tag}_config_bench.log
