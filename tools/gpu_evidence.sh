#!/bin/bash
# One round's evidence under gpurun (1 GPU), in two stages because bench.py's roofline needs the ncu-counted
# constants of exactly this library build (profiles/ncu_constants.json, refreshed in between by
# tools/refresh_profiles.py in the build container):
#   bash tools/gpu_evidence.sh <tag> ncu     GPU test suite, config / polarimetry benches, the ncu set
#   bash tools/gpu_evidence.sh <tag> bench   bench.py: default line, the two other workloads, the reference arm
tag=${1:-r02}; stage=${2:-ncu}
mkdir -p gpurun_out
if [ $stage = ncu ]; then
  ( time timeout 1100 python -m pytest tests -m gpu -q --maxfail=8 --durations=12 ) > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$? $(tail -5 gpurun_out/${tag}_pytest.log | head -1)"
  timeout 300 python tools/config_bench.py --json gpurun_out/${tag}_config_bench.jsonl > gpurun_out/${tag}_config_bench.log 2>&1; echo "cfg rc=$?"
  timeout 200 python tools/config_bench.py --dtype c64 --only c4 --json gpurun_out/${tag}_config_bench_c64.jsonl > gpurun_out/${tag}_config_bench_c64.log 2>&1; echo "cfg64 rc=$?"
  timeout 200 python tools/polarimetry_bench.py --json gpurun_out/${tag}_polarimetry_bench.jsonl > gpurun_out/${tag}_polarimetry_bench.log 2>&1; echo "pol rc=$?"
  timeout 900 bash tools/ncu_round.sh ${tag} > gpurun_out/${tag}_ncu_round.log 2>&1; echo "ncu rc=$?"
  tail -12 gpurun_out/${tag}_ncu_round.log
else
  timeout 400 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"
  for w in c4_aniso c4_tilted; do
    timeout 200 python bench.py --workload $w --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; echo "bench $w rc=$?"
  done
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-seconds 90 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "ref rc=$?"
  cut -c1-400 gpurun_out/${tag}_bench_default.json
fi
