#!/bin/bash
# After `gpurun -- bash tools/gpu_evidence.sh <tag> ncu`: turn the exported ncu pages in gpurun_out/ into the
# tracked summaries under profiles/ and the per-point constants bench.py reads (build container, no GPU).
tag=${1:-r02}
C4=100761600; C5=37785600
gunzip -f gpurun_out/${tag}_*.csv.gz 2>/dev/null
python tools/refresh_profiles.py ${tag}_c4 ${tag}_c4 $C4 --workload c4 > /dev/null
python tools/refresh_profiles.py ${tag}_aniso ${tag}_aniso $C4 --workload c4_aniso > /dev/null
python tools/refresh_profiles.py ${tag}_tilted ${tag}_tilted $C4 --workload c4_tilted > /dev/null
python tools/refresh_profiles.py ${tag}_c4_scat ${tag}_c4_scat $C4 > /dev/null
python tools/refresh_profiles.py ${tag}_tsweep ${tag}_tsweep $C5 > /dev/null
python tools/refresh_profiles.py ${tag}_polar ${tag}_polar $C4 > /dev/null
for w in c4 tilted; do python tools/ncu_stalls.py gpurun_out/${tag}_${w}_srcsass.csv 25 stall_wait > profiles/${tag}_${w}_stalls_by_line.txt; done
for f in pytest.log config_bench.jsonl config_bench_c64.jsonl polarimetry_bench.jsonl launches.csv; do [ -f gpurun_out/${tag}_$f ] && cp gpurun_out/${tag}_$f profiles/${tag}_$f; done
cat profiles/ncu_constants.json
