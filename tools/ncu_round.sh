#!/bin/bash
# One round's ncu evidence (run under gpurun, 1 GPU):  bash tools/ncu_round.sh <tag> [what ...]
#   what: c4 c4_scat aniso tilted tilted_scat tsweep polar launches   (default: all)
# Each capture is `ncu --set full --clock-control none --import-source on` of exactly one timed step
# (warm-up launches skipped); reports land in gpurun_out/<tag>_<what>.ncu-rep.
tag=$1; shift
what=${@:-c4 c4_scat aniso tilted tsweep polar launches}
NCU="ncu --set full --clock-control none --import-source on"
CB="python tools/config_bench.py --reps 1"
for w in $what; do
  case $w in
    # fast-path protocol: 2 launches per reflect (fast build + fix-up); 3 warm-ups + 1 timed per backend
    c4)          $NCU -k regex:reflect_kernel -s 6  -c 2 -o gpurun_out/${tag}_c4          -f $CB --only c4 ;;
    c4_scat)     $NCU -k regex:reflect_kernel -s 14 -c 2 -o gpurun_out/${tag}_c4_scat     -f $CB --only c4 ;;
    aniso)       $NCU -k regex:reflect_kernel -s 6  -c 2 -o gpurun_out/${tag}_aniso       -f $CB --only c4_aniso ;;
    # lean build + fix-up: 2 launches per reflect, as c4
    tilted)      $NCU -k regex:reflect_kernel -s 6  -c 2 -o gpurun_out/${tag}_tilted      -f $CB --only c4_tilted ;;
    tilted_scat) $NCU -k regex:reflect_kernel -s 14 -c 2 -o gpurun_out/${tag}_tilted_scat -f $CB --only c4_tilted ;;
    tsweep)      $NCU -k regex:tsweep_kernel  -s 3  -c 1 -o gpurun_out/${tag}_tsweep      -f $CB --only c5_scaled_nopower ;;
    polar)       $NCU -k regex:polarimetry_kernel -s 3 -c 1 -o gpurun_out/${tag}_polar    -f python tools/polarimetry_bench.py --only "Mueller matrix" --reps 1 ;;
    launches)    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_plain.log 2>&1 &&
                 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
                     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_ncu.log 2>&1 ;;
  esac
  echo "== $w done rc=$?"
  # the reports (25-30 MB each with --import-source) do not fit gpurun's 64 MiB return limit:
  # export the pages the summarisers read, on the box, and drop the report
  rep=gpurun_out/${tag}_${w}.ncu-rep
  if [ -f $rep ]; then
    ncu -i $rep --page raw --csv > gpurun_out/${tag}_${w}_raw.csv 2>/dev/null
    ncu -i $rep --page source --csv | gzip -9 > gpurun_out/${tag}_${w}_source.csv.gz 2>/dev/null
    ncu -i $rep --page source --csv --print-source cuda,sass | gzip -9 > gpurun_out/${tag}_${w}_srcsass.csv.gz 2>/dev/null
    [ -z "$HO_KEEP_REP" ] && rm -f $rep
  fi
done
du -sh gpurun_out; ls gpurun_out | head -60
