#!/bin/bash
# One round's ncu evidence (run under gpurun, 1 GPU):  bash tools/ncu_round.sh <tag> [what ...]
#   what: c4 c4_scat tilted tilted_scat tsweep polar launches   (default: all)
# Each capture is `ncu --set full --clock-control none --import-source on` of exactly one timed step
# (warm-up launches skipped); reports land in gpurun_out/<tag>_<what>.ncu-rep.
tag=$1; shift
what=${@:-c4 c4_scat tilted tilted_scat tsweep polar launches}
NCU="ncu --set full --clock-control none --import-source on"
CB="python tools/config_bench.py --reps 1"
for w in $what; do
  case $w in
    # fast-path protocol: 2 launches per reflect (fast build + fix-up); 3 warm-ups + 1 timed per backend
    c4)          $NCU -k regex:reflect_kernel -s 6  -c 2 -o gpurun_out/${tag}_c4          -f $CB --only c4 ;;
    c4_scat)     $NCU -k regex:reflect_kernel -s 14 -c 2 -o gpurun_out/${tag}_c4_scat     -f $CB --only c4 ;;
    # general build alone: 1 launch per reflect
    tilted)      $NCU -k regex:reflect_kernel -s 3  -c 1 -o gpurun_out/${tag}_tilted      -f $CB --only c4_tilted ;;
    tilted_scat) $NCU -k regex:reflect_kernel -s 7  -c 1 -o gpurun_out/${tag}_tilted_scat -f $CB --only c4_tilted ;;
    tsweep)      $NCU -k regex:tsweep_kernel  -s 3  -c 1 -o gpurun_out/${tag}_tsweep      -f $CB --only c5_scaled_nopower ;;
    polar)       $NCU -k regex:polarimetry_kernel -s 3 -c 1 -o gpurun_out/${tag}_polar    -f python tools/polarimetry_bench.py --only "Mueller matrix" --reps 1 ;;
    launches)    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_plain.log 2>&1 &&
                 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
                     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_ncu.log 2>&1 ;;
  esac
  echo "== $w done rc=$?"
done
ls -la gpurun_out/${tag}_* | head -20
