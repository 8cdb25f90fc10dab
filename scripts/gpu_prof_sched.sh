#!/usr/bin/env bash
# ncu --set full of one launch per (profile, scheduler): PROFS="v2 v4_equirect" SCHEDS="sorted lane"
set -u
mkdir -p gpurun_out
for prof in ${PROFS:-v2 v4_equirect}; do
  for sched in ${SCHEDS:-sorted}; do
    export B200PT_SCHEDULER=$sched
    python scripts/prof_any.py $prof 128 3 > gpurun_out/plain_${prof}_${sched}.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:pt_render -s 2 -c 1 -o gpurun_out/prof_${prof}_${sched} -f \
        python scripts/prof_any.py $prof 128 3 > gpurun_out/ncu_${prof}_${sched}.log 2>&1
    tail -1 gpurun_out/plain_${prof}_${sched}.log; tail -2 gpurun_out/ncu_${prof}_${sched}.log
  done
done
ls -la gpurun_out/*.ncu-rep
