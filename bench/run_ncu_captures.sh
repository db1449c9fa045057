#!/usr/bin/env bash
# ncu captures of every kernel family on ONE GPU (never under a multi-rank command), summarised ON THE BOX into
# gpurun_out/ncu_<name>.md (+ the raw-page CSV); reports larger than 6 MiB are dropped afterwards so the whole directory
# stays under gpurun's 64 MiB copy-back limit.
#   gpurun --timeout 1200 -- 'bash bench/run_ncu_captures.sh'      then copy gpurun_out/ncu_*.md to profiles/
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=.
declare -A K=( [allreduce]="regex:^.*allreduce_kernel" [inplace]="regex:inplace_allreduce_kernel" [exchange]="regex:exchange_kernel"
               [adasum]="regex:adasum_" [optim_sgd]="regex:fused_sgd_kernel" [optim_adam]="regex:fused_adam_kernel" )
declare -A F=( [allreduce]=allreduce [inplace]=inplace [exchange]=exchange [adasum]=adasum [optim_sgd]=optim [optim_adam]=optim )
declare -A C=( [allreduce]=2 [inplace]=1 [exchange]=1 [adasum]=1 [optim_sgd]=1 [optim_adam]=1 )
for name in allreduce inplace exchange adasum optim_sgd optim_adam; do
  timeout 240 ncu --set full --clock-control none --import-source on -c ${C[$name]} -k "${K[$name]}" -o $OUT/prof_$name \
    python bench/ncu_targets.py --only ${F[$name]} 2>&1 | tail -1
  if [ -f $OUT/prof_$name.ncu-rep ]; then
    python bench/ncu_summary.py $OUT/prof_$name.ncu-rep $OUT/ncu_$name.md "$name (one simulated rank, 64 MiB message)" 2>&1 | tail -1
    ncu -i $OUT/prof_$name.ncu-rep --page raw --csv > $OUT/ncu_$name.raw.csv 2>/dev/null
    sz=$(stat -c %s $OUT/prof_$name.ncu-rep)
    [ "$sz" -gt 6291456 ] && rm -f $OUT/prof_$name.ncu-rep
  fi
done
ls -la $OUT | head -30
