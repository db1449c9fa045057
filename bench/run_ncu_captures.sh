#!/usr/bin/env bash
# ncu captures of every kernel family on ONE GPU (never under a multi-rank command), summaries into profiles/.
#   gpurun --timeout 900 -- 'bash bench/run_ncu_captures.sh'   then, here:  bash bench/run_ncu_captures.sh --summarise
set -u
OUT=gpurun_out
mkdir -p $OUT profiles
NCU="ncu --set full --clock-control none --import-source on -c 2"
declare -A K=( [allreduce]="regex:^.*allreduce_kernel" [inplace]="regex:inplace_allreduce_kernel" [exchange]="regex:exchange_kernel"
               [exchange_tma]="regex:exchange_tma_kernel" [pipelined]="regex:pipelined_allreduce_kernel" [adasum]="regex:adasum_"
               [optim_sgd]="regex:fused_sgd_kernel" [optim_adam]="regex:fused_adam_kernel" )
declare -A F=( [allreduce]=allreduce [inplace]=inplace [exchange]=exchange [exchange_tma]=exchange [pipelined]=pipelined [adasum]=adasum
               [optim_sgd]=optim [optim_adam]=optim )
if [ "${1:-}" = "--summarise" ]; then
  for name in "${!K[@]}"; do
    [ -f $OUT/prof_$name.ncu-rep ] && python bench/ncu_summary.py $OUT/prof_$name.ncu-rep profiles/ncu_$name.md "$name (one simulated rank, 64 MiB)"
  done
  exit 0
fi
for name in allreduce inplace exchange exchange_tma pipelined adasum optim_sgd optim_adam; do
  extra=""
  [ "$name" = "exchange_tma" ] && extra="HVD_EXCHANGE_TMA=1"
  env $extra timeout 200 $NCU -k "${K[$name]}" -o $OUT/prof_$name python bench/ncu_targets.py --only ${F[$name]} 2>&1 | tail -2
done
ls -la $OUT/*.ncu-rep
