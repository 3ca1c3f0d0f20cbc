#!/bin/bash
# usage (GPU box): tools/ab_run.sh <variant> ...   - alternates the regular build with each libprl_b200_<variant>.so, 3 rounds
D=$PWD/parallel-reinforcement-learning_b200/prl_b200
for i in 1 2 3; do
  echo -n "base   "; env ${PRL_AB_MODE:-PRL_PROF_GRAPH}=1 python tools/prof_update.py step 65536 5 2>&1 | tail -1
  for v in "$@"; do
    echo -n "$v     "; PRL_B200_LIB=$D/libprl_b200_$v.so env ${PRL_AB_MODE:-PRL_PROF_GRAPH}=1 python tools/prof_update.py step 65536 5 2>&1 | tail -1
  done
done
