for i in 1 2 3; do
  echo -n "A(new) "; PRL_PROF_GRAPH=1 python tools/prof_update.py step 65536 5 2>&1 | tail -1
  echo -n "B(waits) "; PRL_B200_LIB=$PWD/parallel-reinforcement-learning_b200/prl_b200/libprl_b200_waits.so PRL_PROF_GRAPH=1 python tools/prof_update.py step 65536 5 2>&1 | tail -1
done
