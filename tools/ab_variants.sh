for q in 0 1; do
  for n in 4096 16384; do
    SBX_REAL=1 SBX_Q2=$q python tools/profile_kernel.py --n $n --pad 2 --hop $((n/16)) --frames 131072 --iters 5 2>&1 | tail -1 | sed "s/^/real Q2=$q /"
  done
  SBX_Q2=$q python tools/profile_kernel.py --n 2048 --pad 2 --hop 256 --frames 262144 --iters 5 2>&1 | tail -1 | sed "s/^/pair Q2=$q /"
done
