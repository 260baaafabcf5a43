#!/bin/bash
# On the GPU box: one `ncu --set full` capture of a B=256 forward, digested in place (the .ncu-rep is ~40 MB and is not
# brought back): per-launch raw metrics CSV + per-kernel stall-reason digests (tools/ncu_stalls.py).
#   [MODE=as_shipped] tools/ncu_digest.sh <tag>
set -u
tag=$1
out=gpurun_out
rep=/tmp/${tag}.ncu-rep
N=1 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm|umma_atm|umma_ln2|attention_mma|rnn_umma|condition_kernel" -c 25 -o /tmp/${tag} python tools/one_forward.py > $out/${tag}_ncu.log 2>&1
ncu -i $rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py $rep $out/${tag}_summary.csv $out/${tag}_traffic.json > /dev/null 2>&1
: > $out/${tag}_stalls.txt
# launch order of one forward: 0 condition, 1 in_linear, then per layer qkv, attention, out_proj_ln, ff1, ff2_ln; 22 rnn_ih, 23 rnn, 24 head
i=0
for name in condition in_linear qkv attention out_proj_ln ff1 ff2_ln; do
  ncu -i $rep --page source --csv --launch-skip $i --launch-count 1 > /tmp/src.csv 2>/dev/null
  echo "=== launch $i: $name -- $(head -1 /tmp/src.csv | cut -c1-110)" >> $out/${tag}_stalls.txt
  python tools/ncu_stalls.py /tmp/src.csv 14 2>/dev/null | cut -c1-220 >> $out/${tag}_stalls.txt
  i=$((i+1))
done
for pair in "22 rnn_ih" "23 rnn" "24 head"; do
  set -- $pair
  ncu -i $rep --page source --csv --launch-skip $1 --launch-count 1 > /tmp/src.csv 2>/dev/null
  echo "=== launch $1: $2 -- $(head -1 /tmp/src.csv | cut -c1-110)" >> $out/${tag}_stalls.txt
  python tools/ncu_stalls.py /tmp/src.csv 14 2>/dev/null | cut -c1-220 >> $out/${tag}_stalls.txt
done
rm -f $rep /tmp/src.csv
ls -la $out/${tag}_*
