#!/bin/bash
# GPU box: one giant block through the shipped library and every variant in zopfli_b200/_var/ (same output expected),
# then an ncu source-level capture of the shipped k_iterate.   usage: tools/gpu_variants.sh [bytes] [iterations]
N=${1:-1000000}; IT=${2:-5}
mkdir -p gpurun_out
{
echo "== base"; timeout 120 python tools/one_block.py $N $IT 2>&1 | tail -3
for f in zopfli_b200/_var/lib_*.so; do
  echo "== $f"; ZB_LIB=$PWD/$f timeout 120 python tools/one_block.py $N $IT 2>&1 | tail -3
done
} > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
if [ -n "$NCU" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_iterate -c 1 -f -o gpurun_out/k_iterate_cur python tools/one_block.py $N 3 > gpurun_out/ncu.log 2>&1
  python tools/ncu_summary.py gpurun_out/k_iterate_cur.ncu-rep gpurun_out/k_iterate_cur.txt >> gpurun_out/ncu.log 2>&1
  tail -3 gpurun_out/ncu.log
fi
