#!/bin/bash
# scripts/r2_profile.sh -- the ncu captures summarised under profiles/ (run under gpurun on ONE GPU; each command has run
# once without ncu first).  Reports land in gpurun_out/; scripts/ncu_summary.py / ncu_lines.py turn them into text.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
for c in c2 c3 c4 c5; do python scripts/prof_one.py $c 2 > $O/plain_$c.log 2>&1 || exit 1; done
python bench.py --steps 3 --warmup 3 --only-headline --no-cpu-baseline > $O/plain_bench.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:primary_kernel -s 1 -c 1 -o $O/r02_c2 python scripts/prof_one.py c2 2 > $O/ncu_c2.log 2>&1
$NCU -k regex:"primary_kernel|tri_deferred" -s 2 -c 2 -o $O/r02_c4 python scripts/prof_one.py c4 2 > $O/ncu_c4.log 2>&1
$NCU -k regex:shade_expand -s 7 -c 7 -o $O/r02_c3 python scripts/prof_one.py c3 2 > $O/ncu_c3.log 2>&1
$NCU -k regex:shade_expand -s 24 -c 3 -o $O/r02_c5 python scripts/prof_one.py c5 2 > $O/ncu_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --only-headline --no-cpu-baseline > $O/ncu_bench.log 2>&1
# text summaries (what profiles/ keeps); the reports themselves exceed what gpurun brings back (64 MiB): keep only config 4's
mkdir -p $O/r02
for c in c2 c3 c4 c5; do
  python scripts/ncu_summary.py $O/r02_$c.ncu-rep > $O/r02/r02_${c}_ncu_raw.txt 2>&1
  python scripts/ncu_lines.py $O/r02_$c.ncu-rep 45 > $O/r02/r02_${c}_lines.txt 2>&1
done
cp $O/r02_launches_bench_c2.csv $O/r02/
rm -f $O/r02_c2.ncu-rep $O/r02_c3.ncu-rep $O/r02_c5.ncu-rep
ls -la $O/r02 $O/*.ncu-rep
