#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; one GPU).  Every ncu capture follows a plain run of the
# same command that exited 0; numbers printed under ncu are never bench values.  gpurun brings back at most 64 MiB, so
# each .ncu-rep is reduced to its raw-metric CSV + summary on the box (tools/ncusum.py) and deleted.
set -u
O=gpurun_out
cap() {   # cap <name> <kernel regex> <skip> <count> <command...>
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$re --launch-skip $skip --launch-count $cnt -f -o /tmp/$name "$@" > $O/r02_ncu_${name}.log 2>&1
  echo "$name capture rc=$?"
  python tools/ncusum.py /tmp/$name.ncu-rep > $O/r02_ncu_${name}_summary.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/r02_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r02_ncu_${name}_source.csv.gz
  rm -f /tmp/$name.ncu-rep
}
python bench.py --steps 3 --warmup 1 --no-cpu --no-pageable --no-secondary > $O/r02_ncu_plain_bench.json 2> $O/r02_ncu_plain_bench.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --no-secondary > $O/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/sortbench.py 250000000 32 "3:-1:8:32,3:0:8:32,3:-1:8:40" c3 > $O/r02_sortbench_final_c3.log 2>&1 || { echo "plain sortbench failed"; exit 1; }
cat $O/r02_sortbench_final_c3.log
# record passes of the third build (2 warm builds = 8 pass launches at 32 sorted bits, then the from-sequence pass of the third)
cap sort_pass scatter_pass_kernel 9 3 python tools/sortbench.py 250000000 32 "3:-1:8:32" c3
cap sort_pass_seq scatter_pass_kernel 8 1 python tools/sortbench.py 250000000 32 "3:-1:8:32" c3
cap rle rle_kernel 2 1 python tools/sortbench.py 250000000 32 "3:-1:8:32" c3
cap group_detect group_detect_kernel 2 1 python tools/sortbench.py 250000000 32 "3:-1:8:32" c3
python tools/probebench.py > $O/r02_probebench3.log 2>&1 || { echo "plain probebench failed"; exit 1; }
cap probe_lookup probe_lookup_kernel 1 1 python tools/probebench.py
cap probe_emit probe_emit_kernel 0 1 python tools/probebench.py
du -sh $O
