#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; one GPU).  Every ncu capture follows a plain run of the
# same command that exited 0; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
python bench.py --steps 3 --warmup 1 --no-cpu --no-pageable --no-secondary > $O/r02_ncu_plain_bench.json 2> $O/r02_ncu_plain_bench.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --no-secondary > $O/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/sortbench.py 250000000 32 "3:0:8:40" c3 > $O/r02_ncu_plain_sort.log 2>&1 || { echo "plain sortbench failed"; exit 1; }
# the record passes of the third build (2 warm builds = 10 pass launches + the first from-sequence pass of the third)
ncu --set full --clock-control none --import-source on -k regex:scatter_pass_kernel --launch-skip 11 --launch-count 4 -f -o $O/r02_sort_pass \
    python tools/sortbench.py 250000000 32 "3:0:8:40" c3 > $O/r02_ncu_sort.log 2>&1
echo "sort pass capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:scatter_pass_kernel --launch-skip 10 --launch-count 1 -f -o $O/r02_sort_pass_seq \
    python tools/sortbench.py 250000000 32 "3:0:8:40" c3 > $O/r02_ncu_sort_seq.log 2>&1
echo "first pass capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rle_kernel --launch-skip 2 --launch-count 1 -f -o $O/r02_rle \
    python tools/sortbench.py 250000000 32 "3:0:8:40" c3 > $O/r02_ncu_rle.log 2>&1
echo "rle capture rc=$?"
python tools/probebench.py > $O/r02_ncu_plain_probe.log 2>&1 || { echo "plain probebench failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:probe_lookup_kernel --launch-skip 1 --launch-count 1 -f -o $O/r02_probe_lookup \
    python tools/probebench.py > $O/r02_ncu_probe.log 2>&1
echo "probe lookup capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:probe_emit_kernel --launch-skip 0 --launch-count 1 -f -o $O/r02_probe_emit \
    python tools/probebench.py > $O/r02_ncu_probe_emit.log 2>&1
echo "probe emit capture rc=$?"
ls -la $O/*.ncu-rep
