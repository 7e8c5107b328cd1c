#!/bin/bash
# ncu evidence of the feature kernels for the build in the tree (run on the GPU box through gpurun):
#   tools/capture_profiles.sh <tag>     e.g. r02r
# writes text summaries gpurun_out/<tag>/ncu_features_<workload>.txt, profiles/traffic.json (DRAM bytes per launch,
# tied to the source hash) as gpurun_out/<tag>/traffic.json, and the launch list of one bench.py run.
set -u
tag=${1:-cap}
out=gpurun_out/$tag
mkdir -p $out /tmp/ncu_$tag
args=""
for w in melA melB linA mel512 mel256; do
  ncu --set full --clock-control none --import-source on -k regex:features -c 1 -f -o /tmp/ncu_$tag/$w python tools/kbench.py 1 $w > /dev/null 2>&1
  python tools/ncu_report.py /tmp/ncu_$tag/$w.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:features -c 1 python tools/kbench.py 1 $w ($tag build)" > $out/ncu_features_$w.txt 2>/dev/null
done
# n_fft 4096: phase streams through the 1024-point kernel + combine: the combine kernel and the launch list of one run
ncu --set full --clock-control none --import-source on -k regex:combine -c 1 -f -o /tmp/ncu_$tag/mel4096 python tools/kbench.py 1 mel4096 > /dev/null 2>&1
python tools/ncu_report.py /tmp/ncu_$tag/mel4096.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:combine -c 1 python tools/kbench.py 1 mel4096 ($tag build)" > $out/ncu_combine_mel4096.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/launches_mel4096.csv python tools/kbench.py 1 mel4096 > /dev/null 2>&1
python tools/ncu_traffic.py mel80_22k_1k_ragged=/tmp/ncu_$tag/melA.ncu-rep mel128_44k_1k_ragged=/tmp/ncu_$tag/melB.ncu-rep \
  linear_22k_1k_ragged=/tmp/ncu_$tag/linA.ncu-rep mel80_16k_nfft512=/tmp/ncu_$tag/mel512.ncu-rep > /dev/null 2>&1
cp profiles/traffic.json $out/traffic.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/bench_launches.csv python bench.py --steps 2 --warmup 1 > $out/bench_under_ncu.log 2>&1
ls -la $out
