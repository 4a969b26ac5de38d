#!/bin/bash
# Run on the GPU box (under gpurun): writes everything the judge reads into gpurun_out/, which is
# then copied to profiles/ by hand.  $1 = tag (e.g. r1)
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --gb 1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --rans-slices -1 > $O/${TAG}_launches_bench.log 2>&1
# 2. the two hot kernels, full sections + source
ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate_kernel -s 1 -c 1 -o $O/${TAG}_inflate \
    python bench.py --gb 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --rans-slices 0 > $O/${TAG}_inflate_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rans_nx16_decode_kernel -s 2 -c 2 -o $O/${TAG}_rans \
    python bench.py --gb 0.1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --rans-slices -1 > $O/${TAG}_rans_bench.log 2>&1
ncu --set full --clock-control none -k regex:bam_unpack_kernel -c 1 -o $O/${TAG}_bam \
    python bench.py --gb 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --rans-slices 0 > $O/${TAG}_bam_bench.log 2>&1
ls -la $O
# 3. the tok3 name-rebuild kernel (two blocks per warp), bench leg only
ncu --set full --clock-control none --import-source on -k regex:tok3_names -s 1 -c 1 -o $O/${TAG}_tok3 \
    python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --rans-slices 0 --tok3-blocks 4736 > $O/${TAG}_tok3_bench.log 2>&1
