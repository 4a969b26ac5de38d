#!/bin/bash
# Run on the GPU box (under gpurun): one light ncu pass (a handful of metrics, no replay-heavy sections) over every kernel
# the library launches — the bench legs plus the GPU tests of the codecs the bench does not touch — so profiles/ holds
# time / instructions / DRAM bytes / occupancy for ALL kernels, not only the hot ones.  $1 = tag (e.g. r2)
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none -c 600 --csv --log-file $O/${TAG}_all_bench.csv \
    python bench.py --gb 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --rans-slices 1 --tok3-blocks 1184 --cram-tiles 4 > $O/${TAG}_all_bench.log 2>&1
ncu --metrics $M --clock-control none -c 4000 --csv --log-file $O/${TAG}_all_tests.csv \
    python -m pytest -q -m gpu tests/test_gpu_arith.py tests/test_gpu_arith_enc.py tests/test_gpu_rans4x8.py tests/test_gpu_rans4x8_enc.py \
        tests/test_gpu_fqzcomp.py tests/test_gpu_fqzcomp_enc.py tests/test_gpu_tok3_enc.py tests/test_gpu_sam_format.py tests/test_gpu_bgzf_compress.py \
        tests/test_gpu_xform.py tests/test_gpu_bam_pack.py tests/test_gpu_rans_enc.py -x > $O/${TAG}_all_tests.log 2>&1
tail -2 $O/${TAG}_all_tests.log
wc -l $O/${TAG}_all_bench.csv $O/${TAG}_all_tests.csv
