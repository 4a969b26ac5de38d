#!/bin/bash
# tok3 GPU check: parity tests, then the bench leg at two batch sizes (optionally for an A/B library build)
timeout 600 python -m pytest tests/test_gpu_tok3.py -x -q 2>&1 | tail -4
for lib in "" $TOK3_AB_LIBS; do
  for nb in 1184 4736; do
    HGPU_LIB=${lib:+$PWD/htslib_b200/$lib} python bench.py --gb 0.5 --rans-slices 0 --no-cpu-baseline --no-e2e --steps 3 --tok3-blocks $nb 2>gpurun_out/tok3_bench.err |
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['extra'].get('tok3_decode'); print('${lib:-default}', {k:(round(t[k],2) if isinstance(t[k],float) else t[k]) for k in t if k in ('blocks','entropy_ms','rebuild_ms','device_GBps','e2e_GBps','error')})"
  done
done
tail -2 gpurun_out/tok3_bench.err
