#!/bin/bash
mkdir -p gpurun_out
TESTS="test_gpu_encoder" bash scripts/gpu_tests.sh | tail -4
for cfg in "KJC_LANES=1" "KJC_LANES=2" "KJC_LANES=3" "KJC_LANES=4" "KJC_LANES=2 KJC_MICRO_TOKENS=37888" "KJC_LANES=4 KJC_MICRO_TOKENS=9472"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --no-index --no-cpu --steps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['config']['workload'][-40:])"
done 2>&1 | tee gpurun_out/lanes.txt
