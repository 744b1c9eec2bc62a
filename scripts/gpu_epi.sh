#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pair" --no-header -p no:cacheprovider 2>&1 | tail -4
python scripts/gemm_micro.py 2>&1 | grep -E "PAIR|BN=192" | cut -c1-120
for cfg in "X=1" "KJC_PAIR_GEMM=1"; do
echo "== $cfg"; env $cfg timeout 300 python bench.py --no-index --no-cpu --steps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
