#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused_ffn" --no-header -p no:cacheprovider 2>&1 | tail -5
python scripts/ffn_micro.py 2>&1 | head -3
TESTS="test_gpu_encoder" bash scripts/gpu_tests.sh | tail -4
timeout 300 python bench.py --no-index --no-cpu --steps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
