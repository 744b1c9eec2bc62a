#!/bin/bash
for v in parts2 parts3 parts2 parts3; do
  cp kjarni_b200/lib_$v.so kjarni_b200/libkjarni_cuda.so
  echo "== $v"; timeout 300 python bench.py --no-index --no-cpu --steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
