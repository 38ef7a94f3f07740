#!/bin/bash
# Fused build: one full-size CTA per SM against two half-size CTAs per SM (tile-per-CTA work list, staggered start).
set -u
P=temporally-consistent-stereo-matching_b200
run() { python bench.py --skip-cpu --skip-gpu-reference --skip-e2e --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('build_ms %.4f step %.4f checksum %.6f' % (d['phases']['build_ms'], d['ms_per_step'], d['checksum']))"; }
for cfg in "1 10" "2 6" "2 8" "2 5"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_FUSED_CTAS=$1 -DTCS_FUSED_CONV_WARPS=$2 \
       -I include -I $P/csrc -c $P/csrc/corr_build_fused.cu -o $P/build/corr_build_fused.o || exit 1
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  if [ "$1" = "2" ] && [ "$2" = "6" ]; then python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "build or fused or golden" 2>&1 | tail -2; fi
  if [ "$1" = "1" ]; then echo -n "== ctas 1 conv 10: "; run; echo -n "== ctas 1 conv 10 tile mode: "; TCS_FUSED_TILE_MODE=1 run; continue; fi
  for st in 0 5000 10000 20000; do echo -n "== ctas $1 conv $2 stagger $st: "; TCS_FUSED_STAGGER_NS=$st run; done
done
