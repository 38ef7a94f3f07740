#!/bin/bash
# Fused build: one full-size CTA per SM against two half-size CTAs per SM on the shapes that take the fused kernel.
set -u
P=temporally-consistent-stereo-matching_b200
run() { python bench.py --skip-cpu --skip-gpu-reference --skip-e2e --steps 30 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('build_ms %.4f step %.4f value %.0f' % (d['phases']['build_ms'], d['ms_per_step'], d['value']))"; }
for ctas in 2 1; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_FUSED_CTAS=$ctas \
       -I include -I $P/csrc -c $P/csrc/corr_build_fused.cu -o $P/build/corr_build_fused.o || exit 1
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  echo -n "== ctas $ctas 540p x8: "; run
  echo -n "== ctas $ctas 540p x1: "; run --seqs-per-gpu 1
  echo -n "== ctas $ctas 540p x2: "; run --seqs-per-gpu 2
  echo -n "== ctas $ctas 480x640 x8: "; run --height 480 --width 640
  echo -n "== ctas $ctas 540p x8 bf16: "; run --precision bf16
done
