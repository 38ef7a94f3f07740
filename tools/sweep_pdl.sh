#!/bin/bash
# Programmatic dependent launch of the lookup on / off (TCS_LOOKUP_PDL).  (The same on the warp chain's eight kernels was measured
# with this script and a TCS_WARP_PDL switch: warp phase 0.3021 -> 0.3008 ms, nothing at one frame; not kept.)
set -u
P=temporally-consistent-stereo-matching_b200
for cfg in "1 0" "0 0"; do
  set -- $cfg
  for f in corr_lookup warp; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_LOOKUP_PDL=$1 -DTCS_WARP_PDL=$2 \
         -I include -I $P/csrc -c $P/csrc/$f.cu -o $P/build/$f.o || exit 1
  done
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  echo "== lookup pdl $1 warp pdl $2"
  [ "$cfg" = "1 0" ] && timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sequence.py -x -q -m gpu 2>&1 | tail -2
  for rep in 1 2; do timeout 300 python bench.py --skip-cpu --skip-gpu-reference --skip-e2e --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('warp_ms %.4f lookups_ms %.4f step %.4f checksum %.6f' % (d['phases']['warp_ms'], d['phases']['lookups_ms'], d['ms_per_step'], d['checksum']))"; done
  timeout 300 python bench.py --skip-cpu --skip-gpu-reference --skip-e2e --steps 30 --seqs-per-gpu 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('x1: warp_ms %.4f lookups_ms %.4f step %.4f' % (d['phases']['warp_ms'], d['phases']['lookups_ms'], d['ms_per_step']))"
done
