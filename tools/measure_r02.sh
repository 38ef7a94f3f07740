#!/bin/bash
# Round-2 measurement pass on one B200 (run through gpurun): GPU tests, the bench line, the launch list with per-launch
# DRAM bytes in the step's natural cache state, and one full ncu capture of the dominant kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
SMALL="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-gpu-reference"
$SMALL > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -c 900 --csv --log-file gpurun_out/r02_launches.csv $SMALL > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
$SMALL > gpurun_out/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:corr_lookup_r4x4o -s 200 -c 2 -o gpurun_out/r02_prof_lookup \
    $SMALL > gpurun_out/r02_ncu_lookup.log 2>&1
echo "lookup capture rc=$?"
cat gpurun_out/r02_bench.json
