#!/bin/bash
# Runs bench.py on every BASELINE.json shape (1 GPU) and prints one summary line per config.
set -u
out=gpurun_out/configs_r02.jsonl
: > $out
run() { echo "== $*" >&2; python bench.py --steps 10 --warmup 3 --skip-cpu --skip-gpu-reference "$@" 2>/dev/null | tail -1 >> $out; }
run --height 540 --width 960 --seqs-per-gpu 8
run --height 540 --width 960 --seqs-per-gpu 1 --skip-e2e
run --height 480 --width 640 --seqs-per-gpu 8
run --height 375 --width 1242 --seqs-per-gpu 8
run --height 375 --width 1242 --seqs-per-gpu 8 --precision bf16 --skip-e2e
run --height 1080 --width 1920 --seqs-per-gpu 2 --skip-e2e
run --height 1080 --width 1920 --seqs-per-gpu 2 --skip-e2e --mode alternate
python - <<'PY'
import json
for l in open("gpurun_out/configs_r02.jsonl"):
    d = json.loads(l); c = d["config"]; p = d["phases"]
    print("| %s | %s | %s | %d | %.0f | %.3f | %.3f / %.3f / %.3f | %.2f | %s |" % (c["workload"].split(" temporal")[0], c["mode"], c["precision"], c["seqs_per_gpu"], d["value"], d["ms_per_step"], p["build_ms"], p["warp_ms"], p["lookups_ms"], d["roofline"]["frac"], ("%.0f" % d["e2e"]["value"]) if d["e2e"] else "-"))
PY
