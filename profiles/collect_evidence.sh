#!/bin/bash
# Collects the measured evidence of a round on ONE B200 (run under gpurun from the repository root):
#     gpurun --timeout 1500 -- 'bash profiles/collect_evidence.sh r2final'
# Everything lands in gpurun_out/<tag>/; the summaries that are judged are then copied / condensed into profiles/
# in the build container (profiles/summarize_ncu_all.py, summarize_ncu_source.py).
set -u
TAG=${1:-evidence}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/smi.txt" 2>&1

# 1. parity suite through the C ABI
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/tests.txt"

# 2. bench: default workload (BASELINE configs[1]) with the CPU baseline and the secondary 8 x 512^2 measurement; reference arm
python bench.py --ops-out "$OUT/ops.json" > "$OUT/bench.json" 2> "$OUT/bench.err"
python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"

# 3. the other BASELINE configurations, in-graph timeline of one step
python tests/gpu_configs.py --out "$OUT/configs.json" > "$OUT/configs.log" 2>&1
python tests/gpu_timeline.py --out "$OUT/timeline.json" > "$OUT/timeline.txt" 2>&1

# 4. ncu: launch list of the bench command, every launch of one forward (speed-of-light + memory sections; the window is
#    wider than one forward — context conversion launches precede it — and summarize_ncu_all.py aligns on time_mlp_kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > "$OUT/ncu_l.log" 2>&1
python tests/gpu_profile_forward.py --iters 1 --names-out "$OUT/op_names.json" > /dev/null 2>&1
N=$(python -c "import json;print(len(json.load(open('$OUT/op_names.json'))))")
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
    -s "$((N - 4))" -c "$((N + 20))" -o "$OUT/prof_all" -f python tests/gpu_profile_forward.py --iters 2 > "$OUT/ncu_all.log" 2>&1

# 5. compute-sanitizer on smoke() (64 x 64 forward + 3-step DDIM loop checked against the oracle)
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file "$OUT/san_$tool.log" \
      python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/san_$tool.out" 2>&1
  echo "$tool rc=$? $(tail -1 "$OUT/san_$tool.log")" >> "$OUT/san_rc.txt"
done
ls -la "$OUT" | tail -30
