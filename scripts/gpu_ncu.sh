#!/bin/bash
# One `ncu --set full` pass over the main kernels of a bench step, summarised on the GPU box (the .ncu-rep itself is kept
# only when it is small enough to travel back).  Usage: bash scripts/gpu_ncu.sh <tag> [kernel regex] [launch count]
cd "$(dirname "$0")/.."
TAG=${1:-r02}
REGEX=${2:-"k_virial_mass|k_count|k_od_count|k_ready|k_clip_bins_smem|k_hop2_heavy|k_cell_gather"}
COUNT=${3:-16}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --wide-galaxies 0"
mkdir -p gpurun_out
timeout 300 $CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" --launch-count $COUNT -f -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_run_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > /tmp/raw_$TAG.csv 2>/dev/null
python scripts/ncu_full_summary.py /tmp/raw_$TAG.csv > gpurun_out/ncu_full_$TAG.json
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv 2>/dev/null | grep -E "Stall|stall|Issue|Eligible|No Eligible|One or More" | head -400 > gpurun_out/ncu_details_$TAG.csv
sz=$(stat -c %s /tmp/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ]; then cp /tmp/prof_$TAG.ncu-rep gpurun_out/; fi
python - <<PY
import json
for e in json.load(open("gpurun_out/ncu_full_$TAG.json")):
    print(e["kernel"][:34].ljust(34), e.get("duration"), "| dram", e.get("dram_read"), e.get("dram_write"), "| L2", e.get("l2_hit_pct"),
          "| sm%", e.get("sm_pct_of_peak"), "| act", e.get("active_threads_per_inst"), "| warps%", e.get("warps_active_pct"), "| stall", e.get("top_stalls"))
PY
