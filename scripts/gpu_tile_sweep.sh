#!/bin/bash
# On the GPU box: time every tile variant (scripts/build_tile_variants.py) with a short bench run, then run the
# stage-0 / overdensity parity tests against the fastest one.  Output: gpurun_out/tile_sweep.txt
mkdir -p gpurun_out
: > gpurun_out/tile_sweep.txt
for lib in fof_b200/variants/libfofb200_*.so; do
  name=$(basename $lib .so); name=${name#libfofb200_}
  FOFB_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --wide-galaxies 0 > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - "$name" >> gpurun_out/tile_sweep.txt <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/sweep_{name}.json"))
    k = d["roofline"]["kernel_ms_per_step"]
    print(name, round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3), k.get("k_count"), k.get("k_od_count"), d["config"]["clusters_final"], d["config"]["members_final"])
except Exception as e:
    print(name, "FAILED", e)
PY
done
cat gpurun_out/tile_sweep.txt
best=$(grep -v FAILED gpurun_out/tile_sweep.txt | sort -k2 -n | head -1 | cut -d' ' -f1)
echo "best=$best" | tee -a gpurun_out/tile_sweep.txt
export FOFB_LIB=$PWD/fof_b200/variants/libfofb200_$best.so
FILES="${FILES:-test_stage0_gpu test_parity_sizes_gpu test_gpu_golden test_fof_gpu test_edge_cases_gpu test_slabs_gpu}" bash scripts/gpu_diag.sh
