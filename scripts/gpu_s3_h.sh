#!/bin/bash
# session 3, call H: band rasteriser at 7 CTAs per SM under the 196 KB carve-out (60 KB of L1 instead of 28 KB); fresh capture of the band kernel
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
  for combo in "" "band_carveout=85 raster_ctas_per_sm=7" "raster_ctas_per_sm=7" "band_carveout=71 raster_ctas_per_sm=6" "band_carveout=100"; do
    echo "== [$combo]"
    opts=""; for o in $combo; do opts="$opts --opt $o"; done
    for w in cfg5 cfg1; do
      timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/h_$w.json 2> gpurun_out/h_$w.err; tail -1 gpurun_out/h_$w.err
      show gpurun_out/h_$w.json
    done
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_band_kernel -s 3 -c 1 -o gpurun_out/s3_band_cfg5_h -f \
     python bench.py --workload cfg5 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_band_h.log 2>&1; echo "cap rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_kernel -s 3 -c 1 -o gpurun_out/s3_bin_cfg5_h -f \
     python bench.py --workload cfg5 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_bin_h.log 2>&1; echo "cap rc=$?"
