#!/bin/bash
# sustained runs of the final build (rare-deadlock check): edge_n / edge_m (32-channel stem variant), eager + graph, fp32 + odd sizes
set -u
mkdir -p gpurun_out
for cfg in "MODEL=edge_n MODE=detect_graph STEPS=3000" "MODEL=edge_n MODE=fwd STEPS=1500" "MODEL=edge_m B=32 MODE=detect_graph STEPS=1500" "MODEL=edge_m B=32 MODE=fwd STEPS=800" "MODEL=edge_s B=16 S=352 MODE=fwd STEPS=1500"; do
  echo "== $cfg"; env $cfg timeout 170 python scripts/stress.py 2>&1 | tail -3
done | tee gpurun_out/stress_final.log
