#!/bin/bash
# Sweep the host-model options of bench.py (memory format x cudnn.benchmark); prints one summary line per run.
for cl in 0 1; do for cb in 0 1; do
  python bench.py --steps 8 --warmup 6 --no-cpu-baseline --channels-last $cl --cudnn-benchmark $cb 2>/dev/null > /tmp/sweep_$cl$cb.json
  python - "$cl" "$cb" <<'PY'
import json, sys
cl, cb = sys.argv[1], sys.argv[2]
d = json.load(open(f"/tmp/sweep_{cl}{cb}.json"))
print("channels_last", cl, "cudnn_benchmark", cb, round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 2), "ms; hot path",
      d["roofline"]["hot_path_ms_per_step"], "ms; e2e", round(d["e2e"]["value"], 1))
PY
done; done
