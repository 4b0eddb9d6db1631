#!/bin/bash
# short bench + kernel breakdown (used under gpurun while iterating)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_iter.json && python - <<PY
import json
d=json.load(open("gpurun_out/b_iter.json"))
print("fps", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
for k,v in d["kernel_breakdown"].items():
    if v["ms_per_step"] > 0.1: print(" ", k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
