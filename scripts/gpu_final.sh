#!/bin/bash
# round-end check as the driver runs it: GPU tests, smoke, default bench line; then the ncu evidence of the training step
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-final10}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err
python - <<P
import json
for l in open("$OUT/bench_$TAG.json"):
    if l.startswith("{"):
        d = json.loads(l); t = d["train"]
        print("headline", round(d["ms_per_step"], 4), round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3))
        print("train", round(t["ms_per_step"], 4), round(t["value"]), "inline", round(t["pipeline"]["unpipelined_ms_per_step"], 4), "e2e", round(t["e2e"]["ms_per_step"], 4), t["pipeline"]["sm_partition"])
        print("c5", {k: round(v["ms_per_step"], 3) for k, v in d["config5_tempgru_beam4"].items() if isinstance(v, dict)}, "c1", round(d["config1_single_clip"]["latency_ms"], 3))
P
bash scripts/gpu_profile_r2c.sh r2 2>&1 | tail -4
