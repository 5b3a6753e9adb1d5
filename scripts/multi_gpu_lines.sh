#!/bin/bash
# BASELINE.json configs on N GPUs of one box (VERDICT r1 missing #4): the default workload (configs[1]/[2]), configs[3]
# (8 agents, use_ego_color, 512 envs per GPU) and the rasteriser sweep of configs[4], one torchrun line each.
#   bash scripts/multi_gpu_lines.sh N [tag]      -> gpurun_out/<tag>/n<N>_*.json
N=${1:-8}; TAG=${2:-mgpu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$N" -gt 1 ]; then RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; else RUN="python"; fi
$RUN bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > $OUT/n${N}_default.json 2> $OUT/n${N}_default.err
$RUN bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --num-agents 8 --batch-envs 512 --use-ego-color > $OUT/n${N}_configs3_a8_b512.json 2> $OUT/n${N}_configs3.err
for B in 4096 16384 65536; do
  $RUN bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --batch-envs $B > $OUT/n${N}_sweep_b${B}.json 2> $OUT/n${N}_sweep_b${B}.err
done
$RUN scripts/micro/d2h_ceiling.py > $OUT/n${N}_d2h_ceiling.json 2> $OUT/n${N}_d2h.err
for f in $OUT/n${N}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"] if "value" in d else d, "ms/step %.4f" % d.get("ms_per_step", float("nan")) if "ms_per_step" in d else "",
          "e2e %.4g" % d["e2e"]["value"] if d.get("e2e") else "", "raster frac %.3f fill frac %.3f" % (r.get("frac", 0), (r.get("fill_kernel_alone") or {}).get("frac", 0)) if r else "")
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
