#!/bin/bash
# Multi-GPU scaling run on ONE box (gpurun --gpus 8): cfg3 headline at 1/2/4/8 ranks and (SWEEP=1) the cfg-5 sweep at 2/4/8 ranks.
# Writes one JSON line per run under gpurun_out/.  RANKS="2 4 8" CFG3=0 SWEEP=1: the sweep alone.
set -u
P=29500
for N in ${RANKS:-1 2 4 8}; do
  P=$((P+1))
  [ "${CFG3:-1}" = "1" ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N \
      --steps 30 --warmup 5 --no-cpu-baseline --no-roofline --no-eager-reference --out gpurun_out/r02_scale_cfg3_n$N.json \
      > gpurun_out/r02_scale_cfg3_n$N.log 2>&1
  [ "${SWEEP:-0}" = "1" ] && [ "$N" != "1" ] || continue
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N \
      --config sweep --steps 10 --warmup 3 --no-cpu-baseline --out gpurun_out/r02_sweep_n$N.json \
      > gpurun_out/r02_sweep_n$N.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_scale_cfg3_n*.json')):
    d = json.load(open(f))
    print(f, 'n', d['n_gpus'], 'value %.1f e2e %.1f cached %.1f cached_e2e %.1f eager %.1f' % (
        d['value'], d['e2e']['value'], d['cached']['value'], d['cached_e2e']['value'], d['eager_api']['value']))
for f in sorted(glob.glob('gpurun_out/r02_sweep_n[248].json')):
    d = json.load(open(f))
    print(f, [round(c['value'], 1) for c in d['sweep']])
PY
