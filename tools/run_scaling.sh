#!/bin/bash
# Scaling measurements of round 2 on N GPUs of one box (run through `gpurun --gpus N -- bash tools/run_scaling.sh N`):
#   cfg4 weak   (BASELINE.json configs[3]: 64x256x128 cells per GPU, Cx = 0.1)            - one process per GPU (torchrun) and multi_CUDA
#   cfg5 weak   (configs[4]: Cx = 0.5, rain mode, 64 columns per GPU)                      - one process per GPU
#   cfg5 strong (fixed 256x256x128 cells = 3.4e8 SDs split over the GPUs)                  - one process per GPU
# Each line of gpurun_out/r02_scale_*.json is bench.py's JSON line.
N=$1
OUT=gpurun_out
STEPS=${STEPS:-20}
WARM=${WARM:-5}
run() {  # name, extra args...
  name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup $WARM --no-cpu-baseline --no-alt "$@" > $OUT/r02_scale_${name}_n$N.json 2> $OUT/r02_scale_${name}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps $STEPS --warmup $WARM --no-cpu-baseline --no-alt "$@" > $OUT/r02_scale_${name}_n$N.json 2> $OUT/r02_scale_${name}_n$N.err
  fi
  echo "$name n=$N rc=$? $(tail -c 300 $OUT/r02_scale_${name}_n$N.err | tr '\n' ' ' | cut -c1-200)"
}
run cfg4_weak
if [ "$N" != "1" ] && [ -z "$SKIP_MULTICUDA" ]; then
  python bench.py --gpus $N --steps $STEPS --warmup $WARM --no-cpu-baseline --no-alt > $OUT/r02_scale_cfg4_weak_multicuda_n$N.json 2> $OUT/r02_scale_cfg4_weak_multicuda_n$N.err
  echo "cfg4_weak_multicuda n=$N rc=$?"
fi
run cfg5_weak --config cfg5
run cfg5_strong --config cfg5 --scaling strong --nx 256 --steps 10 --warmup 3
