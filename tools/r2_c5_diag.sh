#!/bin/bash
# Diagnostic: bounded config-5 bench on 2 GPUs at reduced size with a hard cap, per-rank stderr and Python tracebacks on abort.
POSES=${1:-400000}; CAP=${2:-70}; TAG=${3:-r2g}
O=gpurun_out; mkdir -p $O
for R in 0 1; do
  RANK=$R LOCAL_RANK=$R WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=29571 VUS_VERBOSE=1 \
  timeout -s ABRT $CAP python -X faulthandler -u bench.py --config C5 --gpus 2 --poses $POSES --steps 1 --warmup 1 --no-e2e --no-profile \
    > $O/${TAG}_diag_rank$R.json 2> $O/${TAG}_diag_rank$R.err &
done
wait
for R in 0 1; do echo "== rank $R"; grep -v "^  try\|^    pcg" $O/${TAG}_diag_rank$R.err | tail -25; done
head -c 400 $O/${TAG}_diag_rank0.json
