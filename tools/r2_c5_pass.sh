#!/bin/bash
# Pose-range partition on N GPUs: 1-rank vs N-rank parity on a converging chain graph, then the bounded full-size config-5 bench.
# usage (GPU box, repo root): bash tools/r2_c5_pass.sh N TAG [poses loops]
N=${1:-2}; TAG=${2:-r2f}; POSES=${3:-20000}; LOOPS=${4:-20}
O=gpurun_out; mkdir -p $O
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
rm -f $O/part_ref_${POSES}.npy
{ run 1 29531 tools/c5_partitioned.py $POSES $LOOPS; for W in 2 4 8; do if [ $W -le $N ]; then run $W $((29531+W)) tools/c5_partitioned.py $POSES $LOOPS; fi; done; } > $O/${TAG}_c5_parity.txt 2>&1
grep -E "world|max pose" $O/${TAG}_c5_parity.txt
run $N 29550 bench.py --config C5 --gpus $N --steps 1 --warmup 1 --no-e2e > $O/${TAG}_c5_full_n$N.json 2> $O/${TAG}_c5_full_n$N.err; echo "bench rc $?"
tail -c 600 $O/${TAG}_c5_full_n$N.err
