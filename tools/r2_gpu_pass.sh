#!/bin/bash
# One GPU pass: tests, bench lines (C3 spec / C1 / C2), ncu launch list and --set full captures of the hot kernels at C3 size.
# usage (from the repo root, on the GPU box): bash tools/r2_gpu_pass.sh TAG
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_gputest.log 2>&1; echo "pytest rc $?" >> $O/${TAG}_gputest.log
tail -3 $O/${TAG}_gputest.log
timeout 400 python bench.py --steps 3 --warmup 3 $BENCH_EXTRA > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc $?"
if [ "$3" == "all" ]; then
timeout 200 python bench.py --config C1 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c1.json 2> $O/${TAG}_bench_c1.err; echo "bench c1 rc $?"
timeout 200 python bench.py --config C2 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc $?"
fi
if [ "$2" == "list" ]; then
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches.csv \
  python tools/ncu_target.py 100000 > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc $?"
fi
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled \
  -k "regex:(SchurBlockBody|StereoPoseBody|StereoLmBody|NodeAsmBody|PairAsmBody|LinStereoTileBody|LinBody|ChunkFwdBody|ChunkBwdBody|BandMatvecBody|ChunkFactorBody|LmInvertBody|DampBody|LmBacksubBody)" \
  -c 36 -o /tmp/${TAG}_hot -f python tools/ncu_target.py 100000 > $O/${TAG}_ncu_hot.log 2>&1; echo "ncu full rc $?"
ncu -i /tmp/${TAG}_hot.ncu-rep --page details > $O/${TAG}_hot_details.txt 2>&1
ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > $O/${TAG}_hot_raw.csv 2>&1
SZ=$(stat -c %s /tmp/${TAG}_hot.ncu-rep 2>/dev/null || echo 0); echo "ncu-rep bytes $SZ"
if [ "$SZ" -gt 0 ] && [ "$SZ" -lt 40000000 ]; then cp /tmp/${TAG}_hot.ncu-rep $O/; fi
fi
ls -la $O | tail -12
