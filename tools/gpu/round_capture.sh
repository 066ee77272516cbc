#!/bin/bash
# GPU-box session of a round: tests, per-shape timing, launch list of the bench command, ncu --set full of three shapes.
# Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
cd "$(dirname "$0")/../.."
python -m pytest tests -m gpu -q -x > gpurun_out/gputests_r3.log 2>&1; tail -3 gpurun_out/gputests_r3.log
tools/time_shapes.sh > gpurun_out/r3_shapes.log 2>&1; grep TIME gpurun_out/r3_shapes.log
python bench.py > gpurun_out/bench_r3.log 2> gpurun_out/bench_r3.err; cut -c1-400 gpurun_out/bench_r3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_r2c_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:ofdm_link_fast -c 1 -f"
$NCU -o gpurun_out/r2c_n1024 python tools/profile_link.py --n 1024 --order 64 --launches 1 > gpurun_out/ncu_r2c_n1024.log 2>&1
$NCU -o gpurun_out/r2c_c1 python tools/profile_link.py --n 64 --order 4 --taps flat_fading --prefix 16 --eq ZF --launches 1 > gpurun_out/ncu_r2c_c1.log 2>&1
$NCU -o gpurun_out/r2c_c3 python tools/profile_link.py --n 64 --order 64 --launches 1 > gpurun_out/ncu_r2c_c3.log 2>&1
$NCU -o gpurun_out/r2c_n4096 python tools/profile_link.py --n 4096 --order 256 --launches 1 > gpurun_out/ncu_r2c_n4096.log 2>&1
ls -la gpurun_out/*r2c*
