python tools/fp32_peak.py 5 > gpurun_out/fp32_peak.log 2>&1
python tools/bench_frames.py > gpurun_out/frames.log 2>&1
python tools/sustained.py 15 > gpurun_out/sustained.log 2>&1
python tools/run_baseline_configs.py > gpurun_out/baseline_configs.log 2>&1
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ofdm_link_fast --launch-skip 2 -c 1 -f -o gpurun_out/r2_final_n1024 python tools/profile_link.py > gpurun_out/ncu_final.log 2>&1
cat gpurun_out/fp32_peak.log gpurun_out/frames.log gpurun_out/sustained.log | head -60
