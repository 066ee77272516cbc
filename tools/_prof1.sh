set -x
python -m pytest tests/test_reference_examples_gpu.py -x -q 2>&1 | tail -3
P="python tools/profile_link.py"
NCU="ncu --set full --clock-control none --import-source on -k regex:ofdm_link_fast --launch-skip 2 -c 1 -f"
{
$P --time 20
$P --time 20 --points 16
$P --n 64 --order 4 --taps flat_fading --prefix 16 --eq ZF --time 20
$P --n 64 --order 64 --time 20
$P --n 256 --order 16 --time 20
$P --n 256 --order 64 --time 20
$P --n 128 --order 16 --time 20
$P --n 512 --order 16 --time 20
$P --n 2048 --order 64 --time 20
$P --n 4096 --order 256 --time 20
$P --n 64 --order 4 --taps default_multipath --modulator SC-OFDM --time 20
} > gpurun_out/r2_times.log 2>&1
$NCU -o gpurun_out/r2_n1024 $P > gpurun_out/ncu_n1024.log 2>&1
$NCU -o gpurun_out/r2_c1_n64 $P --n 64 --order 4 --taps flat_fading --prefix 16 --eq ZF > gpurun_out/ncu_c1.log 2>&1
$NCU -o gpurun_out/r2_c3_n64 $P --n 64 --order 64 > gpurun_out/ncu_c3.log 2>&1
$NCU -o gpurun_out/r2_n256 $P --n 256 --order 16 > gpurun_out/ncu_n256.log 2>&1
cat gpurun_out/r2_times.log | grep TIME
