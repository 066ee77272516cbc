python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests.log
P="python tools/profile_link.py"
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
} > gpurun_out/r2_times_v2.log 2>&1
python bench.py --no-c5 > gpurun_out/bench_v2.log 2> gpurun_out/bench_v2.err
tail -3 gpurun_out/gputests.log; grep TIME gpurun_out/r2_times_v2.log
