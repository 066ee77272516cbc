python -m pytest tests -m gpu -q 2>&1 | tail -8
python tests/fuzz_parity.py 300 77 > gpurun_out/fuzz300.log 2>&1; tail -1 gpurun_out/fuzz300.log; grep -c "ZERO" gpurun_out/fuzz300.log; grep "ZERO" gpurun_out/fuzz300.log | grep -c " gen "; grep MISMATCH gpurun_out/fuzz300.log | head
P="python tools/profile_link.py"
{
$P --n 64 --order 4 --taps flat_fading --prefix 16 --eq ZF --time 20
$P --n 64 --order 64 --eq ZF --time 20
$P --n 64 --order 64 --eq MMSE --time 20
$P --n 1024 --order 64 --eq ZF --time 20
$P --n 64 --order 64 --taps Lin-Phoong_P2 --prefix 1 --prefix-type ZERO --time 20
$P --n 64 --order 64 --taps Lin-Phoong_P2 --prefix 1 --prefix-type CYCLIC --time 20
} > gpurun_out/r2_times_v3.log 2>&1
grep "TIME\|fast_kernel" gpurun_out/r2_times_v3.log
