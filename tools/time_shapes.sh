#!/bin/bash
# Every team shape of the fast kernel through tools/profile_link.py --time 20 (the table of profiles/r2_shapes.md)
cd "$(dirname "$0")/.."
P="python tools/profile_link.py --time 20"
$P --n 1024 --order 64
$P --n 1024 --order 64 --points 16
$P --n 64 --order 4 --taps flat_fading --prefix 16 --eq ZF
$P --n 64 --order 64 --eq ZF
$P --n 64 --order 64
$P --n 256 --order 16
$P --n 256 --order 64
$P --n 128 --order 16
$P --n 512 --order 16
$P --n 2048 --order 64
$P --n 4096 --order 256
$P --n 64 --order 4 --taps Lin-Phoong_P1 --prefix 3 --modulator SC-OFDM
$P --n 1024 --order 64 --eq ZF
$P --n 64 --order 64 --taps Lin-Phoong_P2 --prefix 1 --prefix-type ZERO
$P --n 64 --order 64 --taps Lin-Phoong_P2 --prefix 1
