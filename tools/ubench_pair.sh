#!/bin/bash
# pairs N kch S boxmode do_mma same_slab steps cluster
U=tools/ubench_pair.bin
echo "== pure TMA (no MMA)"
$U 64 192 2 3 0 0 1 200 2
$U 64 192 2 5 0 0 1 200 2
$U 64 192 2 6 0 0 1 200 2
$U 64 192 2 6 1 0 1 200 2
$U 64 192 4 3 0 0 1 200 2
echo "== with MMA N=192"
$U 64 192 2 3 0 1 1 200 2
$U 64 192 2 5 0 1 1 200 2
$U 64 192 2 6 0 1 1 200 2
$U 64 192 4 3 0 1 1 200 2
$U 64 192 2 6 1 1 1 200 2
echo "== with MMA N=96"
$U 64 96 2 3 0 1 1 200 2
$U 64 96 2 6 0 1 1 200 2
$U 64 96 4 3 0 1 1 200 2
$U 64 96 2 6 1 1 1 200 2
echo "== with MMA N=256, 128"
$U 64 256 2 5 0 1 1 200 2
$U 64 128 2 6 0 1 1 200 2
