#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2i_sweeps.txt
OP=sobel NF=64 timeout 300 python scripts/bench_sustained.py default sobel.band_rows=22 sobel.band_rows=30 sobel.band_rows=46 sobel.band_rows=62 sobel.band_rows=94 >> gpurun_out/r2i_sweeps.txt 2>&1
OP=sobel NF=128 timeout 300 python scripts/bench_sustained.py default sobel.band_rows=62 >> gpurun_out/r2i_sweeps.txt 2>&1
OP=warp NF=16 timeout 300 python scripts/bench_sustained.py default warp.tile_rows=32 warp.tile_rows=48 >> gpurun_out/r2i_sweeps.txt 2>&1
OP=warp NF=64 timeout 300 python scripts/bench_sustained.py default >> gpurun_out/r2i_sweeps.txt 2>&1
OP=gaussq5 timeout 300 python scripts/bench_sustained.py default gauss.band_rows=60 >> gpurun_out/r2i_sweeps.txt 2>&1
cat gpurun_out/r2i_sweeps.txt
