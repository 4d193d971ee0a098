#!/bin/bash
# round 2 evidence (1 GPU): smoke, the GPU suite, the bench line, ncu launch list + full captures of the metric kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.txt 2>&1; tail -1 gpurun_out/r2p_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.txt 2>&1; tail -3 gpurun_out/r2p_pytest.txt
timeout 900 python bench.py > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; tail -c 300 gpurun_out/r2p_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r2p_bench_ref.json 2>&1
BARGS="--steps 3 --warmup 3 --e2e-frames 2 --no-cpu --no-extra --sustained-s 0 --no-calls"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py $BARGS > gpurun_out/r2p_ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss5_r2 python bench.py $BARGS > gpurun_out/r2p_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control base --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss5_r2_baseclk python bench.py $BARGS > gpurun_out/r2p_ncu_full_base.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
