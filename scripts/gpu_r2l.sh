#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2l_pytest_$i.txt 2>&1; tail -1 gpurun_out/r2l_pytest_$i.txt; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
