#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2aj_pytest.txt 2>&1; tail -2 gpurun_out/r2aj_pytest.txt
timeout 600 python scripts/bench_all_kernels.py "7x7 sigma 1.2" > gpurun_out/r2aj_kern.txt 2>&1; cut -c1-200 gpurun_out/r2aj_kern.txt
python - <<'PY'
import rustcv_b200 as R
R.imgproc.init(0); R.imgproc.set_option("sepf32.windowed7", 1)
PY
RCV_OPT_WINDOWED7=1 timeout 600 python - <<'PY' 2>&1 | cut -c1-200
import sys, runpy
import rustcv_b200 as R
R.imgproc.init(0); R.imgproc.set_option("sepf32.windowed7", 1)
sys.argv = ["bench_all_kernels.py", "7x7 sigma 1.2, 1080p BGR"]
runpy.run_path("scripts/bench_all_kernels.py", run_name="__main__")
PY
