timeout 600 python -m pytest tests -m gpu -x -q -k "resampl or nvelope" > gpurun_out/r2_pytest17.log 2>&1; tail -n 3 gpurun_out/r2_pytest17.log
python tools/kernel_roofline.py --only Resampler | cut -c1-200
