python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python tools/kernel_roofline.py > gpurun_out/kernel_roofline_r1e.jsonl 2>gpurun_out/kr.err; cut -c1-175 gpurun_out/kernel_roofline_r1e.jsonl; tail -5 gpurun_out/kr.err
