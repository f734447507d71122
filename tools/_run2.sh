nvidia-smi -L
python -m pytest tests/test_shared_source.py -m gpu -x -q -s 2>&1 | tail -8
python bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_n2.json
python bench.py --impl reference --gpus 2 --steps 10 --warmup 2 > gpurun_out/bench_n2_ref.json 2>> gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2_ref.json
